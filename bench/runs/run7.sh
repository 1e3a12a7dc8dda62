mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
python bench/bench_configs.py --sf 10 --reps 5 --only config4,config5 > gpurun_out/configs2.log 2>&1; echo cfg exit $?; cat gpurun_out/configs2.log | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg4.csv python bench/bench_configs.py --sf 10 --reps 1 --only config4 > gpurun_out/ncu_cfg4.log 2>&1; echo "ncu cfg exit $?"
