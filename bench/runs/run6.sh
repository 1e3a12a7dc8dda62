mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg.csv python bench/bench_configs.py --sf 10 --reps 1 --only config4,config5 > gpurun_out/ncu_cfg.log 2>&1; echo "ncu cfg exit $?"
tail -2 gpurun_out/ncu_cfg.log | cut -c1-300
