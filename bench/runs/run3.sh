mkdir -p gpurun_out
python -m pytest tests/test_gpu_jit.py -x -q > gpurun_out/pytest_jit.log 2>&1; echo pytest_jit exit $?; tail -5 gpurun_out/pytest_jit.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
MSC_SCAN_JIT=2 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_alljit.log 2>&1; echo pytest_alljit exit $?; tail -3 gpurun_out/pytest_alljit.log
python bench.py --steps 20 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
MSC_JIT_DUMP_DIR=gpurun_out/jitsrc ncu --set full --clock-control none --import-source on -k regex:msc_jit -c 1 -f -o gpurun_out/prof_r1m python bench.py --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
