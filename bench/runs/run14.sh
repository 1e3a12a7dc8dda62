mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -4 gpurun_out/pytest.log | cut -c1-300
MSC_SCAN_JIT=2 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_alljit.log 2>&1; echo pytest_alljit exit $?; tail -4 gpurun_out/pytest_alljit.log | cut -c1-300
MSC_BENCH_KEEP=1 python bench/bench_configs.py --sf 10 --reps 5 --only q1_by --jit never > gpurun_out/cfg_q1s_nojit.log 2>&1; cat gpurun_out/cfg_q1s_nojit.log | cut -c1-400
python bench/bench_configs.py --sf 10 --reps 5 --only q1_by --jit always > gpurun_out/cfg_q1s_jit.log 2>&1; cat gpurun_out/cfg_q1s_jit.log | cut -c1-400
