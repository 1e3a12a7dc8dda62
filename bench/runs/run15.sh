mkdir -p gpurun_out
MSC_BENCH_KEEP=1 python bench/bench_configs.py --sf 10 --reps 5 --only q1_by --jit never > gpurun_out/cfg_q1s_nojit.log 2>&1; cat gpurun_out/cfg_q1s_nojit.log | cut -c1-500
python bench/bench_configs.py --sf 10 --reps 5 --only q1_by --jit always > gpurun_out/cfg_q1s_jit.log 2>&1; cat gpurun_out/cfg_q1s_jit.log | cut -c1-500
