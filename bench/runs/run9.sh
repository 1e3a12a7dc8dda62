mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 0 -c 1 -f -o gpurun_out/prof_hash python bench/bench_configs.py --sf 10 --reps 1 --only config4 > gpurun_out/ncu_hash.log 2>&1; echo "ncu exit $?"
