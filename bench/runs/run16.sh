mkdir -p gpurun_out
python bench.py --steps 20 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
MSC_JIT_DUMP_DIR=gpurun_out/jitsrc ncu --set full --clock-control none --import-source on -k regex:msc_jit -s 4 -c 1 -f -o gpurun_out/prof_r1n python bench.py --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
ls gpurun_out/jitsrc
