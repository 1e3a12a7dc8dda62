# One GPU session end to end (run under gpurun from the repo root): parity tests, the bench line, the ncu launch list of the
# same command and one --set full capture of the dominant kernel.  Summaries go to profiles/ with bench/ncu_summary.py.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
python bench.py --steps 20 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_1gpu.log').read().strip().splitlines()[-1])
    print(d['value'], d['n_gpus'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_1gpu.log').read()[-1500:]); print(open('gpurun_out/bench_1gpu.err').read()[-2500:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
MSC_JIT_DUMP_DIR=gpurun_out/jitsrc ncu --set full --clock-control none --import-source on -k regex:msc_jit_dense -s 4 -c 1 -f -o gpurun_out/prof python bench.py --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
# here afterwards:  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv; ncu -i ... --page source --csv > src.csv; python bench/ncu_summary.py raw.csv src.csv ...
