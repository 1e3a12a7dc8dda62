# scratch driver for gpurun sessions
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
python bench.py --steps 20 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_1gpu.log').read().strip().splitlines()[-1])
    print(d['value'], d['n_gpus'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['ms_per_step'], d['config']['wall_ms_per_step'], d['gpu_launches'], d['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_1gpu.log').read()[-1500:]); print(open('gpurun_out/bench_1gpu.err').read()[-2500:])
PY
ncu --set full --clock-control none --import-source on -k regex:regvm -c 1 -f -o gpurun_out/prof_r1k python bench.py --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
