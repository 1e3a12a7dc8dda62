# scratch driver for gpurun sessions
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
for sf in 15 1; do
python bench.py --sf $sf --steps 20 > gpurun_out/bench_sf$sf.log 2> gpurun_out/bench_sf$sf.err; echo "bench sf$sf exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_sf$sf.log').read().strip().splitlines()[-1])
    print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['launch'], d['ms_per_step'], d['config']['wall_ms_per_step'], d['gpu_launches'], d['e2e'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_sf$sf.log').read()[-1500:]); print(open('gpurun_out/bench_sf$sf.err').read()[-1500:])
PY
done
