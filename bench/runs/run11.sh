mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -4 gpurun_out/pytest.log
python bench.py --steps 20 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"; tail -3 gpurun_out/bench_1gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1gpu.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['config']['parity_check'])
PY
python bench/step_probe.py 2>&1 | tail -8
