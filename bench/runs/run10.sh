mkdir -p gpurun_out
python -m pytest tests/test_gpu_jit.py tests/test_gpu_dense_variants.py -x -q > gpurun_out/pytest_jit.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest_jit.log
python bench.py --steps 20 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1gpu.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'])
PY
python bench/step_probe.py 2>&1 | tail -8
