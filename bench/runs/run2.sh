python -m pytest tests/test_gpu_jit.py -x -q > gpurun_out/pytest_jit.log 2>&1; echo pytest_jit exit $?; tail -15 gpurun_out/pytest_jit.log
timeout 300 python bench.py --steps 20 --e2e-steps 3 > gpurun_out/bench_jit.log 2> gpurun_out/bench_jit.err; echo "bench exit $?"; tail -3 gpurun_out/bench_jit.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_jit.log').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['kernel'], d['roofline']['launch'], d['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_jit.log').read()[-1500:])
PY
MSC_SCAN_JIT=2 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_alljit.log 2>&1; echo pytest_alljit exit $?; tail -5 gpurun_out/pytest_alljit.log
