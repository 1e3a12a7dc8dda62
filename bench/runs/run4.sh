mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
MSC_SCAN_JIT=2 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_alljit.log 2>&1; echo pytest_alljit exit $?; tail -15 gpurun_out/pytest_alljit.log
python bench/bench_configs.py --sf 10 --reps 5 --jit never > gpurun_out/configs_nojit.log 2>&1; echo cfg exit $?; cat gpurun_out/configs_nojit.log | cut -c1-420
python bench/bench_configs.py --sf 10 --reps 5 --jit always > gpurun_out/configs_jit.log 2>&1; echo cfg exit $?; cat gpurun_out/configs_jit.log | cut -c1-420
python bench.py --steps 20 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1gpu.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['traffic_source'])
PY
python bench/step_probe.py 2>&1 | tail -12
