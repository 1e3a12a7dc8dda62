# scratch driver for gpurun sessions
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
python bench.py --steps 20 > gpurun_out/bench_sf15.log 2> gpurun_out/bench_sf15.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_sf15.log').read().strip().splitlines()[-1])
    print(d['value'], d['n_gpus'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['ms_per_step'], d['config']['wall_ms_per_step'], d['gpu_launches'], d['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_sf15.log').read()[-1500:]); print(open('gpurun_out/bench_sf15.err').read()[-2500:])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_2gpu.log').read().strip().splitlines()[-1])
    print(d['value'], d['n_gpus'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['ms_per_step'], d['config']['wall_ms_per_step'], d['gpu_launches'], d['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_2gpu.log').read()[-1500:]); print(open('gpurun_out/bench_2gpu.err').read()[-2500:])
PY
