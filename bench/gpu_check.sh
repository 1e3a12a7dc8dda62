# scratch driver for gpurun sessions
python -m pytest tests/test_gpu_multi.py tests/test_gpu_dense_variants.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/pytest.log
for n in 1 2; do
if [ $n = 1 ]; then python bench.py --steps 20 > gpurun_out/bench_${n}gpu.log 2> gpurun_out/bench_${n}gpu.err; else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_${n}gpu.log 2> gpurun_out/bench_${n}gpu.err; fi; echo "bench$n exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${n}gpu.log').read().strip().splitlines()[-1])
    print(d['value'], d['n_gpus'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['ms_per_step'], d['config']['wall_ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['roofline']['traffic'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_${n}gpu.log').read()[-1500:]); print(open('gpurun_out/bench_${n}gpu.err').read()[-2500:])
PY
done
