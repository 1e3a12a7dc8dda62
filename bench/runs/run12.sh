mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo pytest exit $?; tail -25 gpurun_out/pytest_multi.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_2gpu.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['roofline']['launch'])
PY
tail -5 gpurun_out/bench_2gpu.err | cut -c1-300
