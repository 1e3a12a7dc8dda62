mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "bench8 exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_8gpu.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['roofline']['launch'], d['config']['parity_check'], d['e2e']['value'])
PY
tail -5 gpurun_out/bench_8gpu.err | cut -c1-300
