# scratch driver for gpurun sessions
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest exit $?; tail -5 gpurun_out/pytest.log
for m in 1 0; do
MSC_SCAN_MASKED=$m python bench.py --steps 10 --e2e-steps 2 > gpurun_out/bench_sf15_m$m.log 2>&1; echo "masked=$m exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_sf15_m$m.log').read().strip().splitlines()[-1])
    print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['launch'], d['ms_per_step'], d['e2e'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_sf15_m$m.log').read()[-1500:])
PY
done
MSC_SCAN_STAGES=2 python bench.py --steps 10 --e2e-steps 1 > gpurun_out/bench_sf15_m1s2.log 2>&1; echo "stages2 exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_sf15_m1s2.log').read().strip().splitlines()[-1])
print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['launch'])"
ncu --set full --clock-control none --import-source on -k regex:regvm -c 1 -f -o gpurun_out/prof_r1h python bench.py --sf 4 --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
