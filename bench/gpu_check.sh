# scratch driver for gpurun sessions
python bench/step_probe.py --sf 15 --steps 50 2>&1 | tail -12
python bench.py --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/plain_sf15.log 2>&1; echo "plain exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:regvm -c 1 -f -o gpurun_out/prof_r1j python bench.py --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"
