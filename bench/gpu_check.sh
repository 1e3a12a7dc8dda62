# One GPU session end to end (run under gpurun from the repo root, ONE GPU): parity tests, both bench arms, the ncu launch
# list of the bench command and --set full captures of the kernels DESIGN.md names.  Summaries go to profiles/ with
# bench/ncu_summary.py (run here afterwards on the exported CSVs).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/r2_pytest.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1; echo "reference arm exit $?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.log 2> gpurun_out/r2_bench_1gpu.err; echo "bench exit $?"
python bench/show_bench.py gpurun_out/r2_bench_1gpu.log
# launch list of the SAME command (cold cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches exit $?"
# --set full, one launch each: the headline kernel (Q1 prepared pass), the fused probe scan of config 5, the streaming aggregate of
# config 4, the join table build, the shared-memory-cell dense kernel (7 groups x 7 aggregates) and the hash scan behind CTA-local tables
export MSC_JIT_DUMP_DIR=gpurun_out/jitsrc_r2 MSC_JIT_CACHE=gpurun_out/jitcache_r2
ncu --set full --clock-control none --import-source on -k regex:msc_jit_dense -s 6 -c 1 -f -o gpurun_out/r2_prof_q1 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-extras > gpurun_out/r2_ncu_q1.log 2>&1; echo "ncu q1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:msc_jit_dense -s 2 -c 1 -f -o gpurun_out/r2_prof_join python bench/cfg_probe.py --config 5 --reps 2 > gpurun_out/r2_ncu_join.log 2>&1; echo "ncu join exit $?"
ncu --set full --clock-control none --import-source on -k regex:msc_jit_runs -s 1 -c 1 -f -o gpurun_out/r2_prof_runs python bench/cfg_probe.py --config 4 --reps 2 > gpurun_out/r2_ncu_runs.log 2>&1; echo "ncu runs exit $?"
ncu --set full --clock-control none -k regex:join_build8_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_build python bench/cfg_probe.py --config 5 --reps 2 > gpurun_out/r2_ncu_build.log 2>&1; echo "ncu build exit $?"
ncu --set full --clock-control none --import-source on -k regex:msc_jit_dense -s 2 -c 1 -f -o gpurun_out/r2_prof_cells python bench/cfg_probe.py --config 6 --reps 2 > gpurun_out/r2_ncu_cells.log 2>&1; echo "ncu cells exit $?"
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_lhash python bench/cfg_probe.py --config 7 --reps 2 > gpurun_out/r2_ncu_lhash.log 2>&1; echo "ncu lhash exit $?"
for n in q1 join runs build cells lhash; do
  ncu -i gpurun_out/r2_prof_$n.ncu-rep --page raw --csv > gpurun_out/r2_prof_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof_$n.ncu-rep --page source --csv > gpurun_out/r2_prof_${n}_src.csv 2>/dev/null
  rm -f gpurun_out/r2_prof_$n.ncu-rep
done
ls gpurun_out/jitcache_r2 2>/dev/null | head
