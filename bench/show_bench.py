#!/usr/bin/env python
"""Print the interesting numbers of bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    print(f"{f}: N={d['n_gpus']} value {d['value']:.4g} rows/s, {d['ms_per_step']:.4f} ms/step, kernel {d['roofline']['kernel_ms']:.4f} ms, "
          f"frac {d['roofline']['frac']:.3f}, launches {d['gpu_launches']}")
    e = d["e2e"]
    print(f"   e2e {e['value']:.4g} rows/s, {e['ms_per_step']:.2f} ms, h2d {e['h2d_gbs']:.1f} GB/s, d2h bytes {e['d2h_bytes_per_step']}")
    print(f"   cpu {d['cpu_baseline']['value']:.4g} ({d['cpu_baseline']['cores']} thr)  py-engine {(d.get('cpu_baseline_python') or {}).get('value')}")
    for k, v in d.get("extra", {}).items():
        if "error" in v:
            print(f"   {k}: ERROR {v['error'][:300]}")
            continue
        if k == "midcard":
            for kk, vv in v.items():
                if isinstance(vv, dict):
                    print(f"   midcard {kk}: {vv['ms']:.4f} ms, kernel {vv['kernel_ms']:.4f} ms (frac {vv['kernel_frac_of_peak_all_gpus']:.3f}), {vv['groups']} groups, "
                          f"{vv['agg_mode']} kind {vv['scan_kind']} local slots {vv['hash_local_slots']} | {vv['parity_check'][:50]}")
            continue
        ms = v.get("ms", v.get("ms_per_step"))
        frac = v.get("frac_of_peak_all_gpus")
        print(f"   {k}: {ms:.4f} ms" + (f", frac {frac:.3f}" if frac is not None else "") + f", exchange {v.get('exchange')}, wall {v.get('bench_wall_s')} s"
              + (f", plan {v.get('plan')}" if 'plan' in v else "") + (f" | {v.get('parity_check','')[:60]}" if 'parity_check' in v else ""))
