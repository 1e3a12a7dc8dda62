#!/usr/bin/env python
"""Build profiles/r02_ncu_summary.json and profiles/r02_traffic.json from the CSV exports bench/gpu_check.sh leaves in
gpurun_out/ (`ncu -i X.ncu-rep --page raw|source --csv`), one entry per captured kernel (bench/ncu_summary.py is the digest of
one capture).

    python bench/profile_digest.py
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3,
        "second": 1e6, "s": 1e6}
SF10_ROWS, SF15_ROWS, BUILD_KEYS = 59999625, 89998471, 2276829
CAPTURES = [
    # name, label, note, rows, traffic workload (None: summary only), algorithmic bytes per row
    ("q1", "r02 prof_q1: msc_jit_dense, TPC-H Q1 prepared pass (fused finish), lineitem sf15, 1 x B200",
     "the headline kernel: scan + filter + GROUP BY + final projection in one launch; DRAM traffic = algorithmic bytes", SF15_ROWS,
     {"sf_per_gpu": 15.0, "layout": "native", "rows_per_gpu": SF15_ROWS, "bytes_per_row_scanned": 25}, 25),
    ("join", "r02 prof_join: msc_jit_dense carrying the join probe (MSC_OP_PROBE), config 5, lineitem sf10 probe side",
     "LIKE lookup, probe of the 2.3 M-key compact table (presence bitmap word tested in the resolve half, 8-byte slots, L2 evict-last), survivors queued "
     "per warp, aggregate per survivor; 5 resident CTAs per SM", SF10_ROWS,
     {"config": "join (extra.join)", "sf": 10.0, "probe_rows": SF10_ROWS, "bytes_per_row_scanned": 9}, 9),
    ("runs", "r02 prof_runs: msc_jit_runs, streaming aggregate over sorted runs, config 4 (GROUP BY l_orderkey), lineitem sf10",
     "no hash table: run heads by comparison, run numbers by warp scan (tile bases fetched one tile ahead); every run's cell is STORED by the segment "
     "holding its first row, leading rows are added behind the warp's barrier, tiles carry into earlier runs through carry cells: no identity fill",
     SF10_ROWS, {"config": "highcard (extra.highcard)", "sf": 10.0, "rows": SF10_ROWS, "bytes_per_row_scanned": 12}, 12),
    ("build", "r02 prof_build: join_build8_kernel, compact join table build over the filtered orders side (2.3 M keys), config 5",
     "CAS on 8-byte slots + presence bitmap", BUILD_KEYS, None, 8),
    ("cells", "r02 prof_cells: msc_jit_dense with its accumulators in shared memory (a copy of every cell per thread), GROUP BY l_shipmode x 7 aggregates, lineitem sf10",
     "49 cells: past 32 the specialised kernel updates the row's own group's cells (LDS / DADD / STS) instead of every group's register under a predicate; "
     "3 CTAs per SM (67.8 KB each)", SF10_ROWS,
     {"config": "midcard l_shipmode (extra.midcard)", "sf": 10.0, "rows": SF10_ROWS, "bytes_per_row_scanned": 13}, 13),
    ("lhash", "r02 prof_lhash: scan_kernel<8, MODE_HASH> behind CTA-local tables, GROUP BY l_quantity (50 groups, FLOAT key) x 7 aggregates, lineitem sf10",
     "512-slot shared-memory table per CTA (find-or-insert + red.shared), folded into the global table when the CTA is done; without it the same scan "
     "takes 18.2 ms of same-address global atomics", SF10_ROWS,
     {"config": "midcard l_quantity (extra.midcard)", "sf": 10.0, "rows": SF10_ROWS, "bytes_per_row_scanned": 12}, 12),
]


def quantity(raw: dict, key: str) -> float:
    value, unit = raw[key]
    return float(value.replace(",", "")) * UNIT[unit]


def main() -> None:
    summary = {"note": "Summaries of `ncu --set full --clock-control none --import-source on` captures (gpurun, 1 x B200, round 2; bench/gpu_check.sh is the "
                       "recipe, bench/ncu_summary.py the digest of one capture, bench/profile_digest.py wrote this file). Times under the profiler are "
                       "cold-cache and serialised: use them for shares and traffic, the live CUDA-event numbers are in bench.py's JSON line.",
               "captures": []}
    traffic = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` captures (gpurun, 1 x B200, round 2); "
                       "bench.py copies `traffic_bytes` of the entry whose workload and kernel match into roofline.traffic", "captures": []}
    for name, label, note, nrows, workload, bytes_per_row in CAPTURES:
        raw_csv, src_csv = OUT / f"r2_prof_{name}_raw.csv", OUT / f"r2_prof_{name}_src.csv"
        if not raw_csv.exists():
            print("missing", raw_csv, file=sys.stderr)
            continue
        entry = json.loads(subprocess.run([sys.executable, str(ROOT / "bench" / "ncu_summary.py"), str(raw_csv), str(src_csv), label, note, str(nrows)],
                                          check=True, capture_output=True, text=True).stdout)
        summary["captures"].append(entry)
        rows = list(csv.reader(open(raw_csv)))
        raw = dict(zip(rows[0], zip(rows[2], rows[1])))
        read, write = quantity(raw, "dram__bytes_read.sum"), quantity(raw, "dram__bytes_write.sum")
        if workload is not None:
            traffic["captures"].append({"workload": workload, "kernel": entry["kernel"], "capture": label.split(":")[0] + " (profiles/r02_ncu_summary.json)",
                                        "dram_bytes_read": int(read), "dram_bytes_write": int(write), "traffic_bytes": int(read + write),
                                        "algorithmic_bytes": nrows * bytes_per_row, "gpu_time_duration_us": round(quantity(raw, "gpu__time_duration.sum"), 3)})
    (ROOT / "profiles" / "r02_ncu_summary.json").write_text(json.dumps(summary, indent=1) + "\n")
    (ROOT / "profiles" / "r02_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    for t in traffic["captures"]:
        print(t["capture"], t["kernel"], f"{t['gpu_time_duration_us']:.1f} us", f"traffic {t['traffic_bytes'] / 1e6:.0f} MB vs algorithmic {t['algorithmic_bytes'] / 1e6:.0f} MB")


if __name__ == "__main__":
    main()
