"""Summarise one `ncu --set full` capture (its `--page raw --csv` and `--page source --csv` exports) into the JSON entry
format of profiles/r01_scan_kernel_ncu_summary.json.

    ncu -i prof.ncu-rep --page raw --csv > prof_raw.csv; ncu -i prof.ncu-rep --page source --csv > prof_src.csv
    python bench/ncu_summary.py prof_raw.csv prof_src.csv "<capture label>" "<note>" <rows scanned> > entry.json
"""
import collections
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def main() -> None:
    raw, src, label, note, nrows = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = dict(zip(hdr, zip(vals, units)))
    entry = {"capture": label, "note": note, "kernel": get["Kernel Name"][0], "rows_scanned": nrows}
    for k in KEEP:
        if k in get:
            v, u = get[k]
            entry[k] = f"{v} {u}".strip()
    stalls = {}
    issued = float(get["smsp__inst_executed.sum"][0]) if "smsp__inst_executed.sum" in get else 0.0
    for h, (v, _) in get.items():
        # warp-level stall breakdown: average warp cycles per issued instruction, by reason
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
            if float(v) >= 0.05:
                stalls[name] = round(float(v), 2)
    entry["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items()))
    entry["warp_instructions_executed"] = int(issued)
    entry["warp_instructions_per_row"] = round(issued / nrows, 4)
    # opcode mix from the source page (SASS rows carry "Instructions Executed")
    srows = list(csv.reader(open(src)))
    hi = next((i for i, r in enumerate(srows) if "Source" in r and "Instructions Executed" in r), None)
    if hi is not None:
        sh = srows[hi]
        si, ei = sh.index("Source"), sh.index("Instructions Executed")
        mix = collections.Counter()
        for r in srows[hi + 1:]:
            try:
                n = float(r[ei])
            except (ValueError, IndexError):
                continue
            toks = r[si].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            mix[op.rstrip(";")] += n
        total = sum(mix.values()) or 1.0
        entry["top_opcodes_pct"] = {k: round(100 * v / total, 1) for k, v in mix.most_common(14)}
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
