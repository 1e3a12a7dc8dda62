#!/usr/bin/env python
"""TPC-H Q1 benchmark of the CUDA engine (BASELINE.json: "TPC-H Q1 rows/sec & scan GB/s vs HBM peak").

    python bench.py --gpus 1 --steps 20 --warmup 3 [--sf 15] [--layout native|wide]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm: the reference's native path, restated (oracle/q1_port.c)

A step is one pass of the hot path over the whole (per-rank) lineitem: fused scan + filter +
GROUP BY (msc_scan_aggregate) followed by the AVG projection over the 3 result groups
(msc_scan_project).  `value` is measured with the columns resident in HBM (CUDA events on the
library's stream, max over ranks); `e2e` re-ingests the BlockFile image from pinned host memory
every step and reads the result back.  Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "tests", ROOT / "bench"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))
os.environ["TZ"] = "UTC"
time.tzset()

Q1_NATIVE_BYTES_PER_ROW = 8 + 4 * 4 + 1   # i64 shipdate + 4 x f32 + u8 returnflag code
Q1_WIDE_BYTES_PER_ROW = 8 + 4 * 8 + 4     # north_star layout: i64 + 4 x f64 + u32 code (44 B/row)
Q1_DISK_BYTES_PER_ROW = 8 + 4 * 4 + 2     # what is copied host->device: + u8 length and 1 byte per flag


def parse_args() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["cuda", "reference"], default="cuda")
    ap.add_argument("--sf", type=float, default=15.0, help="lineitem scale factor PER GPU (sf15 ~ 90M rows)")
    ap.add_argument("--layout", choices=["native", "wide"], default="native")
    ap.add_argument("--e2e-steps", type=int, default=7)
    ap.add_argument("--keep", action="store_true", help="keep the generated table")
    return ap.parse_args()


def peaks() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int) -> None:
        self.device = device
        self.proc: subprocess.Popen | None = None
        self.lines: list[str] = []

    def __enter__(self) -> "ClockSampler":
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self) -> None:
        assert self.proc and self.proc.stdout
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc: object) -> None:
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


class NumaLocal:
    """Run a block on the CPUs next to GPU `index` (NVML's ideal affinity), so that memory first touched inside it --
    the pinned BlockFile image -- lands on the GPU's NUMA node; the previous affinity comes back afterwards (the CPU
    baseline must see all host cores).  With eight ranks reading their images from one socket's memory the host side
    of the PCIe copies, not the links, set the e2e rate."""

    def __init__(self, index: int) -> None:
        self.index, self.old, self.cpus = index, None, None

    def __enter__(self) -> "NumaLocal":
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            old = os.sched_getaffinity(0)
            if cpus & old and (cpus & old) != old:
                os.sched_setaffinity(0, cpus & old)
                self.old, self.cpus = old, sorted(cpus & old)
        except Exception:  # noqa: BLE001  (no NVML, no permission: keep the default placement)
            self.old = None
        return self

    def __exit__(self, *exc) -> None:  # noqa: ANN002
        if self.old is not None:
            os.sched_setaffinity(0, self.old)


def captured_traffic(sf: float, layout: str, rows: int, bytes_per_row: int, kernel: str) -> tuple[float | None, str | None]:
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this very workload and kernel (profiles/)."""
    path = ROOT / "profiles" / "r01_traffic.json"
    if not path.exists():
        return None, None
    for cap in json.loads(path.read_text()).get("captures", []):
        w = cap["workload"]
        same_kernel = kernel.startswith(cap["kernel"].split("::")[-1])
        if same_kernel and (w["sf_per_gpu"], w["layout"], w["rows_per_gpu"], w["bytes_per_row_scanned"]) == (sf, layout, rows, bytes_per_row):
            return float(cap["traffic_bytes"]), f"profiles/r01_traffic.json ({cap['capture']})"
    return None, None


def kernel_name(rows_per_thread: int, kind: int = 0, regs: int = 0) -> str:
    """msc_stats.last_scan_kind 2 = query-specialised kernel (csrc/jit.cu); otherwise last_scan_rows_per_thread is R of
    scan_kernel<R, MODE_DENSE>, or -(8 + 100 * NG) for the regvm variants."""
    if kind == 2:
        return f"msc_jit_dense (specialised for this query by NVRTC, {rows_per_thread} rows per lane, {regs} registers)"
    if rows_per_thread >= 0:
        return f"scan_kernel<R={rows_per_thread}, MODE_DENSE>"
    ng, rows = divmod(-rows_per_thread, 100)
    return f"regvm_dense_kernel_ng{ng} ({rows} rows per lane)"


def table_path(sf: float, rank: int) -> Path:
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    folder = base / f"minispark_b200_bench_{os.getuid()}"
    folder.mkdir(parents=True, exist_ok=True)
    return folder / f"lineitem_q1_sf{sf:g}_rank{rank}.bin"


def ensure_table(sf: float, rank: int) -> tuple[Path, float]:
    import gen_tpch

    path = table_path(sf, rank)
    t0 = time.perf_counter()
    if not path.exists():
        tmp = path.with_suffix(".tmp")
        gen_tpch.write_table(tmp, "lineitem", sf=sf, columns=gen_tpch.Q1_COLUMNS, seed=1234 + 1000 * rank)
        tmp.rename(path)
    return path, time.perf_counter() - t0


def run_q1_port(path: Path, threads: int, max_blocks: int, wire: int) -> dict:
    exe = ROOT / "oracle" / "build" / "q1_port"
    if not exe.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, stdout=subprocess.DEVNULL)
    out = subprocess.run([str(exe), str(path), str(threads), str(max_blocks), str(wire)], check=True, capture_output=True, text=True)
    return json.loads(out.stdout)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def dist_env() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# ----------------------------------------------------------------------------------------------------
def reference_arm(args: argparse.Namespace) -> None:
    """The reference's own CPU implementation of the path (ThreadEngine steps, oracle/q1_port.c) on all
    host threads.  Rank 0 alone runs; a step is one pass over a bounded block sample of the same table."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    path, _ = ensure_table(args.sf, 0)
    threads = host_threads()
    probe = run_q1_port(path, threads, 4, 1)
    per_block = max(probe["seconds"] / max(probe["blocks"], 1), 1e-4)
    budget_s = 120.0 / max(args.steps + args.warmup, 1)  # keep the whole arm within a few minutes
    blocks = int(max(1, min(budget_s / per_block, 1e9)))
    for _ in range(args.warmup):
        run_q1_port(path, threads, blocks, 1)
    times, rows = [], 0
    for _ in range(args.steps):
        r = run_q1_port(path, threads, blocks, 1)
        times.append(r["seconds"])
        rows = r["rows"]
        blocks_used = r["blocks"]
    total = sum(times)
    value = rows * args.steps / total
    line = {
        "impl": "reference", "metric": "tpch_q1_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TPC-H Q1 (examples/benchmark.py:51-68) on synthetic lineitem sf{args.sf:g}", "sf_per_gpu": args.sf,
                   "rows_per_step": rows, "columns_in_file": 6},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": f"{blocks_used} row-blocks ({rows} rows) per step, oracle/q1_port.c: per-block jobs on {threads} threads, "
                                   "decode all file columns -> materialising filter -> f64 hash aggregate -> merge (Zig ThreadEngine steps, "
                                   "PythonEngine arithmetic); file holds only the 6 Q1 columns, which favours the CPU"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def cuda_arm(args: argparse.Namespace) -> None:
    import numpy as np

    import cases
    from minispark_b200 import CudaExecutionEngine
    from minispark_b200 import native as N

    rank, world, local_rank = dist_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier() -> None:
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    path, gen_s = ensure_table(args.sf, rank)
    # every rank owns a whole table of its own (weak scaling), so the engine must not shard it again
    engine = CudaExecutionEngine(device=local_rank, layout=args.layout, shard=(0, 1))
    ns = cases.namespace()
    try:
        # BlockFile image in pinned host memory: the "host buffers" of the e2e measurement
        nbytes = path.stat().st_size
        pinned = C.c_void_p()
        with NumaLocal(local_rank) as numa:  # the image's pages belong on the NUMA node this GPU hangs off
            engine.ctx.call("msc_host_alloc", nbytes, C.byref(pinned))
            view = (C.c_char * nbytes).from_address(pinned.value)
            with open(path, "rb") as f:
                got = f.readinto(view)
        assert got == nbytes
        engine.register_table_image(str(path), pinned.value, nbytes)

        task = engine.sql(cases.Q1_SQL.format(table=str(path))).task
        launches0 = engine.ctx.stats().launches

        # ---- correctness first: full-precision device result vs the f64 C restatement (1e-9) --------
        rel, schema = engine.execute_to_device(task)
        names = [n for n, _ in schema]
        keys = rel.cols[0].dict.export()
        cols = [rel.column_numpy(i) for i in range(len(names))]
        result = {keys[int(cols[0][r])]: {n: cols[i][r].item() for i, n in enumerate(names) if i} for r in range(rel.nrows)}
        rows_total = int(sum(v["count_order"] for v in result.values()))
        engine.release_query()
        check = "skipped (multi-rank result is the merge of all ranks' tables; each table is checked at N=1)"
        if rank == 0 and world == 1:
            oracle = run_q1_port(path, host_threads(), 0, 0)
            assert oracle["rows"] >= rows_total
            for g in oracle["groups"]:
                mine = result[g["key"]]
                assert mine["count_order"] == g["count"], (g["key"], mine["count_order"], g["count"])
                for a, b in (("sum_qty", "sum_qty"), ("sum_base_price", "sum_base_price"), ("sum_disc_price", "sum_disc_price"),
                             ("sum_charge", "sum_charge")):
                    assert abs(mine[a] - g[b]) <= 1e-9 * abs(g[b]), (g["key"], a, mine[a], g[b])
                assert abs(mine["avg_disc"] - g["sum_disc"] / g["count"]) <= 1e-9 * abs(g["sum_disc"] / g["count"])
            check = "ok: 3 groups, counts exact, f64 sums within 1e-9 of oracle/q1_port.c"
        entry = engine._tables[str(path)]
        nrows_table = entry.nrows

        # ---- prepared hot path: compile once; a step = fused scan-aggregate (+ cross-rank merge) + projection ----
        prepared = engine.prepare(task)
        bytes_per_row = prepared.bytes_per_row
        agg_launch: dict = {}

        def step() -> tuple[float, float]:
            final, dev_ms = prepared.run()
            agg_launch.update(prepared.scan_stats)
            engine.release_query()
            return dev_ms, prepared.scan_stats["scan_ms"]

        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        engine.ctx.call("msc_sync")
        scan_ms_all = []
        with ClockSampler(local_rank) as clocks:
            # the timed region: EXACTLY `steps` passes between two events on the library's stream (msc_timer_*), with a
            # barrier + synchronisation on both sides; everything a pass does -- launches, the host work between them,
            # the NCCL all-gather of the partial tables -- lies between the two events
            t0 = time.perf_counter()
            engine.ctx.call("msc_timer_start")
            for _ in range(args.steps):
                _, s = step()
                scan_ms_all.append(s)
            region_ms = C.c_double()
            engine.ctx.call("msc_timer_stop", C.byref(region_ms))
            if dist is not None:
                import torch

                torch.cuda.synchronize(local_rank)
            wall_s = time.perf_counter() - t0
        barrier()
        launches_per_step = None
        l0 = engine.ctx.stats().launches
        step()
        launches_per_step = engine.ctx.stats().launches - l0

        dev_s = max_over_ranks(region_ms.value / 1e3)
        wall_max = max_over_ranks(wall_s)
        total_rows = sum_over_ranks(float(nrows_table))
        value = total_rows * args.steps / dev_s
        if world > 1:  # Q1's filter keeps every generated row, so the merged COUNT must equal all ranks' rows
            assert rows_total == int(total_rows), (rows_total, total_rows)
            check = f"ok: merged COUNT over {world} ranks == {int(total_rows)} input rows; per-table sums are checked at N=1"
        scan_ms = statistics.mean(scan_ms_all)
        achieved = nrows_table * bytes_per_row / (scan_ms * 1e-3) / 1e9
        peak, peak_src = peaks()
        kernel = kernel_name(agg_launch["rows_per_thread"], agg_launch.get("kind", 0), agg_launch.get("regs", 0))
        traffic, traffic_src = captured_traffic(args.sf, args.layout, nrows_table, bytes_per_row, kernel)

        # ---- e2e: pinned host image -> H2D -> decode -> scan -> result back on the host -----------------
        e2e_times, h2d_bytes, d2h_bytes = [], 0, 0
        for i in range(args.e2e_steps + 1):
            engine.drop_table_cache()
            barrier()
            t0 = time.perf_counter()
            rel, schema = engine.execute_to_device(task)
            host_cols = [rel.column_numpy(c) for c in range(len(schema))]
            dt = time.perf_counter() - t0
            d2h_bytes = sum(a.nbytes for a in host_cols)
            h2d_bytes = engine.last_stats.get("ingest_bytes", 0)
            engine.release_query()
            if i > 0:  # first pass warms allocator pools
                e2e_times.append(dt)
        # median: the host link is shared with other tenants of the box, single passes are occasionally several times slower
        e2e_s = max_over_ranks(statistics.median(e2e_times))
        e2e_value = total_rows / e2e_s

        if rank == 0:
            cpu = None
            try:
                threads = host_threads()
                sample_blocks = 0
                r = run_q1_port(path, threads, sample_blocks, 1)
                cpu = {"value": r["rows"] / r["seconds"], "unit": "rows/s", "cores": threads, "kind": "port",
                       "sample": f"all {r['blocks']} row-blocks ({r['rows']} rows) of rank 0's table once ({r['seconds']:.2f} s), oracle/q1_port.c "
                                 f"on {threads} threads (Zig ThreadEngine steps with PythonEngine f64 arithmetic; 6-column file)"}
            except Exception as e:  # noqa: BLE001
                cpu = {"value": None, "unit": "rows/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
            line = {
                "metric": "tpch_q1_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {
                    "workload": f"TPC-H Q1 (examples/benchmark.py:51-68) on synthetic lineitem sf{args.sf:g} per GPU, sharded by row-block",
                    "sf_per_gpu": args.sf, "rows_per_gpu": nrows_table, "layout": args.layout, "bytes_per_row_scanned": bytes_per_row,
                    "l2": "inputs larger than L2 (scanned columns %.2f GB per GPU vs 126 MB L2)" % (nrows_table * bytes_per_row / 1e9),
                    "timing": "two CUDA events on the library's stream bracketing all timed steps (host gaps and the cross-rank exchange included), max over ranks",
                    "wall_ms_per_step": 1e3 * wall_max / args.steps, "parity_check": check,
                },
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "kernel": kernel, "launch": agg_launch,
                             "program": prepared.prog.program.text,
                             "kernel_ms": scan_ms, "algorithmic_bytes_per_launch": nrows_table * bytes_per_row, "peak_source": peak_src,
                             "kernel_ms_includes": ("the in-kernel wait for the slowest rank's partial table (fused exchange): quote the roofline at N=1"
                                                    if world > 1 and str(agg_launch.get("exchange", "")).startswith("nvlink") else "the scan kernel alone"),
                             "north_star_layout_equiv_gbs": nrows_table * Q1_WIDE_BYTES_PER_ROW / (scan_ms * 1e-3) / 1e9 if args.layout == "native" else None},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                        "ms_per_step": 1e3 * e2e_s, "h2d_gbs": h2d_bytes / e2e_s / 1e9,
                        "passes_ms": [round(1e3 * t, 2) for t in e2e_times], "statistic": "median of the passes (wall clock, rank-local)"},
                "gpu_launches": int(launches_per_step * args.steps),
                "clocks": clocks.summary(),
                "setup": {"generate_s": gen_s, "pinned_image_numa_cpus": (f"{numa.cpus[0]}-{numa.cpus[-1]} ({len(numa.cpus)} CPUs next to the GPU)"
                                                                          if numa.cpus else "default placement")},
            }
            print(json.dumps(line))
        engine.ctx.call("msc_host_free", pinned)
    finally:
        engine.close()
        if dist is not None:
            dist.destroy_process_group()
        if not args.keep and os.environ.get("MSC_BENCH_KEEP") is None:
            try:
                path.unlink()
            except OSError:
                pass


def _validated(task):  # noqa: ANN001, ANN202
    from copy import deepcopy

    t = deepcopy(task)
    t.validate_schema()
    return t


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
