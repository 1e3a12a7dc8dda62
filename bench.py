#!/usr/bin/env python
"""TPC-H Q1 benchmark of the CUDA engine (BASELINE.json: "TPC-H Q1 rows/sec & scan GB/s vs HBM peak").

    python bench.py --gpus 1 --steps 20 --warmup 3 [--sf 15] [--layout native|wide]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm: the reference's native path, restated (oracle/q1_port.c)

A step is one pass of the hot path over the whole (per-rank) lineitem: fused scan + filter +
GROUP BY + cross-rank merge + the AVG projection over the 3 result groups, one kernel per rank.
`value` is measured with the columns resident in HBM (CUDA events on the library's stream, max over
ranks); `e2e` goes through the plugin call a DataFrame makes -- `execute_full_task` from a BlockFile
image in pinned host memory (H2D every step) to the result BlockFile, read back by `collect_results`.
Rank 0 prints ONE JSON line.  Its `extra` object carries the other BASELINE.json configs, each on ONE
table sharded by the engine itself (contiguous row-block ranges per rank) and each checked at full size
against the C restatements of the reference (oracle/q1_port.c, oracle/cfg_port.c):
  extra.q1_sharded  Q1 on one lineitem of --strong-sf (strong scaling over the ranks)          config 3
  extra.highcard    GROUP BY l_orderkey SUM / AVG: pre-aggregate, hash partition, row exchange   config 4
  extra.join        orders JOIN lineitem + BETWEEN + LIKE + GROUP BY, both sides exchanged       config 5
  extra.collect     warm DataFrame.collect() of Q1 through the plugin path (repeat = prepared pass)
and `cpu_baseline_python` is the real reference PythonExecutionEngine (baseline/_ref) on a bounded sample.
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

PIPELINE = 3  # prepared passes in flight (msc_prepared_enqueue / msc_prepared_wait)
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
    ap.add_argument("--strong-sf", type=float, default=100.0, help="scale factor of the ONE lineitem extra.q1_sharded shards over all ranks")
    ap.add_argument("--cfg-sf", type=float, default=10.0, help="scale factor of the tables of extra.highcard / extra.join")
    ap.add_argument("--no-extras", action="store_true", help="only the headline Q1 line")
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
        self.nvml_samples: list[tuple[float, float, int]] = []  # (sm MHz, max sm MHz, event-reason bits) straight from NVML
        self._stop = threading.Event()
        self._nvml_thread: threading.Thread | None = None

    def _nvml_loop(self, nvml, handle) -> None:  # noqa: ANN001
        reasons = getattr(nvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or nvml.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            smax = float(nvml.nvmlDeviceGetMaxClockInfo(handle, nvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            smax = 0.0
        while True:
            try:
                self.nvml_samples.append((float(nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM)), smax, int(reasons(handle))))
            except Exception:  # noqa: BLE001
                return
            if self._stop.wait(0.002):
                return

    def __enter__(self) -> "ClockSampler":
        # NVML in a thread of this process: a sample every 2 ms, so that even the ~8 ms timed region of an 8-GPU run is covered
        # (nvidia-smi -lms needs ~0.2 s before its first line: with eight of them starting at once the region was over first)
        try:
            import pynvml as nvml

            nvml.nvmlInit()
            handle = nvml.nvmlDeviceGetHandleByIndex(self.device)
            self._nvml_thread = threading.Thread(target=self._nvml_loop, args=(nvml, handle), daemon=True)
            self._nvml_thread.start()
            return self
        except Exception:  # noqa: BLE001  (no NVML bindings: the command-line tool instead)
            self._nvml_thread = None
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
        if self._nvml_thread is not None:
            self._stop.set()
            self._nvml_thread.join(timeout=1.0)
            return
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
        if self.nvml_samples:
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}  # nvmlClocksEventReason*
            for clock, top, mask in self.nvml_samples:
                sm.append(clock)
                smax.append(top)
                reasons.update(name for name, bit in bits.items() if mask & bit)
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm), "source": "nvml, every 2 ms"}
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
    for name in ("r02_traffic.json", "r01_traffic.json"):  # the newest capture of this workload and kernel
        path = ROOT / "profiles" / name
        if not path.exists():
            continue
        for cap in json.loads(path.read_text()).get("captures", []):
            w = cap["workload"]
            same_kernel = kernel.startswith(cap["kernel"].split("::")[-1])
            if same_kernel and (w.get("sf_per_gpu"), w.get("layout"), w.get("rows_per_gpu"), w.get("bytes_per_row_scanned")) == (sf, layout, rows, bytes_per_row):
                return float(cap["traffic_bytes"]), f"profiles/{name} ({cap['capture']})"
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


def gen_workers() -> int:
    """Processes the table generator may use on this rank: the host's threads shared between the local ranks."""
    return max(1, min(16, host_threads() // max(int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))), 1)))


def ensure_table(sf: float, rank: int) -> tuple[Path, float]:
    import gen_tpch

    path = table_path(sf, rank)
    t0 = time.perf_counter()
    if not path.exists():
        tmp = path.with_suffix(".tmp")
        gen_tpch.write_table(tmp, "lineitem", sf=sf, columns=gen_tpch.Q1_COLUMNS, seed=1234 + 1000 * rank, workers=gen_workers())
        tmp.rename(path)
    return path, time.perf_counter() - t0


def ensure_shared_table(name: str, rank: int, barrier, **gen) -> Path:  # noqa: ANN001, ANN003
    """ONE table for all ranks (they shard it by row-block themselves): rank 0 writes it, everybody waits."""
    import gen_tpch

    path = table_path(0, 0).with_name(name)
    if rank == 0 and not path.exists():
        tmp = path.with_suffix(".tmp")
        gen_tpch.write_table(tmp, workers=host_threads(), **gen)
        tmp.rename(path)
    barrier()
    return path


def run_q1_port(path: Path, threads: int, max_blocks: int, wire: int) -> dict:
    from oracle import ports  # the checker / CPU baseline: the one place bench.py executes oracle/

    return ports.q1(path, threads, max_blocks, wire)


def check_q1(result: dict, oracle: dict) -> None:
    """Full-precision engine result {flag: {column: value}} against q1_port's groups: counts exact, f64 sums / AVG 1e-9."""
    assert sorted(result) == sorted(g["key"] for g in oracle["groups"]), (sorted(result), oracle["groups"])
    for g in oracle["groups"]:
        mine = result[g["key"]]
        assert mine["count_order"] == g["count"], (g["key"], mine["count_order"], g["count"])
        for a in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge"):
            assert abs(mine[a] - g[a]) <= 1e-9 * abs(g[a]), (g["key"], a, mine[a], g[a])
        for a, b in (("avg_qty", "sum_qty"), ("avg_price", "sum_base_price"), ("avg_disc", "sum_disc")):
            assert abs(mine[a] - g[b] / g["count"]) <= 1e-9 * abs(g[b] / g["count"]), (g["key"], a)


def fold_q1_oracles(per_rank: list[dict]) -> dict:
    groups: dict[str, dict] = {}
    for r in per_rank:
        for g in r["groups"]:
            acc = groups.setdefault(g["key"], {k: (v if k == "key" else 0) for k, v in g.items()})
            for k, v in g.items():
                if k != "key":
                    acc[k] += v
    return {"rows": sum(r["rows"] for r in per_rank), "groups": list(groups.values())}


def q1_result(rel, schema) -> dict:  # noqa: ANN001
    names = [n for n, _ in schema]
    keys = rel.cols[0].dict.export()
    cols = [rel.column_numpy(i) for i in range(len(names))]
    return {keys[int(cols[0][r])]: {n: cols[i][r].item() for i, n in enumerate(names) if i} for r in range(rel.nrows)}


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

    def gather_objects(obj):  # noqa: ANN001, ANN202
        if dist is None:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

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
        # every rank checks the MERGED result: its own table through oracle/q1_port.c, the ranks' oracle groups folded in
        # rank order (what the final aggregate over the shuffled partials computes, plan.py:190-199)
        rel, schema = engine.execute_to_device(task)
        result = q1_result(rel, schema)
        rows_total = int(sum(v["count_order"] for v in result.values()))
        engine.release_query()
        mine = run_q1_port(path, max(1, host_threads() // world), 0, 0)
        oracle = fold_q1_oracles(gather_objects(mine))
        assert oracle["rows"] >= rows_total
        check_q1(result, oracle)
        check = (f"ok: one-shot and prepared (timed) results, 3 groups merged over {world} rank(s): counts exact, every f64 SUM / AVG within 1e-9 "
                 "of oracle/q1_port.c run on each rank's table and folded in rank order")
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
            # Passes are enqueued up to PIPELINE deep and collected in order: each is a complete query (one kernel: scan,
            # cross-rank exchange, merge, final projection; its result's row count is read back and checked), but the host
            # does not sit between two kernels -- with several ranks the in-kernel exchange then waits for the slowest GPU,
            # not for the slowest host loop.
            in_flight = 0
            for _ in range(args.steps):
                if prepared.enqueue():
                    in_flight += 1
                    if in_flight == PIPELINE:
                        assert prepared.wait().nrows == 3
                        scan_ms_all.append(engine.ctx.stats().last_scan_ms)  # this pass's kernel, from its own event pair
                        in_flight -= 1
                else:
                    _, s = step()
                    scan_ms_all.append(s)
            while in_flight:
                assert prepared.wait().nrows == 3
                scan_ms_all.append(engine.ctx.stats().last_scan_ms)
                in_flight -= 1
            region_ms = C.c_double()
            engine.ctx.call("msc_timer_stop", C.byref(region_ms))
            if dist is not None:
                import torch

                torch.cuda.synchronize(local_rank)
            wall_s = time.perf_counter() - t0
        barrier()
        launches_per_step = None
        l0 = engine.ctx.stats().launches
        final, _ = prepared.run()
        launches_per_step = engine.ctx.stats().launches - l0
        check_q1(q1_result(final, prepared.plan.schema), oracle)  # the timed path (prepared pass, in-kernel exchange at N > 1)
        engine.release_query()

        dev_s = max_over_ranks(region_ms.value / 1e3)
        wall_max = max_over_ranks(wall_s)
        total_rows = sum_over_ranks(float(nrows_table))
        value = total_rows * args.steps / dev_s
        assert rows_total == int(total_rows), (rows_total, total_rows)  # Q1's filter keeps every generated row
        scan_ms = statistics.mean(scan_ms_all)
        achieved = nrows_table * bytes_per_row / (scan_ms * 1e-3) / 1e9
        peak, peak_src = peaks()
        kernel = kernel_name(agg_launch["rows_per_thread"], agg_launch.get("kind", 0), agg_launch.get("regs", 0))
        traffic, traffic_src = captured_traffic(args.sf, args.layout, nrows_table, bytes_per_row, kernel)

        # ---- e2e: the call a DataFrame makes.  Pinned host BlockFile image -> execute_full_task (H2D of the referenced
        # columns, decode, scan, result BlockFile) -> collect_results (the rows back as Python dicts) --------------------
        e2e_times, h2d_bytes, d2h_bytes = [], 0, 0
        for i in range(args.e2e_steps + 1):
            engine.drop_table_cache()
            barrier()
            t0 = time.perf_counter()
            jobs = engine.execute_full_task(task)
            rows_back = list(engine.collect_results(jobs))
            dt = time.perf_counter() - t0
            d2h_bytes = sum(f.file_path.stat().st_size for j in jobs for f in j.output_files)
            h2d_bytes = engine.last_stats.get("ingest_bytes", 0)
            assert len(rows_back) == 3 and sum(r["count_order"] for r in rows_back) == int(total_rows)
            if i > 0:  # first pass warms allocator pools
                e2e_times.append(dt)
        # median: the host link is shared with other tenants of the box, single passes are occasionally several times slower
        e2e_s = max_over_ranks(statistics.median(e2e_times))
        e2e_value = total_rows / e2e_s

        # ---- what the host can deliver: every rank copies its pinned image to the device at the same time, nothing else running
        # (the ceiling of the e2e path's H2D leg on this box; with 8 ranks the host side, not the links, sets it)
        raw_gbs = None
        try:
            scratch = C.c_void_p()
            engine.ctx.call("msc_dev_alloc", nbytes, C.byref(scratch))
            best = None
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                engine.ctx.call("msc_memcpy_h2d", scratch, pinned, nbytes)
                dt = max_over_ranks(time.perf_counter() - t0)
                best = dt if best is None else min(best, dt)
            engine.ctx.call("msc_dev_free", scratch)
            raw_gbs = nbytes / best / 1e9
        except Exception:  # noqa: BLE001
            raw_gbs = None

        # ---- the plugin path warm: DataFrame.collect() of the same SQL, columns resident; the second identical task tree
        # is prepared by the engine itself (plan cache), later ones are one launch + result BlockFile + collect_results
        collect_times, collect_plan = [], None
        for i in range(6):
            barrier()
            t0 = time.perf_counter()
            rows_back = engine.sql(cases.Q1_SQL.format(table=str(path))).collect()
            collect_times.append(time.perf_counter() - t0)
            collect_plan = engine.last_stats.get("plan")
        assert len(rows_back) == 3
        collect_ms = 1e3 * max_over_ranks(statistics.median(collect_times[2:]))

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
                    "timing": "two CUDA events on the library's stream bracketing all timed steps (host gaps and the cross-rank exchange included), max over ranks; "
                              f"passes are enqueued up to {PIPELINE} deep (msc_prepared_enqueue / msc_prepared_wait) and every result's row count is read back",
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
                        "h2d_raw_gbs_per_gpu": raw_gbs, "h2d_frac_of_raw": (h2d_bytes / e2e_s / 1e9) / raw_gbs if raw_gbs else None,
                        "h2d_raw_how": f"one cudaMemcpyAsync of the whole pinned image ({nbytes} bytes) per rank, all {world} rank(s) at the same time, best of 3",
                        "passes_ms": [round(1e3 * t, 2) for t in e2e_times], "statistic": "median of the passes (wall clock, rank-local)",
                        "path": "engine.execute_full_task(task) on a pinned host BlockFile image (columns dropped from the device before every pass) "
                                "-> result BlockFile -> engine.collect_results(...) rows"},
                "gpu_launches": int(launches_per_step * args.steps),
                "clocks": clocks.summary(),
                "setup": {"generate_s": gen_s, "pinned_image_numa_cpus": (f"{numa.cpus[0]}-{numa.cpus[-1]} ({len(numa.cpus)} CPUs next to the GPU)"
                                                                          if numa.cpus else "default placement")},
                "extra": {"collect": {"ms": collect_ms, "plan": collect_plan, "passes_ms": [round(1e3 * t, 3) for t in collect_times],
                                      "what": "warm engine.sql(Q1).collect(): parse, plan cache hit, ONE kernel, result BlockFile, collect_results (columns resident)",
                                      "rows_per_s": total_rows / (collect_ms * 1e-3)}},
            }
        engine.ctx.call("msc_host_free", pinned)
        engine.close()
        engine = None
        # ---- the other BASELINE configs, each on ONE table the engine shards itself -------------------------------------
        if not args.no_extras:
            tools = {"barrier": barrier, "max_over_ranks": max_over_ranks, "sum_over_ranks": sum_over_ranks, "gather_objects": gather_objects}
            extras = run_extras(args, rank, world, local_rank, tools)
            if rank == 0:
                line["extra"].update(extras)
                if world == 1:
                    line["cpu_baseline_python"] = python_engine_baseline()
        if rank == 0:
            print(json.dumps(line))
    finally:
        if engine is not None:
            engine.close()
        if dist is not None:
            dist.destroy_process_group()
        if args.keep is False and os.environ.get("MSC_BENCH_DROP") is not None:  # tables stay in /dev/shm for the next N of a scaling run
            try:
                path.unlink()
            except OSError:
                pass


# ----------------------------------------------------------------------------------------------------
def python_engine_baseline() -> dict:
    """The REAL reference PythonExecutionEngine (baseline/_ref, unmodified) on a bounded sample of the Q1 workload: it is
    single-threaded by construction (execution.py:69-83) and runs ~1e5 rows/s, so the sample is sf0.05 (~0.3 M rows);
    larger scale factors would be linear extrapolations and are not quoted."""
    import gen_tpch

    sample = table_path(0, 0).with_name("lineitem_q1_sf0.05_pyengine.bin")
    try:
        if not sample.exists():
            gen_tpch.write_table(sample, "lineitem", sf=0.05, columns=gen_tpch.Q1_COLUMNS)
        out = subprocess.run([sys.executable, str(ROOT / "bench" / "ref_python_engine.py"), str(sample)], capture_output=True, text=True, timeout=300)
        r = json.loads(out.stdout.strip().splitlines()[-1])
        if "unavailable" in r:
            return {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": r["unavailable"]}
        return {"value": r["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "reference",
                "sample": f"TPC-H Q1 through DataFrame.collect() on the unmodified reference PythonExecutionEngine (baseline/_ref, Python {r['python']}), "
                          f"lineitem sf0.05 = {r['rows']} rows of the 6 Q1 columns in {r['seconds']:.2f} s; single-threaded by construction"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": f"failed: {e!r}"[:300]}


def run_extras(args: argparse.Namespace, rank: int, world: int, local_rank: int, tools: dict) -> dict:
    """extra.q1_sharded / extra.highcard / extra.midcard / extra.join: every rank runs them (collective); rank 0's dict is printed."""
    import gen_tpch
    from minispark_b200 import CudaExecutionEngine

    out: dict = {}
    # (no shard argument: the engine takes (rank, world) from the process group and shards every table it opens)
    engine = CudaExecutionEngine(device=local_rank, layout=args.layout)
    try:
        assert engine.shard == (rank, world)
        for name, fn in (("q1_sharded", extra_q1_sharded), ("highcard", extra_highcard), ("midcard", extra_midcard), ("join", extra_join)):
            t0 = time.perf_counter()
            try:
                res = fn(args, engine, rank, world, tools)
                ok = True
            except Exception as e:  # noqa: BLE001
                res, ok = {"error": f"{type(e).__name__}: {e}"[:400]}, False
                try:
                    engine.release_query()
                except Exception:  # noqa: BLE001
                    pass
            res["bench_wall_s"] = round(time.perf_counter() - t0, 2)
            out[name] = res
            if not all(tools["gather_objects"](ok)):  # a rank failed: the others must not walk into its collectives
                break
            engine.drop_table_cache()
    finally:
        engine.close()
    return out


def _column_bytes(engine, table: Path, names: list[str]) -> int:  # noqa: ANN001
    """Bytes per row of the device-resident columns `names` of a table, at the widths actually allocated."""
    from minispark_b200 import native as N

    entry = engine._tables[str(table)]
    index = {n: i for i, (n, _) in enumerate(entry.schema)}
    return sum(N.PHYS_WIDTH[entry.columns[index[n]].phys] for n in names)


def _time_one_shot(engine, task, reps: int, tools: dict) -> tuple[float, list[float], dict, int]:  # noqa: ANN001
    """Median over `reps` executions (after one untimed) of the max-over-ranks wall time of execute_to_device, the device idle
    on both sides.  -> (seconds, per-pass seconds, last_stats of the last pass, local result rows)"""
    times, stats, nrows = [], {}, 0
    for i in range(reps + 1):
        tools["barrier"]()
        engine.ctx.call("msc_sync")
        t0 = time.perf_counter()
        rel, _ = engine.execute_to_device(task)
        engine.ctx.call("msc_sync")
        dt = time.perf_counter() - t0
        stats, nrows = dict(engine.last_stats), rel.nrows
        engine.release_query()
        if i:
            times.append(tools["max_over_ranks"](dt))
    return statistics.median(times), times, stats, nrows


def extra_q1_sharded(args, engine, rank: int, world: int, tools: dict) -> dict:  # noqa: ANN001
    """BASELINE config 3: Q1 on ONE lineitem, row-blocks dealt to the ranks by the engine (strong scaling)."""
    import cases
    import gen_tpch

    sf = args.strong_sf
    free = os.statvfs(table_path(0, 0).parent).f_bavail * os.statvfs(table_path(0, 0).parent).f_frsize
    need = 6.0e6 * sf * 26 * 1.2
    while sf > 5 and need > free * 0.5:  # a small /dev/shm: shrink rather than fail (recorded below)
        sf, need = sf / 2, need / 2
    path = ensure_shared_table(f"lineitem_q1_sf{sf:g}_shared.bin", rank, tools["barrier"], table="lineitem", sf=sf, columns=gen_tpch.Q1_COLUMNS, seed=4321)
    task = engine.sql(cases.Q1_SQL.format(table=str(path))).task
    t0 = time.perf_counter()
    prepared = engine.prepare(task)
    ingest_s = time.perf_counter() - t0
    steps = max(min(args.steps, 20), 3)
    for _ in range(3):
        final, _ = prepared.run()
        engine.release_query()
    tools["barrier"]()
    engine.ctx.call("msc_sync")
    engine.ctx.call("msc_timer_start")
    scan_ms = []
    for _ in range(steps):
        final, _ = prepared.run()
        scan_ms.append(prepared.scan_stats["scan_ms"])
        engine.release_query()
    region = C.c_double()
    engine.ctx.call("msc_timer_stop", C.byref(region))
    dev_s = tools["max_over_ranks"](region.value / 1e3)
    final, _ = prepared.run()
    result = q1_result(final, prepared.plan.schema)
    engine.release_query()
    local_rows = prepared.nrows
    total_rows = int(tools["sum_over_ranks"](float(local_rows)))
    parity = "not checked on this rank"
    if rank == 0:
        oracle = run_q1_port(path, host_threads(), 0, 0)
        assert oracle["rows"] == total_rows, (oracle["rows"], total_rows)
        check_q1(result, oracle)
        parity = "ok: merged result (every rank holds it) vs oracle/q1_port.c over the whole file: counts exact, f64 SUM / AVG within 1e-9"
    peak, _ = peaks()
    ms = 1e3 * dev_s / steps
    gbs = total_rows * prepared.bytes_per_row / (ms * 1e-3) / 1e9
    return {"workload": f"TPC-H Q1 on ONE synthetic lineitem sf{sf:g} ({total_rows} rows), contiguous row-block ranges per rank (engine sharding)",
            "scaling": "strong", "sf": sf, "rows": total_rows, "rows_this_rank": local_rows, "ms_per_step": ms, "rows_per_s": total_rows / (ms * 1e-3),
            "kernel_ms_this_rank": statistics.mean(scan_ms), "scanned_gbs_all_gpus": gbs, "frac_of_peak_all_gpus": gbs / (peak * world),
            "exchange": prepared.scan_stats.get("exchange", "none (one rank)" if world == 1 else "partial-table all-gather"),
            "ingest_and_prepare_s": round(ingest_s, 2), "steps": steps, "parity_check": parity}


def extra_highcard(args, engine, rank: int, world: int, tools: dict) -> dict:  # noqa: ANN001
    """BASELINE config 4: SELECT l_orderkey, SUM(l_quantity), AVG(l_extendedprice) FROM lineitem GROUP BY l_orderkey."""
    import numpy as np

    import cases
    from oracle import ports

    ns = cases.namespace()
    sf = args.cfg_sf
    lineitem = ensure_shared_table(f"lineitem_cfg_sf{sf:g}.bin", rank, tools["barrier"], table="lineitem", sf=sf,
                                   columns=["l_orderkey", "l_quantity", "l_extendedprice", "l_shipmode"])
    task = ns.DataFrame(engine).table(str(lineitem)).group_by(ns.Col("l_orderkey")).agg(
        ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.avg(ns.Col("l_extendedprice")).alias("p")).task
    reps = max(min(args.steps, 7), 3)
    sec, passes, stats, local_groups = _time_one_shot(engine, task, reps, tools)
    local_rows = engine._tables[str(lineitem)].nrows
    total_rows = int(tools["sum_over_ranks"](float(local_rows)))
    bytes_per_row = _column_bytes(engine, lineitem, ["l_orderkey", "l_quantity", "l_extendedprice"])
    sent = tools["sum_over_ranks"](float(stats.get("exchange_bytes_sent", 0) if world > 1 else 0))
    exch_s = tools["max_over_ranks"](float(stats.get("exchange_host_s", 0.0) if world > 1 else 0.0))
    # parity at full size: the complete result on every rank (rank-ordered gather of the partitions), rank 0 checks it
    rel, _ = engine.execute_to_device(task, replicate=True)
    keys, q, p = (rel.column_numpy(i) for i in range(3))
    engine.release_query()
    parity = "not checked on this rank"
    if rank == 0:
        want = ports.highcard(lineitem, table_path(0, 0).with_name("highcard_oracle.bin"))
        order = np.argsort(keys, kind="stable")
        assert want["rows"] == total_rows and len(keys) == want["groups"], (want["rows"], total_rows, len(keys), want["groups"])
        assert np.array_equal(keys[order], want["keys"]), "group keys differ"
        np.testing.assert_allclose(q[order], want["sum_q"], rtol=1e-9, atol=0)
        np.testing.assert_allclose(p[order], want["sum_p"] / want["count"], rtol=1e-9, atol=0)
        parity = f"ok: {len(keys)} groups gathered from {world} rank(s): keys exact, SUM and AVG within 1e-9 of oracle/cfg_port.c (f64) over the whole file"
    peak, _ = peaks()
    gbs = total_rows * bytes_per_row / sec / 1e9
    return {"workload": f"GROUP BY l_orderkey SUM(l_quantity), AVG(l_extendedprice) on ONE lineitem sf{sf:g} ({total_rows} rows), engine sharding",
            "sf": sf, "rows": total_rows, "groups": int(tools["sum_over_ranks"](float(local_groups))) if stats.get("result_partitioned") else local_groups,
            "ms": 1e3 * sec, "passes_ms": [round(1e3 * t, 3) for t in passes], "rows_per_s": total_rows / sec, "bytes_per_row_scanned": bytes_per_row,
            "algorithmic_bytes": total_rows * bytes_per_row, "scanned_gbs_all_gpus": gbs, "frac_of_peak_all_gpus": gbs / (peak * world),
            "agg_mode": stats.get("agg_mode"), "agg_scan_ms_this_rank": stats.get("agg_scan_ms"), "exchange": stats.get("exchange") or "none (one rank)",
            "exchange_bytes_sent_all_ranks": int(sent), "exchange_host_ms": 1e3 * exch_s,
            "exchange_gbs": (sent / exch_s / 1e9) if exch_s > 0 else None, "result_partitioned": bool(stats.get("result_partitioned")),
            "timing": "wall clock around execute_to_device (one-shot: lowering + all launches + host waits), device idle on both sides, max over ranks, median of the passes",
            "parity_check": parity}


def extra_midcard(args, engine, rank: int, world: int, tools: dict) -> dict:  # noqa: ANN001
    """GROUP BY with few groups and seven aggregates (six accumulators) on the config-4 lineitem: by l_shipmode (7 groups, a
    dictionary key: dense table, 49 cells -> the specialised kernel's shared-memory cell form) and by l_quantity (50 groups, a
    FLOAT key: hash aggregate behind CTA-local tables)."""
    import numpy as np

    import cases
    from oracle import ports

    ns = cases.namespace()
    sf = args.cfg_sf
    lineitem = ensure_shared_table(f"lineitem_cfg_sf{sf:g}.bin", rank, tools["barrier"], table="lineitem", sf=sf,
                                   columns=["l_orderkey", "l_quantity", "l_extendedprice", "l_shipmode"])
    peak, _ = peaks()
    out: dict = {"sf": sf}
    for key in ("l_shipmode", "l_quantity"):
        task = cases.midcard_frame(ns, lineitem, key, engine).task
        reps = max(min(args.steps, 9), 5)
        sec, passes, stats, _ = _time_one_shot(engine, task, reps, tools)
        local_rows = engine._tables[str(lineitem)].nrows
        total_rows = int(tools["sum_over_ranks"](float(local_rows)))
        bytes_per_row = _column_bytes(engine, lineitem, [key, "l_quantity", "l_extendedprice"] if key != "l_quantity" else ["l_quantity", "l_extendedprice"])
        # (a one-shot hash aggregate is followed by its final projection, whose launch overwrites kernel_ms: take the aggregate scan's own time)
        kernel_ms = tools["max_over_ranks"](float((stats.get("agg_scan_ms") if stats.get("agg_mode") == "hash" else stats.get("kernel_ms")) or 0.0))
        rel, schema = engine.execute_to_device(task, replicate=True)
        names = [n for n, _ in schema]
        cols = {n: rel.column_numpy(i) for i, n in enumerate(names)}
        keys = [rel.cols[0].dict.export()[c] for c in cols[key].tolist()] if key == "l_shipmode" else [float(v) for v in cols[key].tolist()]
        engine.release_query()
        parity = "not checked on this rank"
        if rank == 0:
            want = ports.groupby(lineitem, key)
            ref = {g["key"]: g for g in want["groups"]}
            assert want["rows"] == total_rows and sorted(keys) == sorted(ref), (want["rows"], total_rows, len(keys), len(ref))
            for i, k in enumerate(keys):
                g = ref[k]
                assert int(cols["n"][i]) == g["count"] and float(cols["min_p"][i]) == g["min_p"] and float(cols["max_p"][i]) == g["max_p"], (key, k)
                for name in ("sum_q", "sum_p", "sum_pq"):
                    assert abs(float(cols[name][i]) - g[name]) <= 1e-9 * abs(g[name]), (key, k, name, float(cols[name][i]), g[name])
                assert abs(float(cols["avg_p"][i]) - g["sum_p"] / g["count"]) <= 1e-9 * abs(g["sum_p"] / g["count"]), (key, k)
            parity = f"ok: {len(keys)} groups vs oracle/cfg_port.c (f64) over the whole file: COUNT / MIN / MAX exact, SUM and AVG within 1e-9"
        gbs = total_rows * bytes_per_row / sec / 1e9
        out[key] = {"workload": f"GROUP BY {key}: COUNT, SUM x3, MIN, MAX, AVG on ONE lineitem sf{sf:g} ({total_rows} rows), engine sharding",
                    "rows": total_rows, "groups": len(keys), "ms": 1e3 * sec, "passes_ms": [round(1e3 * t, 3) for t in passes], "rows_per_s": total_rows / sec,
                    "bytes_per_row_scanned": bytes_per_row, "algorithmic_bytes": total_rows * bytes_per_row,
                    "scanned_gbs_all_gpus": gbs, "frac_of_peak_all_gpus": gbs / (peak * world),
                    "kernel_ms": kernel_ms, "kernel_frac_of_peak_all_gpus": (total_rows * bytes_per_row / (kernel_ms * 1e-3) / 1e9 / (peak * world)) if kernel_ms else None,
                    "agg_mode": stats.get("agg_mode"), "scan_kind": stats.get("scan_kind"), "scan_smem": stats.get("scan_smem"),
                    "hash_local_slots": stats.get("hash_local_slots") if stats.get("agg_mode") == "hash" else None,
                    "plan": stats.get("plan"), "parity_check": parity}
    out["timing"] = ("ms: wall clock around execute_to_device, device idle on both sides, max over ranks, median of the passes; kernel_ms: CUDA events around "
                     "the aggregate scan of the last pass (msc_stats.last_kernel_ms / last_scan_ms), max over ranks")
    return out


def extra_join(args, engine, rank: int, world: int, tools: dict) -> dict:  # noqa: ANN001
    """BASELINE config 5: orders JOIN lineitem ON orderkey WHERE o_orderdate BETWEEN .. AND l_shipmode LIKE '%AIR%' GROUP BY o_orderpriority."""
    from datetime import datetime

    import cases
    from oracle import ports

    ns = cases.namespace()
    sf = args.cfg_sf
    lineitem = ensure_shared_table(f"lineitem_cfg_sf{sf:g}.bin", rank, tools["barrier"], table="lineitem", sf=sf,
                                   columns=["l_orderkey", "l_quantity", "l_extendedprice", "l_shipmode"])
    orders = ensure_shared_table(f"orders_cfg_sf{sf:g}.bin", rank, tools["barrier"], table="orders", sf=sf,
                                 columns=["o_orderkey", "o_orderdate", "o_orderpriority"])
    lo, hi, needle = "1994-01-01", "1994-12-31", "AIR"
    o = ns.DataFrame(engine).table(str(orders)).alias("o")
    l = ns.DataFrame().table(str(lineitem)).alias("l")
    task = (o.join(l, on=ns.Col("o.o_orderkey") == ns.Col("l.l_orderkey"), how="inner")
            .filter(ns.Col("o.o_orderdate").between(lo, hi)).filter(ns.Col("l.l_shipmode").like(f"%{needle}%"))
            .group_by(ns.Col("o.o_orderpriority")).agg(ns.F.count().alias("n"), ns.F.sum(ns.Col("l.l_extendedprice")).alias("rev"))).task
    reps = max(min(args.steps, 7), 3)
    sec, passes, stats, _ = _time_one_shot(engine, task, reps, tools)
    rows_l = int(tools["sum_over_ranks"](float(engine._tables[str(lineitem)].nrows)))
    rows_o = int(tools["sum_over_ranks"](float(engine._tables[str(orders)].nrows)))
    bytes_l = _column_bytes(engine, lineitem, ["l_orderkey", "l_extendedprice", "l_shipmode"])
    bytes_o = _column_bytes(engine, orders, ["o_orderkey", "o_orderdate", "o_orderpriority"])
    algo = rows_l * bytes_l + rows_o * bytes_o
    rel, schema = engine.execute_to_device(task, replicate=True)
    keys = rel.cols[0].dict.export()
    cols = [rel.column_numpy(i) for i in range(3)]
    got = {keys[int(cols[0][r])]: (int(cols[1][r]), float(cols[2][r])) for r in range(rel.nrows)}
    engine.release_query()
    parity = "not checked on this rank"
    if rank == 0:
        us = lambda text: int(datetime.fromisoformat(text).timestamp() * 1_000_000)  # noqa: E731  (TZ=UTC)
        want = ports.join(orders, lineitem, us(lo), us(hi), needle)
        assert {g["key"] for g in want["groups"]} == set(got), (want["groups"], got)
        for g in want["groups"]:
            n, rev = got[g["key"]]
            assert n == g["count"], (g["key"], n, g["count"])
            assert abs(rev - g["sum"]) <= 1e-9 * abs(g["sum"]), (g["key"], rev, g["sum"])
        parity = (f"ok: {len(got)} groups merged over {world} rank(s): COUNT exact (join cardinality {sum(g['count'] for g in want['groups'])} after the filters), "
                  "SUM within 1e-9 of oracle/cfg_port.c (f64) over the whole files")
    peak, _ = peaks()
    gbs = algo / sec / 1e9
    return {"workload": f"orders JOIN lineitem ON orderkey WHERE o_orderdate BETWEEN '{lo}' AND '{hi}' AND l_shipmode LIKE '%{needle}%' GROUP BY o_orderpriority, "
                        f"ONE orders ({rows_o} rows) and ONE lineitem ({rows_l} rows) sf{sf:g}, engine sharding",
            "sf": sf, "lineitem_rows": rows_l, "orders_rows": rows_o, "ms": 1e3 * sec, "passes_ms": [round(1e3 * t, 3) for t in passes],
            "lineitem_rows_per_s": rows_l / sec, "algorithmic_bytes": algo, "referenced_gbs_all_gpus": gbs, "frac_of_peak_all_gpus": gbs / (peak * world),
            "exchange": stats.get("exchanges") or "none (one rank)",
            "exchange_bytes_sent_all_ranks": int(tools["sum_over_ranks"](float(stats.get("exchange_bytes_sent", 0) if world > 1 else 0))),
            "timing": "wall clock around execute_to_device (one-shot: lowering + all launches + host waits), device idle on both sides, max over ranks, median of the passes",
            "parity_check": parity}


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
