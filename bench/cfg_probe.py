#!/usr/bin/env python
"""Where a one-shot query's wall time goes (BASELINE configs 4 / 5 on one GPU): every library call with its host duration, and
the Python time between the calls.

    python bench/cfg_probe.py --sf 10 --config 4|5 [--jit always]
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "bench"):
    sys.path.insert(0, str(p))
os.environ["TZ"] = "UTC"
time.tzset()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--jit", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--profile", type=int, default=0, help="cProfile this many executions instead of the per-call timeline")
    args = ap.parse_args()
    import cases
    import gen_tpch
    from minispark_b200 import CudaExecutionEngine
    from minispark_b200 import native as N

    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    folder = base / f"minispark_b200_cfg_{os.getuid()}"
    folder.mkdir(parents=True, exist_ok=True)
    lineitem, orders = folder / f"lineitem6_sf{args.sf:g}.bin", folder / f"orders_sf{args.sf:g}.bin"
    if int(os.environ.get("RANK", "0")) != 0:
        time.sleep(1.0)
        while not (lineitem.exists() and orders.exists() and (folder / "ready").exists()):
            time.sleep(0.2)
    if not lineitem.exists():
        gen_tpch.write_table(lineitem, "lineitem", sf=args.sf, columns=["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipmode"], workers=16)
    if not orders.exists():
        gen_tpch.write_table(orders, "orders", sf=args.sf, columns=["o_orderkey", "o_orderdate", "o_orderpriority"], workers=16)
    (folder / "ready").touch()
    ns = cases.namespace()

    def build(e):  # noqa: ANN001, ANN202
        if args.config == 4:
            return ns.DataFrame(e).table(str(lineitem)).group_by(ns.Col("l_orderkey")).agg(
                ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.avg(ns.Col("l_extendedprice")).alias("p"))
        if args.config in (6, 7, 8, 9):  # dense / mid-cardinality GROUP BY: 7 groups (dictionary key), 50 (numeric key), 1000, 4000
            key = {6: ns.Col("l_shipmode"), 7: ns.Col("l_quantity"), 8: (ns.Col("l_orderkey") % 1000).alias("k"),
                   9: (ns.Col("l_orderkey") % 4000).alias("k")}[args.config]
            return ns.DataFrame(e).table(str(lineitem)).group_by(key).agg(
                ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.sum(ns.Col("l_extendedprice")).alias("p"),
                ns.F.sum(ns.Col("l_extendedprice") * (ns.Lit(1) - ns.Col("l_discount"))).alias("dp"), ns.F.avg(ns.Col("l_quantity")).alias("aq"),
                ns.F.avg(ns.Col("l_extendedprice")).alias("ap"), ns.F.avg(ns.Col("l_discount")).alias("ad"), ns.F.count())
        o = ns.DataFrame(e).table(str(orders)).alias("o")
        l = ns.DataFrame().table(str(lineitem)).alias("l")
        return (o.join(l, on=ns.Col("o.o_orderkey") == ns.Col("l.l_orderkey"), how="inner")
                .filter(ns.Col("o.o_orderdate").between("1994-01-01", "1994-12-31")).filter(ns.Col("l.l_shipmode").like("%AIR%"))
                .group_by(ns.Col("o.o_orderpriority")).agg(ns.F.count(), ns.F.sum(ns.Col("l.l_extendedprice")).alias("rev")))

    log: list[tuple[str, float, float]] = []
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:  # under torchrun: ONE table, sharded by the engine; rank 0 prints
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    with CudaExecutionEngine(device=local, shard=None if world > 1 else (0, 1), jit=args.jit) as e:
        task = build(e).task
        for _ in range(3):
            e.execute_to_device(task)
            e.release_query()
        if args.profile:
            import cProfile
            import pstats

            prof = cProfile.Profile()
            prof.enable()
            for _ in range(args.profile):
                e.execute_to_device(task)
                e.release_query()
            prof.disable()
            if rank == 0:
                st = pstats.Stats(prof)
                st.sort_stats("tottime").print_stats(28)
                st.sort_stats("cumulative").print_stats(40)
            return
        orig_call, orig_check = N.Context.call, N.Context.check
        t_base = [0.0]

        def call(self, name, *a):  # noqa: ANN001, ANN002, ANN202
            t0 = time.perf_counter()
            try:
                return orig_call(self, name, *a)
            finally:
                log.append((name, t0 - t_base[0], time.perf_counter() - t0))

        N.Context.call = call
        lib = e.ctx.lib
        for fname in ("msc_rel_info", "msc_rel_cols", "msc_dict_size", "msc_rel_free", "msc_shuffle_begin", "msc_shuffle_finish", "msc_shuffle_allgather"):
            orig = getattr(lib, fname)

            def wrap(*a, _orig=orig, _n=fname):  # noqa: ANN002, ANN202
                t0 = time.perf_counter()
                try:
                    return _orig(*a)
                finally:
                    log.append((_n + "*", t0 - t_base[0], time.perf_counter() - t0))

            # (ctypes function objects are looked up by attribute on the CDLL: shadow them on the instance)
            lib.__dict__[fname] = wrap
        for rep in range(args.reps):
            log.clear()
            e.ctx.call("msc_sync")
            log.clear()
            t_base[0] = time.perf_counter()
            rel, _ = e.execute_to_device(task)
            e.ctx.call("msc_sync")
            total = time.perf_counter() - t_base[0]
            e.release_query()
            in_calls = sum(d for _, _, d in log)
            if rep == args.reps - 1 and rank == 0:
                print(f"config {args.config} sf{args.sf:g} jit={e.jit}: wall {1e3 * total:.3f} ms, inside library calls {1e3 * in_calls:.3f} ms, "
                      f"python {1e3 * (total - in_calls):.3f} ms, {len(log)} calls")
                print("  join:", e.last_stats.get("join"), "| exchanges:", e.last_stats.get("exchanges"), "| plan:", e.last_stats.get("plan"))
                print("  aggregate:", e.last_stats.get("agg_mode"), "scan kind", e.last_stats.get("agg_scan_kind"), "scan ms", e.last_stats.get("agg_scan_ms"), "local slots", e.last_stats.get("hash_local_slots"), "attempts", e.last_stats.get("hash_attempts"),
                      "| groups out:", rel.nrows if hasattr(rel, "nrows") else None)
                print("  stats:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in e.last_stats.items() if k not in ("join", "exchanges", "plan")})
                for name, at, d in log:
                    print(f"  {1e3 * at:8.3f} ms  +{1e3 * d:7.3f}  {name}")
        N.Context.call, N.Context.check = orig_call, orig_check


if __name__ == "__main__":
    main()
