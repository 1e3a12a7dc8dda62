#!/usr/bin/env python
"""Timings of BASELINE.json configs 4 and 5 on one B200 (they are parity cases for bench.py; this prints numbers).

    python bench/bench_configs.py --sf 2 [--reps 5]

config 4: SELECT l_orderkey, SUM(l_quantity), AVG(l_extendedprice) FROM lineitem GROUP BY l_orderkey   (hash aggregate)
config 5: orders JOIN lineitem ON o_orderkey = l_orderkey WHERE o_orderdate BETWEEN .. AND l_shipmode LIKE '%AIR%'
          GROUP BY o_orderpriority  (hash join with late materialisation, then a dense aggregate)
Tables are device resident when the clock starts (first execution ingests them); one JSON line per config.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "bench"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))
os.environ["TZ"] = "UTC"
time.tzset()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--jit", default=None, help="engine jit mode: auto | always | never")
    ap.add_argument("--only", default="", help="comma-separated config name prefixes to run")
    args = ap.parse_args()
    import cases
    import gen_tpch
    from minispark_b200 import CudaExecutionEngine

    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    folder = base / f"minispark_b200_cfg_{os.getuid()}"
    folder.mkdir(parents=True, exist_ok=True)
    lineitem, orders = folder / f"lineitem6_sf{args.sf:g}.bin", folder / f"orders_sf{args.sf:g}.bin"
    if not lineitem.exists():
        gen_tpch.write_table(lineitem, "lineitem", sf=args.sf, columns=["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipmode"])
    if not orders.exists():
        gen_tpch.write_table(orders, "orders", sf=args.sf, columns=["o_orderkey", "o_orderdate", "o_orderpriority"])
    ns = cases.namespace()

    def config4(e):  # noqa: ANN001, ANN202
        return ns.DataFrame(e).table(str(lineitem)).group_by(ns.Col("l_orderkey")).agg(
            ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.avg(ns.Col("l_extendedprice")).alias("p"))

    def config5(e):  # noqa: ANN001, ANN202
        o = ns.DataFrame(e).table(str(orders)).alias("o")
        l = ns.DataFrame().table(str(lineitem)).alias("l")
        return (o.join(l, on=ns.Col("o.o_orderkey") == ns.Col("l.l_orderkey"), how="inner")
                .filter(ns.Col("o.o_orderdate").between("1994-01-01", "1994-12-31"))
                .filter(ns.Col("l.l_shipmode").like("%AIR%"))
                .group_by(ns.Col("o.o_orderpriority")).agg(ns.F.count(), ns.F.sum(ns.Col("l.l_extendedprice")).alias("rev")))

    def filter_project(e):  # noqa: ANN001, ANN202  (FilterTask + ProjectTask, tasks.py:79-84,167-177: about half of the rows survive)
        return (ns.DataFrame(e).table(str(lineitem)).filter(ns.Col("l_quantity") > 25.0)
                .select(ns.Col("l_orderkey"), (ns.Col("l_extendedprice") * ns.Col("l_quantity")).alias("v"), ns.Col("l_shipmode")))

    def project_only(e):  # noqa: ANN001, ANN202
        return ns.DataFrame(e).table(str(lineitem)).select((ns.Col("l_extendedprice") * ns.Col("l_quantity")).alias("v"))

    def q1_by_shipmode(e):  # noqa: ANN001, ANN202  (Q1's aggregates over 7 groups: 49 accumulator cells)
        disc = ns.Col("l_extendedprice") * (ns.Lit(1) - ns.Col("l_discount"))
        return ns.DataFrame(e).table(str(lineitem)).group_by(ns.Col("l_shipmode")).agg(
            ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.sum(ns.Col("l_extendedprice")).alias("p"), ns.F.sum(disc).alias("dp"),
            ns.F.sum(disc * (ns.Lit(1) + ns.Col("l_tax"))).alias("ch"), ns.F.avg(ns.Col("l_discount")).alias("d"), ns.F.count().alias("n"))

    configs = (("q1_by_shipmode", q1_by_shipmode), ("config4_high_cardinality_group_by", config4), ("config5_join_filter_like_group_by", config5),
               ("filter_project", filter_project), ("project_only", project_only))
    only = [x for x in args.only.split(",") if x]
    with CudaExecutionEngine(device=0, shard=(0, 1), jit=args.jit) as e:
        for name, build in configs:
            if only and not any(name.startswith(x) for x in only):
                continue
            task = build(e).task
            times, rows_out = [], 0
            for i in range(args.reps + 1):
                t0 = time.perf_counter()
                rel, _ = e.execute_to_device(task)
                e.ctx.call("msc_sync")
                dt = time.perf_counter() - t0
                rows_out = rel.nrows
                stats = dict(e.last_stats)
                e.release_query()
                if i:
                    times.append(dt)
            nl = e._tables[str(lineitem)].nrows
            med = statistics.median(times)
            print(json.dumps({"config": name, "sf": args.sf, "lineitem_rows": nl, "result_rows": rows_out, "ms": round(1e3 * med, 3),
                              "lineitem_rows_per_s": nl / med, "passes_ms": [round(1e3 * t, 2) for t in times],
                              "last_kernel_ms": stats.get("kernel_ms"), "agg_mode": stats.get("agg_mode"), "jit": e.jit,
                              "scan_kind": stats.get("scan_kind"), "agg_scan_kind": stats.get("agg_scan_kind"), "agg_scan_ms": stats.get("agg_scan_ms"), "jit_compiles": e.ctx.stats().jit_compiles}), flush=True)
    if os.environ.get("MSC_BENCH_KEEP") is None:
        lineitem.unlink(missing_ok=True)
        orders.unlink(missing_ok=True)


if __name__ == "__main__":
    main()
