#!/usr/bin/env python
"""Host-side breakdown of one prepared Q1 pass (where do the microseconds between the launches go?).

    python bench/step_probe.py --sf 15 --steps 50

Wraps the C-ABI calls of PreparedAggregate.run with wall-clock timers.  Diagnostic only.
"""

from __future__ import annotations

import argparse
import collections
import os
import sys
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
    ap.add_argument("--sf", type=float, default=15.0)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    import bench as B
    import cases
    from minispark_b200 import CudaExecutionEngine
    from minispark_b200 import native as N

    path, _ = B.ensure_table(args.sf, 0)
    engine = CudaExecutionEngine(device=0, shard=(0, 1))
    task = engine.sql(cases.Q1_SQL.format(table=str(path))).task
    prepared = engine.prepare(task)
    acc: dict[str, float] = collections.defaultdict(float)
    real_call = N.Context.call

    def timed_call(self, name, *a):  # noqa: ANN001, ANN002, ANN202
        t0 = time.perf_counter()
        try:
            return real_call(self, name, *a)
        finally:
            acc[name] += time.perf_counter() - t0

    for _ in range(5):
        prepared.run()
        engine.release_query()
    engine.ctx.call("msc_sync")
    N.Context.call = timed_call
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        prepared.run()
        t2 = time.perf_counter()
        engine.release_query()
        acc["<run() total>"] += t2 - t1
        acc["<release_query>"] += time.perf_counter() - t2
    total = time.perf_counter() - t0
    N.Context.call = real_call
    print(f"{args.steps} passes, {1e3 * total / args.steps:.3f} ms per pass, scan kernel {prepared.scan_stats['scan_ms']:.3f} ms")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"  {k:28s} {1e6 * v / args.steps:8.1f} us per pass")
    engine.close()
    if os.environ.get("MSC_BENCH_KEEP") is None:
        path.unlink(missing_ok=True)


if __name__ == "__main__":
    main()
