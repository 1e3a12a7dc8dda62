#!/usr/bin/env python
"""Where does the end-to-end time of one Q1 pass go?  (pinned BlockFile image -> H2D -> decode -> scan -> host)

    python bench/e2e_probe.py --sf 4 [--reps 3]

Prints the raw pinned->device copy bandwidth of the box, then per pass: msc_table_load wall time
(`ingest_ms`), bytes copied, whole-query wall time.  Diagnostic only; bench.py reports the numbers.
"""

from __future__ import annotations

import argparse
import ctypes as C
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
    ap.add_argument("--sf", type=float, default=4.0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import bench as B
    import cases
    from minispark_b200 import CudaExecutionEngine

    path, gen_s = B.ensure_table(args.sf, 0)
    engine = CudaExecutionEngine(device=0, shard=(0, 1))
    nbytes = path.stat().st_size
    pinned = C.c_void_p()
    t0 = time.perf_counter()
    engine.ctx.call("msc_host_alloc", nbytes, C.byref(pinned))
    t1 = time.perf_counter()
    view = (C.c_char * nbytes).from_address(pinned.value)
    with open(path, "rb") as f:
        f.readinto(view)
    t2 = time.perf_counter()
    print(f"table {nbytes / 1e9:.3f} GB  generate {gen_s:.2f}s  host_alloc {t1 - t0:.3f}s  read {t2 - t1:.3f}s", flush=True)

    dev = C.c_void_p()
    chunk = min(nbytes, 1 << 30)
    engine.ctx.call("msc_dev_alloc", chunk, C.byref(dev))
    for i in range(3):
        t0 = time.perf_counter()
        engine.ctx.call("msc_memcpy_h2d", dev, pinned, chunk)
        dt = time.perf_counter() - t0
        print(f"raw pinned->device copy of {chunk / 1e9:.2f} GB: {dt * 1e3:.1f} ms = {chunk / dt / 1e9:.1f} GB/s", flush=True)
    # per-chunk bandwidth across the whole image (host memory locality / IOMMU effects show up as slow chunks)
    step = 256 << 20
    for rep in range(2):
        line = []
        for off in range(0, nbytes, step):
            n = min(step, nbytes - off)
            t0 = time.perf_counter()
            engine.ctx.call("msc_memcpy_h2d", dev, C.c_void_p(pinned.value + off), n)
            line.append(f"{n / (time.perf_counter() - t0) / 1e9:.0f}")
        print(f"per-256MB-chunk GB/s (rep {rep}): " + " ".join(line), flush=True)
    engine.ctx.call("msc_dev_free", dev)

    engine.register_table_image(str(path), pinned.value, nbytes)
    task = engine.sql(cases.Q1_SQL.format(table=str(path))).task
    for i in range(args.reps + 1):
        engine.drop_table_cache()
        t0 = time.perf_counter()
        rel, schema = engine.execute_to_device(task)
        t1 = time.perf_counter()
        host = [rel.column_numpy(c) for c in range(len(schema))]
        t2 = time.perf_counter()
        st = dict(engine.last_stats)
        engine.release_query()
        print(f"pass {i}: total {1e3 * (t2 - t0):.1f} ms  execute_to_device {1e3 * (t1 - t0):.1f}  ingest_ms {st.get('ingest_ms', 0):.1f}  "
              f"query_s {1e3 * st.get('query_s', 0):.1f} ms  bytes {st.get('ingest_bytes', 0) / 1e9:.3f} GB  "
              f"-> {st.get('ingest_bytes', 0) / max(st.get('ingest_ms', 1), 1e-9) / 1e6:.1f} GB/s in msc_table_load", flush=True)
    engine.ctx.call("msc_host_free", pinned)
    engine.close()
    if os.environ.get("MSC_BENCH_KEEP") is None:
        path.unlink(missing_ok=True)


if __name__ == "__main__":
    main()
