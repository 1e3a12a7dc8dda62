#!/usr/bin/env python
"""Times the REAL reference PythonExecutionEngine (unmodified, from the git-ignored copy under baseline/_ref) on TPC-H Q1.

    python bench/ref_python_engine.py <lineitem BlockFile> [max seconds]

Baseline infrastructure only (bench.py's `cpu_baseline_python`): the reference engine is single-threaded by construction
(src/mini_spark/execution.py:69-83), so this is a 1-core number; it is quoted next to the GPU engine's, never as a target.
`baseline/_ref/mini_spark` is the reference's package directory as `pip install --target baseline/_ref` would place it
(the pip build itself needs hatchling and Python >= 3.13, both absent here); `baseline/_ref/perfetto` is a
no-op stand-in for the tracing dependency the reference imports at module load (src/mini_spark/utils.py:15-16).
Prints one JSON object.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time
from pathlib import Path

os.environ["TZ"] = "UTC"
time.tzset()
ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"


def main() -> None:
    table = sys.argv[1]
    if not (REF / "mini_spark").is_dir():
        print(json.dumps({"unavailable": "baseline/_ref/mini_spark is missing (build() copies it where /root/reference exists)"}))
        return
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT / "tests"))
    os.chdir(tempfile.mkdtemp(prefix="ref_engine_"))  # the reference writes ./shuffle relative to the CWD (constants.py:11)
    import cases
    from mini_spark.execution import PythonExecutionEngine
    from mini_spark.io import BlockFile

    ns = cases.namespace("reference")
    rows = sum(BlockFile(Path(table)).rows_per_block()) if hasattr(BlockFile, "rows_per_block") else None
    t0 = time.perf_counter()
    with PythonExecutionEngine() as engine:
        result = cases.q1(ns, table, engine).collect()
    dt = time.perf_counter() - t0
    count = sum(r["count_order"] for r in result)
    print(json.dumps({"rows": rows or count, "seconds": dt, "rows_per_s": (rows or count) / dt, "groups": len(result), "count": count,
                      "python": sys.version.split()[0]}))


if __name__ == "__main__":
    main()
