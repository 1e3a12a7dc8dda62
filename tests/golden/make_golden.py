"""Generate golden vectors by running the REAL reference (PythonExecutionEngine) in the build container.

    TZ=UTC python tests/golden/make_golden.py

Needs /root/reference (absent on the GPU box, which only ever reads the committed JSON).  The
reference imports ``perfetto`` at module import time (src/mini_spark/utils.py:15-16); a no-op stub
is put on sys.path first.  Outputs:
  tests/golden/df_cases.json   every tests/cases.py DF_CASES query the reference can run
  tests/golden/q1_small.json   TPC-H Q1 over the generator's 3-block sf0.002 lineitem
"""

from __future__ import annotations

import os
import sys
import tempfile
import time
from pathlib import Path

os.environ["TZ"] = "UTC"
time.tzset()
HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REFERENCE = Path("/root/reference/src")


def install_perfetto_stub(folder: Path) -> None:
    pkg = folder / "perfetto"
    (pkg / "protos" / "perfetto" / "trace").mkdir(parents=True)
    (pkg / "trace_builder").mkdir(parents=True)
    for d in (pkg, pkg / "protos", pkg / "protos" / "perfetto", pkg / "protos" / "perfetto" / "trace", pkg / "trace_builder"):
        (d / "__init__.py").write_text("")
    (pkg / "protos" / "perfetto" / "trace" / "perfetto_trace_pb2.py").write_text(
        "class TrackEvent:\n    TYPE_SLICE_BEGIN = 1\n    TYPE_SLICE_END = 2\n")
    (pkg / "trace_builder" / "proto_builder.py").write_text(
        "import types\nclass TracePacket: pass\nclass TraceProtoBuilder:\n"
        "    def add_packet(self):\n        ns = types.SimpleNamespace\n"
        "        return ns(track_event=ns(), track_descriptor=ns())\n    def serialize(self): return b''\n")
    sys.path.insert(0, str(folder))


def main() -> None:
    scratch = Path(tempfile.mkdtemp(prefix="golden_"))
    install_perfetto_stub(scratch / "stub")
    sys.path.insert(0, str(REFERENCE))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    sys.path.insert(0, str(ROOT / "bench"))
    os.chdir(scratch)  # the reference writes ./shuffle relative to the CWD (constants.py:11)

    import cases
    import gen_tpch
    from golden import golden_io
    from mini_spark.execution import PythonExecutionEngine  # the real reference

    ref = cases.namespace("reference")
    tables = cases.write_tables(scratch / "tables", ref)  # written by the reference's own BlockFile
    out: dict[str, list[dict]] = {}
    for name, (build, _, _) in sorted(cases.DF_CASES.items()):
        try:
            with PythonExecutionEngine() as engine:
                out[name] = build(ref, tables, engine).collect()
        except Exception as e:  # noqa: BLE001
            print(f"reference cannot run {name}: {type(e).__name__}: {e}")
    golden_io.dump(HERE / "df_cases.json", out)
    print(f"df_cases.json: {len(out)} cases")

    lineitem = scratch / "lineitem_small.bin"
    gen_tpch.write_table(lineitem, "lineitem", sf=0.002, rows_per_block=4096)
    with PythonExecutionEngine() as engine:
        q1 = cases.q1(ref, str(lineitem), engine).collect()
    golden_io.dump(HERE / "q1_small.json", {"q1_wire": q1})
    print(f"q1_small.json: {len(q1)} groups")


if __name__ == "__main__":
    main()
