"""Builds every query of tests/cases.py TWICE -- with this package's host mirror and with the REAL reference classes
(``mini_spark`` imported from the reference checkout) -- and checks that ``lowering.lower_task`` turns both task trees
into the same logical plan: the boundary of ``CudaExecutionEngine.execute_full_task`` is the reference's own ``Task`` /
``Col`` tree, recognised by class and attribute names.  Run by tests/test_reference_boundary.py in a
subprocess (CPU only; needs /root/reference, so it is skipped on the GPU box).

    python tests/ref_boundary_runner.py <scratch folder>
"""
import os
import sys
import time
from copy import deepcopy
from pathlib import Path

os.environ["TZ"] = "UTC"
time.tzset()
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "bench", ROOT / "tests" / "golden"):
    sys.path.insert(0, str(p))


def plan_text(plan) -> str:  # noqa: ANN001
    def walk(node) -> list[str]:  # noqa: ANN001
        out = [type(node).__name__ + " " + ", ".join(f"{n}:{t.name}" for n, t in node.schema)]
        for attr in ("child", "left", "right"):
            sub = getattr(node, attr, None)
            if sub is not None:
                out += walk(sub)
        return out

    return plan.describe() + "\n" + "\n".join(walk(plan))


def main() -> None:
    scratch = Path(sys.argv[1])
    import make_golden

    make_golden.install_perfetto_stub(scratch / "stub")
    sys.path.insert(0, "/root/reference/src")
    os.chdir(scratch)
    import cases
    import gen_tpch
    from minispark_b200 import lowering as L

    mirror, ref = cases.namespace("mirror"), cases.namespace("reference")
    assert ref.DataFrame.__module__.startswith("mini_spark."), ref.DataFrame.__module__
    tables = cases.write_tables(scratch / "tables")
    lineitem = scratch / "lineitem.bin"
    gen_tpch.write_table(lineitem, "lineitem", sf=0.001, rows_per_block=2048)
    builders = {name: (lambda ns, build=build: build(ns, tables, None)) for name, (build, _, _) in cases.DF_CASES.items()}
    builders["q1"] = lambda ns: cases.q1(ns, str(lineitem))
    checked = 0
    for name, build in sorted(builders.items()):
        texts = []
        for ns in (mirror, ref):
            task = deepcopy(build(ns).task)
            assert type(task).__module__.split(".")[0] == ("minispark_b200" if ns is mirror else "mini_spark"), type(task)
            task.validate_schema()
            texts.append(plan_text(L.lower_task(task)))
        assert texts[0] == texts[1], f"{name}: mirror-built and reference-built trees lower differently\n{texts[0]}\n---\n{texts[1]}"
        checked += 1
    print("boundary ok", checked)


if __name__ == "__main__":
    main()
