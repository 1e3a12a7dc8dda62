"""Dense GROUP BY through every variant of the register-resident interpreter (csrc/gen_regvm.py).

Group counts 2, 3, 4 take the masked variants (per-group reduction in registers), 7 takes the generic
one (one shared-memory update per row).  Results are compared with the f64 oracle at 1e-9 through
``execute_to_device`` (no f32 narrowing) and with the wire result.  A non-finite input must give the
reference's answer too: the masked variants would leak NaN into other groups, so the library reruns
such a scan on the generic kernel (scan.cu).
"""

from __future__ import annotations

import math

import pytest

import cases
from minispark_b200 import BlockFile, CudaExecutionEngine
from minispark_b200.constants import ColumnType
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu

KEYS = [("l_linestatus", 2), ("l_returnflag", 3), ("l_shipinstruct", 4), ("l_shipmode", 7)]


def _query(ns, table, key, engine=None, filtered=True):
    df = ns.DataFrame(engine).table(table)
    if filtered:  # drops about half of the rows, and every row of some lanes
        df = df.filter(ns.Col("l_quantity") > 25.0).filter(ns.Col("l_shipdate") <= "1997-06-01")
    disc = ns.Col("l_extendedprice") * (ns.Lit(1) - ns.Col("l_discount"))
    return df.group_by(ns.Col(key)).agg(
        ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.sum(disc).alias("dp"), ns.F.sum(disc * (ns.Lit(1) + ns.Col("l_tax"))).alias("ch"),
        ns.F.avg(ns.Col("l_tax")).alias("t"), ns.F.count().alias("n"))


def _device_rows(engine, task):
    rel, schema = engine.execute_to_device(task)
    names = [n for n, _ in schema]
    cols = [rel.column_numpy(i) for i in range(len(names))]
    keys = rel.cols[0].dict.export()
    rows = [{names[0]: keys[int(cols[0][r])], **{n: cols[i][r].item() for i, n in enumerate(names) if i}} for r in range(rel.nrows)]
    stats = dict(engine.last_stats)
    engine.release_query()
    return rows, stats


@pytest.mark.parametrize("filtered", [False, True], ids=["all_rows", "filtered"])
@pytest.mark.parametrize(("key", "ngroups"), KEYS, ids=[k for k, _ in KEYS])
def test_dense_group_by_variants_match_f64_oracle(small_lineitem, key, ngroups, filtered):
    ns = cases.namespace()
    want = O.run_task(_query(ns, small_lineitem, key, filtered=filtered).task, wire=False)
    assert len(want) == ngroups
    with CudaExecutionEngine() as e:
        got, stats = _device_rows(e, _query(ns, small_lineitem, key, filtered=filtered).task)
        assert stats["agg_mode"] == "dense"
        wire = _query(ns, small_lineitem, key, e, filtered=filtered).collect()
    O.assert_rows_equal(wire, O.run_task(_query(ns, small_lineitem, key, filtered=filtered).task, wire=True), rel=5e-7)
    want = {r[key]: r for r in want}
    assert sorted(want) == sorted(r[key] for r in got)
    for row in got:
        ref = want[row[key]]
        assert row["n"] == ref["n"]
        for k in ("q", "dp", "ch", "t"):
            assert abs(row[k] - ref[k]) <= 1e-9 * abs(ref[k]), (row[key], k, row[k], ref[k])


@pytest.mark.parametrize("bad", [math.inf, -math.inf, math.nan], ids=["inf", "-inf", "nan"])
def test_non_finite_values_stay_in_their_group(tmp_path, bad):
    """Python sums keep inf / nan inside the group they occur in; so must the GPU (rerun on the generic kernel)."""
    path = tmp_path / "nf.bin"
    rows = [{"k": "abc"[i % 3], "v": float(i % 17) + 0.25} for i in range(3000)]
    rows[1234]["v"] = bad  # group 'b'
    BlockFile(path, [("k", ColumnType.STRING), ("v", ColumnType.FLOAT)]).write_rows(rows)
    ns = cases.namespace()

    def build(engine=None):
        return ns.DataFrame(engine).table(str(path)).group_by(ns.Col("k")).agg(ns.F.sum(ns.Col("v") * ns.Lit(2)).alias("s"), ns.F.count().alias("n"))

    want = {r["k"]: r for r in O.run_task(build().task, wire=False)}
    with CudaExecutionEngine() as e:
        got, _ = _device_rows(e, build().task)
    assert sorted(r["k"] for r in got) == ["a", "b", "c"]
    for row in got:
        ref = want[row["k"]]
        assert row["n"] == ref["n"] == 1000
        if math.isnan(ref["s"]):
            assert math.isnan(row["s"])
        else:
            assert row["s"] == ref["s"], (row["k"], row["s"], ref["s"])
    assert math.isfinite(want["a"]["s"]) and math.isfinite(want["c"]["s"]) and not math.isfinite(want["b"]["s"])
