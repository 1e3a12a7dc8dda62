"""High-cardinality GROUP BY on integer keys, at sizes where the library counts the key column's runs first
(>= 65536 rows): a sorted key (lineitem's l_orderkey) takes the streaming aggregate over its runs (MODE_RUNS, no hash
table), an unsorted one the hash table sized by the run count.  Both against the f64 oracle."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import pytest

import cases
from minispark_b200 import CudaExecutionEngine
from minispark_b200 import native as N
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "bench"))


@pytest.fixture(scope="module")
def lineitem_72k(tmp_path_factory) -> str:
    import gen_tpch

    path = tmp_path_factory.mktemp("hc") / "lineitem.bin"
    gen_tpch.write_table(path, "lineitem", sf=0.012, columns=["l_orderkey", "l_suppkey", "l_quantity", "l_extendedprice"], rows_per_block=30000)
    return str(path)


def _query(ns, table, key, engine=None):
    return ns.DataFrame(engine).table(table).group_by(ns.Col(key)).agg(
        ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.avg(ns.Col("l_extendedprice")).alias("p"), ns.F.count().alias("n"),
        ns.F.min(ns.Col("l_quantity")).alias("lo"), ns.F.max(ns.Col("l_extendedprice")).alias("hi"))


@pytest.mark.parametrize(("key", "kind", "jit"), [("l_orderkey", "RUNS", "auto"), ("l_orderkey", "JIT", "always"), ("l_suppkey", "VM", "auto")],
                         ids=["sorted_key_streams", "sorted_key_streams_specialised", "unsorted_key_hashes"])
def test_high_cardinality_group_by_matches_f64_oracle(lineitem_72k, key, kind, jit):
    ns = cases.namespace()
    want = {r[key]: r for r in O.run_task(_query(ns, lineitem_72k, key).task, wire=False)}
    with CudaExecutionEngine(jit=jit) as e:
        rel, schema = e.execute_to_device(_query(ns, lineitem_72k, key).task)
        assert e.last_stats["agg_mode"] == "hash"
        expected = {N.K[f"MSC_SCAN_KIND_{kind}"]}
        if kind == "RUNS" and os.environ.get("MSC_SCAN_JIT") == "2":  # the whole-suite rerun that forces specialised kernels
            expected.add(N.K["MSC_SCAN_KIND_JIT"])
        assert e.last_stats["agg_scan_kind"] in expected
        names = [n for n, _ in schema]
        cols = [rel.column_numpy(i) for i in range(len(names))]
        got = {int(cols[0][r]): {n: cols[i][r].item() for i, n in enumerate(names)} for r in range(rel.nrows)}
        e.release_query()
        wire = _query(ns, lineitem_72k, key, e).collect()
    assert len(got) == len(want) == rel.nrows and sorted(got) == sorted(want)
    for k, ref in want.items():
        row = got[k]
        assert row["n"] == ref["n"]
        for name in ("q", "p", "lo", "hi"):
            assert abs(row[name] - ref[name]) <= 1e-9 * max(abs(ref[name]), 1e-300), (k, name, row[name], ref[name])
    O.assert_rows_equal(wire, O.run_task(_query(ns, lineitem_72k, key).task, wire=True), rel=5e-7)
