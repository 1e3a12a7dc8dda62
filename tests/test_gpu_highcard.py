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


@pytest.mark.parametrize("jit", ["auto", "always"], ids=["interpreted", "specialised"])
@pytest.mark.parametrize("nrows", [65536, 70001, 200003])
def test_streaming_aggregate_over_runs_of_every_length(tmp_path, nrows, jit):
    """Runs of 1 row, a few rows, exactly half a tile (128), more than a tile (300, 1000 rows: tiles without any run start),
    ending at and straddling tile edges.  The sums are exact in f64, so any order of the atomics must give the same bits."""
    import numpy as np

    from minispark_b200 import BlockFile
    from minispark_b200.constants import ColumnType

    rng = np.random.default_rng(nrows)
    lengths = []
    while sum(lengths) < nrows:
        lengths.append(int(rng.choice([1, 1, 1, 2, 3, 4, 5, 7, 64, 127, 128, 129, 255, 256, 300, 1000], p=[0.2, 0.1, 0.1, 0.15, 0.1, 0.1, 0.05, 0.05, 0.02, 0.02, 0.03, 0.02, 0.02, 0.02, 0.01, 0.01])))
    keys = np.repeat(np.arange(len(lengths), dtype=np.int64) * 3 - 1000, lengths)[:nrows]
    x = rng.integers(0, 4000, nrows) / 8.0
    i = rng.integers(-10**6, 10**6, nrows)
    table = tmp_path / f"runs{nrows}.bin"
    BlockFile(table, [("k", ColumnType.INTEGER), ("x", ColumnType.FLOAT), ("i", ColumnType.INTEGER)]).write_data((keys.tolist(), x.tolist(), i.tolist()))
    ns = cases.namespace()
    uniq, start = np.unique(keys, return_index=True)
    with CudaExecutionEngine(jit=jit) as e:
        q = ns.DataFrame(e).table(str(table)).group_by(ns.Col("k")).agg(
            ns.F.sum(ns.Col("x")).alias("s"), ns.F.count().alias("n"), ns.F.min(ns.Col("i")).alias("lo"), ns.F.max(ns.Col("x")).alias("hx"))
        for attempt in range(3):
            if attempt == 2:  # the table leaves the device and comes back: what was derived from its old copy must be gone
                e.drop_table_cache()
            rel, schema = e.execute_to_device(q.task)
            assert e.ctx.stats().last_agg_runs == 1
            assert e.ctx.stats().last_run_index_hit == (1 if attempt == 1 else 0)  # the key column's run index is kept with the table
            if jit == "always":
                assert e.last_stats["agg_scan_kind"] == N.K["MSC_SCAN_KIND_JIT"]
            cols = [rel.column_numpy(c) for c in range(5)]
            assert rel.nrows == len(uniq) and np.array_equal(cols[0], uniq)
            assert np.array_equal(cols[1], np.add.reduceat(x, start)), np.flatnonzero(cols[1] != np.add.reduceat(x, start))[:8]
            assert np.array_equal(cols[2], np.diff(np.append(start, nrows)))
            assert np.array_equal(cols[3], np.minimum.reduceat(i, start))
            assert np.array_equal(cols[4], np.maximum.reduceat(x, start))
            e.release_query()
