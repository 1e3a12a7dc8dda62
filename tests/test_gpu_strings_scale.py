"""STRING columns at scale (reference LIKE: src/mini_spark/sql.py:178-194): a column of more than a
million DISTINCT values.  Its device dictionary then is the column -- offsets + bytes of every value, plus one code per
row -- and LIKE / equality / ordering run as kernels over those raw bytes (msc_dict_like) or over the codes."""

from __future__ import annotations

import re
import sys
from pathlib import Path

import numpy as np
import pytest

import cases
from minispark_b200 import CudaExecutionEngine

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "bench"))


@pytest.fixture(scope="module")
def comments(tmp_path_factory):
    import gen_tpch

    path = tmp_path_factory.mktemp("strings") / "lineitem_comments.bin"
    gen_tpch.write_table(path, "lineitem", sf=0.2, columns=["l_orderkey", "l_quantity", "l_comment"])
    # the oracle's view of the column, decoded with numpy (the row loop of the pure-Python oracle takes minutes at this size)
    from minispark_b200.io import BlockFile

    bf = BlockFile(path)
    values: list[str] = []
    qty: list[float] = []
    for b in range(len(bf.block_starts)):
        cols = bf.read_block_data_columns_by_id(b) if hasattr(bf, "read_block_data_columns_by_id") else None
        if cols is None:
            break
        values.extend(cols[2])
        qty.extend(cols[1])
    return str(path), values, np.asarray(qty)


def test_like_and_equality_over_a_million_distinct_strings(comments):
    path, values, qty = comments
    assert len(values) > 1_100_000 and len(set(values)) > 1_000_000
    ns = cases.namespace()
    with CudaExecutionEngine() as engine:
        # LIKE '%foo%' -> rows in input order (reference: re.match of the escaped pattern, sql.py:178-179)
        got = ns.DataFrame(engine).table(path).filter(ns.Col("l_comment").like("%foo%")).select(ns.Col("l_comment"), ns.Col("l_quantity")).collect()
        want = [(v, q) for v, q in zip(values, qty) if "foo" in v]
        assert len(want) > 100
        assert [(r["l_comment"], r["l_quantity"]) for r in got] == [(v, float(q)) for v, q in want]
        # a pattern with _ and an anchored prefix, aggregated
        rx = re.compile("^" + re.escape("ab_d%").replace("%", ".*").replace("_", ".") + "$", re.S)
        n = ns.DataFrame(engine).table(path).filter(ns.Col("l_comment").like("ab_d%")).group_by(ns.Col("l_quantity")).agg(ns.F.count()).collect()
        assert sum(r["count"] for r in n) == sum(1 for v in values if rx.match(v))
        # equality with one of the values, and with a string that is not there
        probe = values[len(values) // 3]
        hits = ns.DataFrame(engine).table(path).filter(ns.Col("l_comment") == probe).select(ns.Col("l_orderkey")).collect()
        assert len(hits) == values.count(probe) >= 1
        assert ns.DataFrame(engine).table(path).filter(ns.Col("l_comment") == "no such comment").collect() == []
        # ordering against a literal (Python string order)
        lo = ns.DataFrame(engine).table(path).filter(ns.Col("l_comment") < "b").group_by(ns.Col("l_quantity")).agg(ns.F.count()).collect()
        assert sum(r["count"] for r in lo) == sum(1 for v in values if v < "b")
        # GROUP BY the near-unique column itself
        groups = ns.DataFrame(engine).table(path).group_by(ns.Col("l_comment")).agg(ns.F.count()).collect()
        assert len(groups) == len(set(values)) and sum(r["count"] for r in groups) == len(values)
        print("ingest_ms", engine.last_stats.get("ingest_ms"))
