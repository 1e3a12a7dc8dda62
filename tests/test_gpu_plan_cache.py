"""Repeated queries through the plugin path (DataFrame.collect -> execute_full_task -> result BlockFile -> collect_results):
the second execution of an identical aggregate task tree becomes a prepared pass (one launch), results stay the oracle's,
and the cache lets go when the table changes -- the counterpart of the reference ThreadEngine's compile cache
(src/mini_spark/execution.py:139-160), which is keyed by the generated program."""

from __future__ import annotations

import os
import time

import pytest

import cases
from minispark_b200 import CudaExecutionEngine
from minispark_b200.execution import task_fingerprint
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu


def _gen(path, sf, seed=1234):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "bench"))
    import gen_tpch

    gen_tpch.write_table(path, "lineitem", sf=sf, rows_per_block=4096, seed=seed)
    return str(path)


def test_repeated_collect_reuses_a_prepared_pass(tmp_path):
    ns = cases.namespace()
    table = _gen(tmp_path / "lineitem.bin", 0.002)
    want = O.run_task(cases.q1(ns, table).task, wire=True)
    with CudaExecutionEngine() as engine:
        plans = []
        for _ in range(4):
            got = cases.q1(ns, table, engine).collect()  # a NEW DataFrame / task tree every time, like a user's loop
            O.assert_rows_equal(got, want, rel=5e-7)
            plans.append(engine.last_stats["plan"])
        assert plans[0] == "one-shot" and all(p.startswith("prepared") for p in plans[1:]), plans
        assert engine.last_stats["scan_kind"] == 2  # MSC_SCAN_KIND_JIT: the kernel specialised for this query
        # the SQL front end builds the same tree: it shares the entry
        got = engine.sql(cases.Q1_SQL.format(table=table)).collect()
        O.assert_rows_equal(got, want, rel=5e-7)
        # another literal is another query
        other = ns.DataFrame(engine).table(table).filter(ns.Col("l_shipdate") <= "1995-01-01").group_by(ns.Col("l_returnflag")).agg(ns.F.count())
        assert task_fingerprint(other.task) != task_fingerprint(cases.q1(ns, table).task)
        O.assert_rows_equal(other.collect(), O.run_task(ns.DataFrame(None).table(table).filter(ns.Col("l_shipdate") <= "1995-01-01")
                                                        .group_by(ns.Col("l_returnflag")).agg(ns.F.count()).task, wire=True))
        assert engine.last_stats["plan"] == "one-shot"
        # the table changes on disk: the prepared pass must not serve the old columns
        time.sleep(0.01)
        _gen(tmp_path / "lineitem.bin", 0.003, seed=99)
        os.utime(table, ns=(time.time_ns(), time.time_ns()))
        want2 = O.run_task(cases.q1(ns, table).task, wire=True)
        assert want2 != want
        got = cases.q1(ns, table, engine).collect()
        assert engine.last_stats["plan"] == "one-shot"
        O.assert_rows_equal(got, want2, rel=5e-7)
        got = cases.q1(ns, table, engine).collect()
        assert engine.last_stats["plan"].startswith("prepared")
        O.assert_rows_equal(got, want2, rel=5e-7)


def test_queries_that_are_not_kept(tmp_path, tables):
    """High-cardinality aggregates, joins below the aggregate and plain scans keep running one-shot, correctly."""
    ns = cases.namespace()
    with CudaExecutionEngine() as engine:
        for name in ("join_then_filter_group", "group_by_int_expr", "filter", "concat_filter"):
            build, _, ordered = cases.DF_CASES[name]
            want = O.run_task(build(ns, tables, None).task, wire=True)
            for _ in range(3):
                O.assert_rows_equal(build(ns, tables, engine).collect(), want, ordered=ordered)
            assert engine.last_stats["plan"].startswith("one-shot"), name  # (never a prepared pass; repeats get specialised kernels)
        # with the cache off nothing is prepared
        engine.plan_cache_enabled = False
        build = cases.DF_CASES["groupby_multi"][0]
        for _ in range(3):
            O.assert_rows_equal(build(ns, tables, engine).collect(), O.run_task(build(ns, tables, None).task, wire=True))
            assert engine.last_stats["plan"] == "one-shot"
        engine.plan_cache_enabled = True
        for _ in range(3):
            O.assert_rows_equal(build(ns, tables, engine).collect(), O.run_task(build(ns, tables, None).task, wire=True))
        assert engine.last_stats["plan"].startswith("prepared")
