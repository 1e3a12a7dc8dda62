"""Parity of the CUDA engine (through the C-ABI) with the reference's golden vectors and the oracle.

All tests here need a B200 (``-m gpu``).  They read like the reference's engine-parametrised tests
(tests/test_execution.py, tests/test_e2e.py) with ``CudaExecutionEngine`` as the engine.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

import cases
from golden import golden_io
from minispark_b200 import CudaExecutionEngine, DataFrame
from minispark_b200.parser import parse_sql
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def engine():
    with CudaExecutionEngine() as e:
        yield e


@pytest.mark.parametrize(("name", "sql", "expected"), cases.SQL_CASES, ids=[c[0] for c in cases.SQL_CASES])
def test_e2e_sql_vectors(engine, tables, name, sql, expected):
    df = parse_sql(sql.format(**tables))
    df.engine = engine
    O.assert_rows_equal(df.collect(), expected, ordered=name in cases.ORDERED_SQL)


@pytest.mark.parametrize("name", sorted(cases.DF_CASES))
def test_dataframe_cases(engine, tables, name):
    build, expected, ordered = cases.DF_CASES[name]
    got = build(cases.namespace(), tables, engine).collect()
    fixture = golden_io.load(GOLDEN / "df_cases.json")
    if expected is not None:
        O.assert_rows_equal(got, expected, ordered=ordered)
    if name in fixture:  # rows produced by the real reference PythonExecutionEngine
        O.assert_rows_equal(got, fixture[name], ordered=ordered)
    oracle_rows = O.run_task(build(cases.namespace(), tables, None).task, wire=True)
    O.assert_rows_equal(got, oracle_rows, ordered=ordered)


def test_engine_sql_entry_point(engine, tables):
    rows = engine.sql(f"SELECT * FROM '{tables['fruits']}';").collect()
    assert rows == [dict(r) for r in cases.FRUITS]


@pytest.mark.parametrize("layout", ["native", "wide"])
def test_q1_small_matches_reference_fixture_and_f64_oracle(small_lineitem, layout):
    fixture = golden_io.load(GOLDEN / "q1_small.json")["q1_wire"]
    with CudaExecutionEngine(layout=layout) as e:
        got = cases.q1(cases.namespace(), small_lineitem, e).collect()
        # wire parity with the real reference.  The reference rounds every per-block partial sum to f32
        # before merging (tasks.py:373 -> io.py:91-94); the GPU rounds once, so multi-block FLOAT sums may
        # differ by a few f32 ulps (<= 4 ulp = 5e-7 relative).  Counts stay exact.
        O.assert_rows_equal(got, fixture, rel=5e-7)
        # full precision (no f32 narrowing): 1e-9 relative against the f64 oracle, ints exact
        rel, schema = e.execute_to_device(cases.q1(cases.namespace(), small_lineitem).task)
        names = [n for n, _ in schema]
        keys = rel.cols[0].dict.export()
        dev = {}
        cols = [rel.column_numpy(i) for i in range(len(names))]
        for r in range(rel.nrows):
            dev[keys[int(cols[0][r])]] = {n: cols[i][r].item() for i, n in enumerate(names) if i}
        e.release_query()
    oracle = O.run_task(cases.q1(cases.namespace(), small_lineitem).task, wire=False)
    assert len(oracle) == len(dev) == 3
    for row in oracle:
        mine = dev[row["l_returnflag"]]
        for k, v in row.items():
            if k == "l_returnflag":
                continue
            if isinstance(v, int):
                assert mine[k] == v, k
            else:
                assert abs(mine[k] - v) <= 1e-9 * abs(v), (k, mine[k], v)


def test_q1_sql_text(small_lineitem):
    fixture = golden_io.load(GOLDEN / "q1_small.json")["q1_wire"]
    with CudaExecutionEngine() as e:
        got = e.sql(cases.Q1_SQL.format(table=small_lineitem)).collect()
    O.assert_rows_equal(got, fixture, rel=5e-7)


def test_high_cardinality_group_by(small_lineitem):
    ns = cases.namespace()

    def build(engine):
        return ns.DataFrame(engine).table(small_lineitem).group_by(ns.Col("l_orderkey")).agg(
            ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.avg(ns.Col("l_extendedprice")).alias("p"), ns.F.count())

    with CudaExecutionEngine() as e:
        got = build(e).collect()
        assert e.last_stats["agg_mode"] == "hash"
    O.assert_rows_equal(got, O.run_task(build(None).task, wire=True), rel=5e-7)


def test_join_filter_like_on_generated_tables(small_lineitem, tmp_path):
    import gen_tpch

    orders = str(tmp_path / "orders_small.bin")
    gen_tpch.write_table(orders, "orders", sf=0.002, rows_per_block=1024)
    ns = cases.namespace()

    def build(engine):
        o = ns.DataFrame(engine).table(orders).alias("o")
        l = ns.DataFrame().table(small_lineitem).alias("l")
        return (o.join(l, on=ns.Col("o.o_orderkey") == ns.Col("l.l_orderkey"), how="inner")
                .filter(ns.Col("o.o_orderdate").between("1994-01-01", "1994-12-31"))
                .filter(ns.Col("l.l_shipmode").like("%AIR%"))
                .group_by(ns.Col("o.o_orderpriority")).agg(ns.F.count(), ns.F.sum(ns.Col("l.l_extendedprice")).alias("rev")))

    with CudaExecutionEngine() as e:
        got = build(e).collect()
    want = O.run_task(build(None).task, wire=True)
    assert len(want) > 0
    O.assert_rows_equal(got, want, rel=5e-7)


def test_division_by_zero_raises(engine, tables):
    from minispark_b200 import Col, ExecutionError

    with pytest.raises(ExecutionError):
        DataFrame(engine).table(tables["users"]).select(Col("age") / (Col("user_id") - 3)).collect()


def test_int_overflow_on_write_raises(engine, tmp_path):
    from minispark_b200 import BlockFile, Col

    path = tmp_path / "big.bin"
    BlockFile(path).write_rows([{"num1": 2**31 - 1, "num2": 2**31 - 1}])
    with pytest.raises(OverflowError):  # PythonExecutionEngine raises OverflowError at io.py:90 (ThreadEngine wraps to -2)
        DataFrame(engine).table(str(path)).select(Col("num1") + Col("num2")).collect()


def test_table_cache_and_reingest(engine, tables):
    df = lambda: DataFrame(engine).table(tables["orders"])  # noqa: E731
    a = df().collect()
    engine.drop_table_cache()
    assert df().collect() == a


def test_trace_of_a_query(tables, tmp_path):
    """TRACER.save() after a query gives a Perfetto trace with the reference's slices (execution.py:69-78) on the main
    track and the library's CUDA-event durations on the GPU's track."""
    from minispark_b200.utils import TRACER, parse_trace

    TRACER.enable()
    try:
        with CudaExecutionEngine() as engine:
            build = cases.DF_CASES["groupby_multi"][0]
            rows = build(cases.namespace(), tables, engine).collect()
            assert len(rows) == 3
            TRACER.save(str(tmp_path / "trace.pftrace"))
    finally:
        TRACER.enable(False)
    packets = parse_trace((tmp_path / "trace.pftrace").read_bytes())
    names = [p["event"]["name"] for p in packets if "event" in p and p["event"]["name"]]
    assert "execute full task" in names and "Execution" in names and "write result BlockFile" in names
    assert any(n.startswith("scan: filter + aggregate") for n in names) and any(n.startswith("load table block") for n in names)
    gpu_tracks = [p["track"]["uuid"] for p in packets if "track" in p and p["track"]["name"].startswith("GPU")]
    assert len(gpu_tracks) == 1
    on_gpu = [p for p in packets if "event" in p and p["event"]["track"] == gpu_tracks[0]]
    assert len(on_gpu) >= 4 and len(on_gpu) % 2 == 0
    begins = sum(1 for p in packets if "event" in p and p["event"]["type"] == 1)
    ends = sum(1 for p in packets if "event" in p and p["event"]["type"] == 2)
    assert begins == ends
