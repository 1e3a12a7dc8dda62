"""Hash join on the GPU in both forms -- the lookup fused into the consuming scan (MSC_OP_PROBE; build sides without
duplicate keys) and the materialising build / probe / emit (msc_hash_join) -- against the oracle's restatement of
BroadcastHashJoinTask.generate_chunks (src/mini_spark/tasks.py:201-240).  Multi-join plans (the parser loops
over joins, parser.py:131-133) and joins whose result is not aggregated (row order = right-row major) included."""

from __future__ import annotations

import numpy as np
import pytest

import cases
from minispark_b200 import BlockFile, CudaExecutionEngine
from minispark_b200.constants import ColumnType
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu
JOIN_CASES = ["self_join", "join_string_keys", "join_then_filter_group"]
JOIN_SQL = [c for c in cases.SQL_CASES if "JOIN" in c[1]]


@pytest.fixture(params=["fused", "fused-materialised-build", "pairs"])
def engine(request):
    with CudaExecutionEngine() as e:
        e.fused_probe = request.param != "pairs"
        e.fused_build = request.param == "fused"
        yield e


@pytest.mark.parametrize("name", JOIN_CASES)
def test_dataframe_joins(engine, tables, name):
    build, _, ordered = cases.DF_CASES[name]
    want = O.run_task(build(cases.namespace(), tables, None).task, wire=True)
    O.assert_rows_equal(build(cases.namespace(), tables, engine).collect(), want, ordered=ordered)


@pytest.mark.parametrize(("name", "sql", "expected"), JOIN_SQL, ids=[c[0] for c in JOIN_SQL])
def test_sql_joins(engine, tables, name, sql, expected):
    got = engine.sql(sql.format(**tables)).collect()
    O.assert_rows_equal(got, expected)
    if engine.fused_probe:  # users.user_id is unique: every one of the reference's SQL join vectors takes the fused form
        assert engine.last_stats["join"].startswith("lookup fused"), engine.last_stats["join"]


@pytest.fixture(scope="module")
def star(tmp_path_factory):
    """A small star schema: facts(n rows) -> dims by integer id and by string code; one dimension has a duplicate key."""
    folder = tmp_path_factory.mktemp("star")
    rng = np.random.default_rng(5)
    n = 20_000
    facts = {"f_id": list(range(n)), "f_dim": rng.integers(0, 600, n).tolist(), "f_code": [f"c{int(v):03d}" for v in rng.integers(0, 50, n)],
             "f_x": (rng.integers(0, 8000, n) / 8.0).tolist()}
    BlockFile(folder / "facts.bin", [("f_id", ColumnType.INTEGER), ("f_dim", ColumnType.INTEGER), ("f_code", ColumnType.STRING),
                                     ("f_x", ColumnType.FLOAT)]).write_data(tuple(facts.values()))
    dims = {"d_id": list(range(0, 500)), "d_name": [f"name{int(i) % 7}" for i in range(500)], "d_w": [float(i % 13) for i in range(500)]}
    BlockFile(folder / "dims.bin", [("d_id", ColumnType.INTEGER), ("d_name", ColumnType.STRING), ("d_w", ColumnType.FLOAT)]).write_data(tuple(dims.values()))
    codes = {"k_code": [f"c{i:03d}" for i in range(0, 40)], "k_region": [f"r{i % 4}" for i in range(40)]}
    BlockFile(folder / "codes.bin", [("k_code", ColumnType.STRING), ("k_region", ColumnType.STRING)]).write_data(tuple(codes.values()))
    dup = {"u_id": [1, 2, 2, 3, 599], "u_tag": ["a", "b", "c", "d", "e"]}
    BlockFile(folder / "dup.bin", [("u_id", ColumnType.INTEGER), ("u_tag", ColumnType.STRING)]).write_data(tuple(dup.values()))
    return {k: str(folder / f"{k}.bin") for k in ("facts", "dims", "codes", "dup")}


def _star_queries(ns, t, e):
    C, F = ns.Col, ns.F

    def tbl(name, first=False):
        return (ns.DataFrame(e) if first else ns.DataFrame()).table(t[name])

    return {
        # not aggregated: rows in the probe side's order (reference: right-row major)
        "rows_int_key": (tbl("dims", True).join(tbl("facts"), on=C("d_id") == C("f_dim"), how="inner")
                         .filter(C("f_x") > 900).select(C("f_id"), C("d_name"), C("f_x"), (C("d_w") * C("f_x")).alias("wx")), True),
        "agg_int_key": (tbl("dims", True).join(tbl("facts"), on=C("d_id") == C("f_dim"), how="inner")
                        .filter(C("d_w") > 2.0).group_by(C("d_name")).agg(F.count(), F.sum(C("f_x") * C("d_w")).alias("s"), F.max(C("f_id")).alias("m")), False),
        "str_key": (tbl("codes", True).join(tbl("facts"), on=C("k_code") == C("f_code"), how="inner")
                    .group_by(C("k_region")).agg(F.count(), F.sum(C("f_x")).alias("s")), False),
        "no_build_columns": (tbl("dims", True).join(tbl("facts"), on=C("d_id") == C("f_dim"), how="inner").group_by(C("f_code")).agg(F.count()), False),
        "duplicate_build_keys": (tbl("dup", True).join(tbl("facts"), on=C("u_id") == C("f_dim"), how="inner")
                                 .group_by(C("u_tag")).agg(F.count(), F.sum(C("f_x")).alias("s")), False),
        "filtered_build_side": (tbl("dims", True).filter(C("d_id") < 100).select(C("d_id").alias("id2"), C("d_name"))
                                .join(tbl("facts").filter(C("f_x") < 500).select(C("f_dim"), (C("f_x") + 1).alias("x1")), on=C("id2") == C("f_dim"), how="inner")
                                .group_by(C("d_name")).agg(F.sum(C("x1")).alias("s"), F.count()), False),
        # three tables: (dims JOIN facts) JOIN codes -- the inner join result is the build side of the outer one here ...
        "three_tables": (tbl("dims", True).join(tbl("facts"), on=C("d_id") == C("f_dim"), how="inner")
                         .join(tbl("codes"), on=C("f_code") == C("k_code"), how="inner")
                         .group_by(C("k_region")).agg(F.count(), F.sum(C("f_x")).alias("s")), False),
        # ... and here a dimension joins a join's rows from the left
        "three_tables_rows": (tbl("codes", True).join(tbl("facts").filter(C("f_x") > 950), on=C("k_code") == C("f_code"), how="inner")
                              .select(C("f_id"), C("f_dim"), C("k_region"))
                              .join(tbl("dims").select(C("d_id"), C("d_name")), on=C("f_dim") == C("d_id"), how="inner"), False),
        # computed, negative keys (they still fit the compact 8-byte-slot table) ...
        "negative_keys": (tbl("dims", True).select((C("d_id") - 300).alias("k1"), C("d_name"))
                          .join(tbl("facts").select((C("f_dim") - 300).alias("k2"), C("f_x")), on=C("k1") == C("k2"), how="inner")
                          .group_by(C("d_name")).agg(F.count(), F.min(C("k2")).alias("lo")), False),
        # ... and keys that do not: FLOAT keys are f64 bit patterns (the build falls back to 16-byte slots)
        "wide_keys": (tbl("dims", True).select((C("d_id") * 0.5).alias("k1"), C("d_w"))
                      .join(tbl("facts").select((C("f_dim") * 0.5).alias("k2"), C("f_id")), on=C("k1") == C("k2"), how="inner")
                      .filter(C("f_id") < 2000).group_by(C("d_w")).agg(F.count(), F.max(C("f_id")).alias("m")), False),
    }


def test_timestamp_join_keys(engine, tables):
    """TIMESTAMP keys are 64-bit microsecond counts: the wide table format; every order date is unique."""
    ns = cases.namespace()

    def q(e):
        a = ns.DataFrame(e).table(tables["orders"]).select(ns.Col("order_date").alias("d1"), ns.Col("product"))
        b = ns.DataFrame().table(tables["orders"]).select(ns.Col("order_date").alias("d2"), ns.Col("price"))
        return a.join(b, on=ns.Col("d1") == ns.Col("d2"), how="inner").filter(ns.Col("price") > 40).select(ns.Col("product"), ns.Col("price"), ns.Col("d2"))

    O.assert_rows_equal(q(engine).collect(), O.run_task(q(None).task, wire=True))


@pytest.mark.parametrize("name", ["negative_keys", "wide_keys", "rows_int_key", "agg_int_key", "str_key", "no_build_columns", "duplicate_build_keys", "filtered_build_side",
                                  "three_tables", "three_tables_rows"])
def test_star_schema_joins(engine, star, name):
    ns = cases.namespace()
    df, ordered = _star_queries(ns, star, engine)[name]
    want = O.run_task(_star_queries(ns, star, None)[name][0].task, wire=True)
    assert len(want) > 0
    got = df.collect()
    O.assert_rows_equal(got, want, ordered=ordered and engine.fused_probe, rel=5e-7)
    if engine.fused_probe and name == "duplicate_build_keys":
        assert engine.last_stats["join"].startswith("build / probe"), engine.last_stats["join"]
    if engine.fused_probe and name in ("rows_int_key", "agg_int_key", "str_key", "no_build_columns"):
        assert engine.last_stats["join"].startswith("lookup fused"), engine.last_stats["join"]
        assert ("one scan over the build side" in engine.last_stats["join"]) == engine.fused_build
