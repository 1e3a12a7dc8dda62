"""The oracle is pinned against the reference's own golden vectors (CPU only)."""

from __future__ import annotations

import json
from pathlib import Path

import pytest

import cases
from golden import golden_io
from minispark_b200.parser import parse_sql
from oracle import py_oracle as O

GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize(("name", "sql", "expected"), cases.SQL_CASES, ids=[c[0] for c in cases.SQL_CASES])
def test_oracle_matches_reference_e2e_vectors(tables, name, sql, expected):
    df = parse_sql(sql.format(**tables))
    got = O.run_task(df.task, wire=True)
    O.assert_rows_equal(got, expected, ordered=name in cases.ORDERED_SQL)


@pytest.mark.parametrize("name", [n for n, c in cases.DF_CASES.items() if c[1] is not None])
def test_oracle_matches_reference_dataframe_vectors(tables, name):
    build, expected, ordered = cases.DF_CASES[name]
    got = O.run_task(build(cases.namespace(), tables, None).task, wire=True)
    O.assert_rows_equal(got, expected, ordered=ordered)


@pytest.mark.parametrize("name", sorted(cases.DF_CASES))
def test_oracle_matches_fixtures_generated_by_the_real_reference(tables, name):
    """tests/golden/df_cases.json was produced by the real PythonExecutionEngine (make_golden.py)."""
    fixture = golden_io.load(GOLDEN / "df_cases.json")
    if name not in fixture:
        pytest.skip("case not runnable on the reference (engine-specific)")
    build, _, ordered = cases.DF_CASES[name]
    got = O.run_task(build(cases.namespace(), tables, None).task, wire=True)
    O.assert_rows_equal(got, fixture[name], ordered=ordered)


def test_oracle_q1_matches_reference_fixture(small_lineitem):
    fixture = golden_io.load(GOLDEN / "q1_small.json")
    got = O.run_task(cases.q1(cases.namespace(), small_lineitem).task, wire=True)
    O.assert_rows_equal(got, fixture["q1_wire"])
    f64 = O.run_task(cases.q1(cases.namespace(), small_lineitem).task, wire=False)
    O.assert_rows_equal(f64, fixture["q1_wire"], rel=2e-6)  # f32 quantisation of the wire result
