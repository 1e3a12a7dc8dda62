"""The C restatements (oracle/q1_port.c, oracle/cfg_port.c) -- bench.py's full-size checkers and the reference arm -- are
pinned on CPU: against the fixture the REAL reference produced (tests/golden/q1_small.json) and against the Python oracle."""

from __future__ import annotations

import struct
import sys
from datetime import datetime
from pathlib import Path

import numpy as np
import pytest

import cases
from golden import golden_io
from oracle import ports
from oracle import py_oracle as O

GOLDEN = Path(__file__).parent / "golden"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "bench"))


def _f32(x: float) -> float:
    return struct.unpack("<f", struct.pack("<f", x))[0]


def _rows(port: dict, wire: bool) -> list[dict]:
    """q1_port output in the shape of the query's result rows (AVG = SUM / COUNT, plan.py:200-203)."""
    rnd = (lambda v: _f32(float(v))) if wire else float  # (%.17g prints whole sums without a decimal point)
    out = []
    for g in port["groups"]:
        n = g["count"]
        # wire mode: the final aggregate and the AVG projection share a stage (plan.py:199-203), so AVG divides the f64 sum of
        # the f32 partials and only the result row is narrowed (io.py:91-94)
        sq, sp, sd = float(g["sum_qty"]), float(g["sum_base_price"]), float(g["sum_disc"])
        out.append({"l_returnflag": g["key"], "sum_qty": rnd(sq), "sum_base_price": rnd(sp), "sum_disc_price": rnd(g["sum_disc_price"]),
                    "sum_charge": rnd(g["sum_charge"]), "avg_qty": rnd(sq / n), "avg_price": rnd(sp / n), "avg_disc": rnd(sd / n),
                    "count_order": n})
    return out


def test_q1_port_wire_matches_the_real_reference_fixture(small_lineitem):
    fixture = golden_io.load(GOLDEN / "q1_small.json")["q1_wire"]
    port = ports.q1(small_lineitem, threads=1, wire=1)  # one thread: partials merge in block order, like the reference's stage loop
    assert port["rows"] == sum(r["count_order"] for r in fixture)
    O.assert_rows_equal(_rows(port, wire=True), fixture)  # f32-exact


@pytest.mark.parametrize("threads", [1, 4])
def test_q1_port_f64_matches_python_oracle(small_lineitem, threads):
    want = O.run_task(cases.q1(cases.namespace(), small_lineitem).task, wire=False)
    got = _rows(ports.q1(small_lineitem, threads=threads, wire=0), wire=False)
    O.assert_rows_equal(got, want, rel=1e-12)


def test_q1_port_block_sample(small_lineitem):
    part = ports.q1(small_lineitem, threads=2, max_blocks=2, wire=0)
    assert part["blocks"] == 2 and part["rows"] == 8192


@pytest.fixture(scope="module")
def tpch_pair(tmp_path_factory):
    import gen_tpch

    folder = tmp_path_factory.mktemp("cfg")
    lineitem, orders = folder / "lineitem.bin", folder / "orders.bin"
    gen_tpch.write_table(lineitem, "lineitem", sf=0.002, rows_per_block=2048,
                         columns=["l_orderkey", "l_quantity", "l_extendedprice", "l_shipmode"])
    gen_tpch.write_table(orders, "orders", sf=0.002, rows_per_block=1024, columns=["o_orderkey", "o_orderdate", "o_orderpriority"])
    return str(lineitem), str(orders), folder


def test_highcard_port_matches_python_oracle(tpch_pair):
    lineitem, _, folder = tpch_pair
    ns = cases.namespace()
    df = ns.DataFrame(None).table(lineitem).group_by(ns.Col("l_orderkey")).agg(
        ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.avg(ns.Col("l_extendedprice")).alias("p"), ns.F.count().alias("n"))
    want = sorted(O.run_task(df.task, wire=False), key=lambda r: r["l_orderkey"])
    got = ports.highcard(lineitem, folder / "hc.bin")
    assert got["groups"] == len(want) and got["rows"] == sum(r["n"] for r in want)
    assert got["keys"].tolist() == [r["l_orderkey"] for r in want]
    assert got["count"].tolist() == [r["n"] for r in want]
    np.testing.assert_allclose(got["sum_q"], [r["q"] for r in want], rtol=1e-12)
    np.testing.assert_allclose(got["sum_p"] / got["count"], [r["p"] for r in want], rtol=1e-12)


@pytest.mark.parametrize(("lo", "hi", "needle"), [("1994-01-01", "1994-12-31", "AIR"), ("1992-01-01", "1998-12-31", "R"),
                                                 ("1995-03-01", "1995-03-02", "SHIP")])
def test_join_port_matches_python_oracle(tpch_pair, lo, hi, needle):
    lineitem, orders, _ = tpch_pair
    ns = cases.namespace()
    o = ns.DataFrame(None).table(orders).alias("o")
    l = ns.DataFrame().table(lineitem).alias("l")
    df = (o.join(l, on=ns.Col("o.o_orderkey") == ns.Col("l.l_orderkey"), how="inner")
          .filter(ns.Col("o.o_orderdate").between(lo, hi)).filter(ns.Col("l.l_shipmode").like(f"%{needle}%"))
          .group_by(ns.Col("o.o_orderpriority")).agg(ns.F.count().alias("n"), ns.F.sum(ns.Col("l.l_extendedprice")).alias("rev")))
    want = {r["o_orderpriority"]: r for r in O.run_task(df.task, wire=False)}
    us = lambda s: int(datetime.fromisoformat(s).timestamp() * 1_000_000)  # noqa: E731  (TZ=UTC, conftest)
    got = ports.join(orders, lineitem, us(lo), us(hi), needle)
    assert {g["key"] for g in got["groups"]} == set(want)
    for g in got["groups"]:
        assert g["count"] == want[g["key"]]["n"]
        assert abs(g["sum"] - want[g["key"]]["rev"]) <= 1e-12 * abs(g["sum"])
    nlines = O.run_task(ns.DataFrame(None).table(lineitem).group_by(ns.Col("l_shipmode")).agg(ns.F.count().alias("n")).task)
    assert got["pairs"] == sum(r["n"] for r in nlines)  # every lineitem row has exactly one order


@pytest.mark.parametrize("key", ["l_shipmode", "l_quantity", "l_orderkey"])
def test_groupby_port_matches_python_oracle(tpch_pair, key):
    lineitem, _, _ = tpch_pair
    ns = cases.namespace()
    want = {r[key]: r for r in O.run_task(cases.midcard_frame(ns, lineitem, key).task, wire=False)}
    got = ports.groupby(lineitem, key)
    assert got["rows"] == sum(r["n"] for r in want.values()) and {g["key"] for g in got["groups"]} == set(want)
    for g in got["groups"]:
        ref = want[g["key"]]
        assert g["count"] == ref["n"] and g["min_p"] == ref["min_p"] and g["max_p"] == ref["max_p"]
        for name in ("sum_q", "sum_p", "sum_pq"):
            assert abs(g[name] - ref[name]) <= 1e-12 * abs(ref[name]), (key, g["key"], name)
        assert abs(g["sum_p"] / g["count"] - ref["avg_p"]) <= 1e-12 * abs(ref["avg_p"])
