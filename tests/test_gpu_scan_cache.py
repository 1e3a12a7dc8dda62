"""Compiled scans are kept per plan node and rebound to the next run's rows (execution._CompiledScan): repeated one-shot
queries must give the same rows as the first run, also when the data behind the same plan changes shape (another literal
through a new task is a new plan; the same task over a reloaded table rebinds to new pointers)."""

from __future__ import annotations

import numpy as np
import pytest

import cases
from minispark_b200 import BlockFile, CudaExecutionEngine
from minispark_b200.constants import ColumnType
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu


def _tables(tmp_path, n=5000):
    rng = np.random.default_rng(n)
    left = tmp_path / "dim.bin"
    right = tmp_path / "fact.bin"
    BlockFile(left, [("d_id", ColumnType.INTEGER), ("d_name", ColumnType.STRING), ("d_w", ColumnType.FLOAT)]).write_data(
        (list(range(200)), [f"n{i % 7}" for i in range(200)], (rng.integers(0, 100, 200) / 4.0).tolist()))
    BlockFile(right, [("f_id", ColumnType.INTEGER), ("f_x", ColumnType.FLOAT), ("f_tag", ColumnType.STRING)]).write_data(
        (rng.integers(0, 260, n).tolist(), (rng.integers(0, 4000, n) / 8.0).tolist(), [("AIR", "RAIL", "SHIP", "REG AIR")[i] for i in rng.integers(0, 4, n)]))
    return str(left), str(right)


def _queries(ns, left, right, engine=None):
    def joined():  # (a frame is a builder: every query gets its own)
        d = ns.DataFrame(engine).table(left).alias("d")
        f = ns.DataFrame().table(right).alias("f")
        return d.join(f, on=ns.Col("d.d_id") == ns.Col("f.f_id"), how="inner")

    return {
        "join_agg": joined().filter(ns.Col("d.d_w") > 5.0).filter(ns.Col("f.f_tag").like("%AIR%")).group_by(ns.Col("d.d_name")).agg(
            ns.F.count().alias("n"), ns.F.sum(ns.Col("f.f_x")).alias("s")),
        # 7 groups x 7 accumulator cells behind a fused probe: the specialised kernel's shared-memory-cell form with the compacted tail
        "join_agg_wide": joined().filter(ns.Col("f.f_tag").like("%A%")).group_by(ns.Col("d.d_name")).agg(
            ns.F.count().alias("n"), ns.F.sum(ns.Col("f.f_x")).alias("sx"), ns.F.sum(ns.Col("d.d_w")).alias("sw"), ns.F.min(ns.Col("f.f_x")).alias("lo"),
            ns.F.max(ns.Col("f.f_x")).alias("hi"), ns.F.avg(ns.Col("d.d_w")).alias("aw"), ns.F.sum(ns.Col("f.f_x") * ns.Col("d.d_w")).alias("sxw")),
        "join_select": joined().filter(ns.Col("f.f_x") > 100.0).select(ns.Col("d.d_name"), ns.Col("f.f_x"), (ns.Col("d.d_w") * 2).alias("w2")),
        "hash_agg": ns.DataFrame(engine).table(right).group_by(ns.Col("f_id")).agg(ns.F.sum(ns.Col("f_x")).alias("s"), ns.F.count().alias("n")),
        "filter_project": ns.DataFrame(engine).table(right).filter(ns.Col("f_tag") == "SHIP").select(ns.Col("f_id"), (ns.Col("f_x") + 1).alias("x1")),
    }


def test_repeated_queries_rebind_their_compiled_scans(tmp_path):
    ns = cases.namespace()
    left, right = _tables(tmp_path)
    want = {name: O.run_task(q.task, wire=True) for name, q in _queries(ns, left, right).items()}
    with CudaExecutionEngine() as e:
        frames = _queries(ns, left, right, e)
        for name, q in frames.items():
            rebound = []
            for attempt in range(4):
                if attempt == 3:  # the tables leave the device and come back at other addresses
                    e.drop_table_cache()
                O.assert_rows_equal(q.collect(), want[name], ordered=name in ("join_select", "filter_project"), rel=5e-7)
                rebound.append(e.last_stats.get("scans_rebound", 0))
            assert rebound[0] == 0 and rebound[3] == 0, (name, rebound)
            assert max(rebound[1:3]) > 0 or e.last_stats.get("plan", "").startswith("prepared"), (name, rebound, e.last_stats.get("plan"))


def test_scan_cache_can_be_switched_off(tmp_path):
    ns = cases.namespace()
    left, right = _tables(tmp_path, 3000)
    with CudaExecutionEngine() as e:
        e.scan_cache_enabled = False
        for name, q in _queries(ns, left, right, e).items():
            for _ in range(2):
                got = q.collect()
                assert e.last_stats.get("scans_rebound", 0) == 0
            O.assert_rows_equal(got, O.run_task(_queries(ns, left, right)[name].task, wire=True), ordered=name in ("join_select", "filter_project"), rel=5e-7)
