"""Prepared passes enqueued ahead of the host (msc_prepared_enqueue / msc_prepared_wait): every pass is a complete query
whose result must be the oracle's, results come back in order, and the ring refuses more passes than it has room for."""

from __future__ import annotations

import pytest

import cases
from minispark_b200 import CudaExecutionEngine
from minispark_b200 import native as N
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu


def _rows(final, plan):
    names = [n for n, _ in plan.schema]
    keys = final.cols[0].dict.export()
    cols = [final.column_numpy(i) for i in range(len(names))]
    return {keys[int(cols[0][r])]: {n: cols[i][r].item() for i, n in enumerate(names) if i} for r in range(final.nrows)}


def test_pipelined_prepared_passes(small_lineitem):
    ns = cases.namespace()
    want = {r["l_returnflag"]: r for r in O.run_task(cases.q1(ns, small_lineitem).task, wire=False)}
    with CudaExecutionEngine() as engine:
        prepared = engine.prepare(cases.q1(ns, small_lineitem).task)
        assert not prepared.enqueue()  # the first passes set the prepared object up
        for _ in range(2):
            prepared.run()
            engine.release_query()
        depth = N.K["MSC_PREPARED_RING"]
        results = []
        for _ in range(3):  # three rounds of a full ring
            for _ in range(depth):
                assert prepared.enqueue()
            assert not prepared.enqueue()  # ring full
            for _ in range(depth):
                results.append(_rows(prepared.wait(), prepared.plan))
        assert len(results) == 3 * depth
        for rows in results:
            assert sorted(rows) == sorted(want)
            for k, ref in want.items():
                for name, v in rows[k].items():
                    assert abs(v - ref[name]) <= 1e-9 * max(abs(ref[name]), 1e-300), (k, name, v, ref[name])
        final, _ = prepared.run()  # the synchronous form still works afterwards
        assert _rows(final, prepared.plan).keys() == want.keys()
        assert engine.ctx.stats().last_scan_ms > 0
