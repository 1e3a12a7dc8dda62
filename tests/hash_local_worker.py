"""Worker of test_gpu_hash_local.py: hash-mode GROUP BY on the GPU against numpy, in a process of its own so that the
library's environment switches (MSC_HASH_LOCAL_SLOTS, MSC_HASH_OPTIMISTIC) can be set per run.

    python tests/hash_local_worker.py <folder> <nrows> <distinct> [expect_attempts]
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))

import cases  # noqa: E402
from minispark_b200 import BlockFile, CudaExecutionEngine  # noqa: E402
from minispark_b200.constants import ColumnType  # noqa: E402


def main() -> None:
    folder, nrows, distinct = Path(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    expect_attempts = int(sys.argv[4]) if len(sys.argv) > 4 else None
    rng = np.random.default_rng(nrows + distinct)
    # skewed keys: a few hot groups and a long tail, negative values too; a float key column with the same cardinality
    k = np.where(rng.random(nrows) < 0.5, rng.zipf(1.3, nrows) % distinct, rng.integers(0, distinct, nrows)).astype(np.int64) - distinct // 3
    f = (k % 97).astype(np.float64) * 0.25 - 3.0
    i = rng.integers(-10**6, 10**6, nrows)
    x = rng.integers(0, 4000, nrows) / 8.0  # exact in f32, and every partial sum is exact in f64: any order gives the same bits
    table = folder / f"h_{nrows}_{distinct}.bin"
    BlockFile(table, [("k", ColumnType.INTEGER), ("f", ColumnType.FLOAT), ("i", ColumnType.INTEGER), ("x", ColumnType.FLOAT)]).write_data(
        (k.tolist(), f.tolist(), i.tolist(), x.tolist()))
    ns = cases.namespace()

    def expected(keys: np.ndarray) -> dict:
        order = np.argsort(keys, kind="stable")
        ks, xs, is_ = keys[order], x[order], i[order]
        uniq, start = np.unique(ks, return_index=True)
        return {key: (s, n, lo, hi, hx) for key, s, n, lo, hi, hx in zip(
            uniq.tolist(), np.add.reduceat(xs, start).tolist(), np.diff(np.append(start, len(ks))).tolist(),
            np.minimum.reduceat(is_, start).tolist(), np.maximum.reduceat(is_, start).tolist(), np.maximum.reduceat(xs, start).tolist())}

    with CudaExecutionEngine() as e:
        for name, key_expr, keys in (("integer key", ns.Col("k"), k), ("float key", ns.Col("f"), f),
                                     ("expression key", (ns.Col("k") % 13).alias("m"), np.mod(k, 13))):
            q = ns.DataFrame(e).table(str(table)).group_by(key_expr).agg(
                ns.F.sum(ns.Col("x")).alias("s"), ns.F.count().alias("n"), ns.F.min(ns.Col("i")).alias("lo"), ns.F.max(ns.Col("i")).alias("hi"),
                ns.F.max(ns.Col("x")).alias("hx"), ns.F.avg(ns.Col("x")).alias("a"))
            got = q.collect()
            want = expected(keys)
            assert e.last_stats["agg_mode"] == "hash", e.last_stats
            assert len(got) == len(want), (name, len(got), len(want))
            for row in got:
                row = tuple(row.values()) if isinstance(row, dict) else tuple(row)  # (collect() gives {column: value} rows, dataframe.py:71-79)
                s, n, lo, hi, hx = want[row[0]]
                # FLOAT results travel as f32 (the reference's wire and file format, io.py:91-94): the exact f64 sum, rounded once
                assert row[1:6] == (float(np.float32(s)), n, lo, hi, hx), (name, row, want[row[0]])
                assert row[6] == float(np.float32(s / n)), (name, row, s / n)
            if name == "integer key" and expect_attempts is not None:
                assert e.last_stats["hash_attempts"] == expect_attempts, e.last_stats
            print(name, "ok:", len(got), "groups, local slots", e.last_stats["hash_local_slots"], "attempts", e.last_stats["hash_attempts"])
    print("OK")


if __name__ == "__main__":
    main()
