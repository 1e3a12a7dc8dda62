"""Edge cases of the hot path on the GPU (through the C-ABI), checked against the oracle.

Row counts around the kernels' tile sizes (a regvm warp tile is 256 rows, column allocations are padded to
8192), empty and one-row tables, ragged multi-block files, filters that leave nothing (globally, per group,
per lane), strings of length 0 and 254 (the format's maximum, io.py:43), dictionaries wider than u8 / u16
codes.  The reference's own edge tests are tests/test_io.py:14-98 (block splitting) and the empty results
of tests/test_e2e.py; the rest pins what a from-scratch engine most easily gets wrong.
"""

from __future__ import annotations

import numpy as np
import pytest

import cases
from minispark_b200 import BlockFile, CudaExecutionEngine
from minispark_b200.constants import ColumnType
from oracle import py_oracle as O

pytestmark = pytest.mark.gpu
SCHEMA = [("k", ColumnType.STRING), ("i", ColumnType.INTEGER), ("x", ColumnType.FLOAT)]


def _write(path, nrows, rows_per_block=None, nkeys=3):
    rng = np.random.default_rng(nrows * 7 + nkeys)
    keys = [f"key{j}" for j in range(nkeys)]
    k = [keys[j] for j in rng.integers(0, nkeys, nrows)]
    i = rng.integers(-1000, 1000, nrows).tolist()
    x = (rng.integers(0, 4000, nrows) / 8.0).tolist()  # exact in f32
    bf = BlockFile(path, SCHEMA)
    if nrows == 0:
        bf.write_rows([])
        return str(path)
    if rows_per_block is None:
        bf.write_data((k, i, x))
        return str(path)
    # several ragged row-blocks (the package writer only splits at 2 Mi rows): serialise the blocks directly
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "bench"))
    import gen_tpch

    def blocks():
        for lo in range(0, nrows, rows_per_block):
            ks = k[lo:lo + rows_per_block]
            raw = [t.encode() for t in ks]
            yield len(ks), {"k": ([len(r) for r in raw], b"".join(raw)), "i": i[lo:lo + rows_per_block], "x": x[lo:lo + rows_per_block]}

    with open(path, "wb") as f:
        gen_tpch.write_stream(f, SCHEMA, blocks())
    assert len(BlockFile(path).block_starts) == -(-nrows // rows_per_block)
    return str(path)


def _queries(ns, table, engine=None):
    def df():
        return ns.DataFrame(engine).table(table)

    return {
        "dense": df().group_by(ns.Col("k")).agg(ns.F.sum(ns.Col("x")).alias("s"), ns.F.sum(ns.Col("x") * ns.Col("x")).alias("s2"), ns.F.count().alias("n"),
                                               ns.F.min(ns.Col("i")).alias("lo"), ns.F.max(ns.Col("x")).alias("hi")),
        "dense_sum_only": df().filter(ns.Col("i") >= 0).group_by(ns.Col("k")).agg(ns.F.sum(ns.Col("x") + 1).alias("s"), ns.F.avg(ns.Col("x")).alias("a")),
        "hash": df().group_by(ns.Col("i")).agg(ns.F.sum(ns.Col("x")).alias("s"), ns.F.count().alias("n")),
        "project": df().filter(ns.Col("x") > 250.0).select(ns.Col("k"), (ns.Col("i") * 2).alias("i2"), ns.Col("x")),
        "nothing_survives": df().filter(ns.Col("i") > 5000).group_by(ns.Col("k")).agg(ns.F.sum(ns.Col("x")).alias("s"), ns.F.count().alias("n")),
        "one_group_survives": df().filter(ns.Col("k") == "key1").group_by(ns.Col("k")).agg(ns.F.sum(ns.Col("x")).alias("s"), ns.F.count().alias("n")),
    }


# jit="always": every dense aggregate and filter / project scan below runs on a kernel compiled for its program (csrc/jit.cu)
JIT_MODES = pytest.mark.parametrize("jit", ["auto", "always"], ids=["interpreted", "specialised"])


@JIT_MODES
@pytest.mark.parametrize("nrows", [0, 1, 255, 256, 257, 8191, 8192, 8193, 20000])
def test_row_counts_around_tile_sizes(tmp_path, nrows, jit):
    ns = cases.namespace()
    table = _write(tmp_path / f"t{nrows}.bin", nrows)
    with CudaExecutionEngine(jit=jit) as e:
        for name, q in _queries(ns, table, e).items():
            got = q.collect()
            want = O.run_task(_queries(ns, table)[name].task, wire=True)
            O.assert_rows_equal(got, want, ordered=name == "project")


def test_ragged_blocks_and_many_groups(tmp_path):
    ns = cases.namespace()
    table = _write(tmp_path / "ragged.bin", 5000, rows_per_block=777, nkeys=4)
    with CudaExecutionEngine() as e:
        for name, q in _queries(ns, table, e).items():
            # several blocks: the reference rounds every block's partial sum to f32 before merging (tasks.py:373 ->
            # io.py:91-94), the GPU rounds once, hence the few-ulp tolerance on FLOAT sums
            O.assert_rows_equal(q.collect(), O.run_task(_queries(ns, table)[name].task, wire=True), ordered=name == "project",
                                rel=0.0 if name == "project" else 5e-7)


@JIT_MODES
@pytest.mark.parametrize("nkeys", [1, 2, 4, 5, 12, 300])
def test_group_counts_across_kernel_variants(tmp_path, nkeys, jit):
    """Interpreted: 1..4 groups take the masked regvm variants, 5 and 12 the generic one; specialised: up to 32 accumulator
    cells compile for 128 registers, up to 56 for 168, more go back to the interpreter; 300 keys: u16 codes, too many cells
    -> hash mode on the codes."""
    ns = cases.namespace()
    table = _write(tmp_path / f"g{nkeys}.bin", 3000, nkeys=nkeys)
    with CudaExecutionEngine(jit=jit) as e:
        for name in ("dense", "dense_sum_only", "one_group_survives"):
            got = _queries(ns, table, e)[name].collect()
            O.assert_rows_equal(got, O.run_task(_queries(ns, table)[name].task, wire=True))


def test_string_lengths_0_and_254_and_wide_dictionary(tmp_path):
    ns = cases.namespace()
    path = tmp_path / "s.bin"
    long_a, long_b = "a" * 254, "a" * 253 + "b"
    rows = [{"s": ["", long_a, long_b, "x"][j % 4], "u": f"u{j % 70000:05d}", "v": j % 7} for j in range(70001)]
    BlockFile(path).write_rows(rows)

    def q1(e=None):
        return ns.DataFrame(e).table(str(path)).group_by(ns.Col("s")).agg(ns.F.count().alias("n"), ns.F.sum(ns.Col("v")).alias("t"))

    def q2(e=None):  # 70000 distinct strings: u32 codes, hash aggregate on the codes
        return ns.DataFrame(e).table(str(path)).group_by(ns.Col("u")).agg(ns.F.count().alias("n"))

    def q3(e=None):
        return ns.DataFrame(e).table(str(path)).filter(ns.Col("v") == 3).filter(ns.Col("s").like("a%b")).select(ns.Col("s"), ns.Col("u"))

    with CudaExecutionEngine() as e:
        O.assert_rows_equal(q1(e).collect(), O.run_task(q1().task, wire=True))
        O.assert_rows_equal(q2(e).collect(), O.run_task(q2().task, wire=True))
        O.assert_rows_equal(q3(e).collect(), O.run_task(q3().task, wire=True), ordered=True)


def test_result_blockfile_spans_several_blocks(tmp_path, monkeypatch):
    """A result larger than a row-block is written as several blocks (reference io.py:74-109 splits at ROWS_PER_BLOCK; its
    own test patches the constant, tests/test_io.py) and reads back through the reference-format reader, row for row."""
    import minispark_b200.io as mio

    monkeypatch.setattr(mio, "ROWS_PER_BLOCK", 1000)
    table = _write(tmp_path / "t.bin", 4321)
    ns = cases.namespace()
    with CudaExecutionEngine() as engine:
        df = ns.DataFrame(engine).table(table).filter(ns.Col("i") > -900).select(ns.Col("k"), ns.Col("i"), (ns.Col("x") * 2).alias("x2"))
        jobs = engine.execute_full_task(df.task)
        files = [f.file_path for j in jobs for f in j.output_files]
        assert len(files) == 1
        bf = BlockFile(files[0])
        want = O.run_task(ns.DataFrame(None).table(table).filter(ns.Col("i") > -900).select(ns.Col("k"), ns.Col("i"), (ns.Col("x") * 2).alias("x2")).task, wire=True)
        assert len(bf.block_starts) == -(-len(want) // 1000) and len(bf.block_starts) >= 4
        O.assert_rows_equal(list(engine.collect_results(jobs)), want, ordered=True)


def test_select_star_over_a_wide_join_is_split_into_passes(tmp_path):
    """More output columns than one scan binds (MSC_VM_MAX_OUT = 24): `SELECT *` over the join of two 14-column tables runs
    as several passes over the same rows whose columns are stitched together (the reference has no such limit)."""
    ncols = 14
    rng = np.random.default_rng(11)
    n = 3000

    def table(path, prefix, keys):
        schema = [(f"{prefix}{c}", ColumnType.INTEGER if c % 3 else ColumnType.FLOAT) for c in range(ncols)]
        schema[1] = (f"{prefix}1", ColumnType.STRING)
        cols = []
        for c, (_, t) in enumerate(schema):
            if c == 0:
                cols.append(keys)
            elif t == ColumnType.STRING:
                cols.append([f"s{int(v)}" for v in rng.integers(0, 9, len(keys))])
            elif t == ColumnType.FLOAT:
                cols.append((rng.integers(0, 800, len(keys)) / 8.0).tolist())
            else:
                cols.append(rng.integers(-99, 99, len(keys)).tolist())
        schema[0] = (f"{prefix}0", ColumnType.INTEGER)
        cols[0] = list(keys)
        BlockFile(path, schema).write_data(tuple(cols))
        return str(path)

    a = table(tmp_path / "a.bin", "a", list(range(200)))
    b = table(tmp_path / "b.bin", "b", rng.integers(0, 260, n).tolist())
    ns = cases.namespace()

    def q(e):
        return ns.DataFrame(e).table(a).join(ns.DataFrame().table(b), on=ns.Col("a0") == ns.Col("b0"), how="inner").filter(ns.Col("b2") > -50)

    want = O.run_task(q(None).task, wire=True)
    assert len(want[0]) == 2 * ncols and len(want) > 100
    for fused in (True, False):
        with CudaExecutionEngine() as engine:
            engine.fused_probe = fused
            O.assert_rows_equal(q(engine).collect(), want)
