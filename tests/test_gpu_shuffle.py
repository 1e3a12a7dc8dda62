"""The shuffle's building blocks on ONE GPU, through the C-ABI, against numpy (bit-exact):

* msc_partition (reference: WriteToShufflePartitions.write, tasks.py:347-375; zig fill_buckets, task_utils.zig:53-98):
  every row lands in the partition its key hashes to, partitions are contiguous, counts are exact and the order inside a
  partition is the input order (the reference appends to each bucket in row order);
* the peer-memory exchange with a world of one rank (begin / finish through the rank's own slot), including a second
  exchange into the same slot and the all-rows-to-all-ranks mode;
* the small table all-gather.
Two-rank runs of the same code are in tests/test_gpu_multi.py.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from minispark_b200 import native as N
from minispark_b200.distributed import Comm, PeerShuffle

pytestmark = pytest.mark.gpu


def _mix64(k: np.ndarray) -> np.ndarray:
    """murmur3 fmix64 (csrc/common.cuh msc_mix64), on the key's 64-bit pattern."""
    k = k.astype(np.uint64)
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xFF51AFD7ED558CCD)
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xC4CEB9FE1A85EC53)
        k ^= k >> np.uint64(33)
    return k


def _parts(keys: np.ndarray, nparts: int) -> np.ndarray:
    return ((_mix64(keys.astype(np.int64).view(np.uint64)) >> np.uint64(32)) % np.uint64(nparts)).astype(np.int64)


def _upload(ctx: N.Context, arrays: list[np.ndarray], physes: list[int]) -> int:
    out = C.c_void_p()
    ctx.call("msc_rel_alloc", len(arrays[0]), N.int32_array(physes), len(physes), C.byref(out))
    binds = (N.ColBind * len(physes))()
    ctx.check(ctx.lib.msc_rel_cols(out, binds, len(physes)))
    for b, a in zip(binds, arrays):
        if a.nbytes:
            ctx.call("msc_memcpy_h2d", C.c_void_p(b.data), a.ctypes.data_as(C.c_void_p), a.nbytes)
    return out.value


def _download(ctx: N.Context, rel: int, dtypes: list[str]) -> list[np.ndarray]:
    nrows, ncols = C.c_uint64(), C.c_int32()
    ctx.check(ctx.lib.msc_rel_info(C.c_void_p(rel), C.byref(nrows), C.byref(ncols)))
    assert ncols.value == len(dtypes)
    binds = (N.ColBind * len(dtypes))()
    ctx.check(ctx.lib.msc_rel_cols(C.c_void_p(rel), binds, len(dtypes)))
    out = []
    for b, dt in zip(binds, dtypes):
        a = np.zeros(nrows.value, dtype=dt)
        if a.nbytes:
            ctx.call("msc_memcpy_d2h", a.ctypes.data_as(C.c_void_p), C.c_void_p(b.data), a.nbytes)
        out.append(a)
    return out


@pytest.fixture(scope="module")
def ctx():
    c = N.Context(0)
    yield c
    c.close()


def _table(n: int, seed: int, key_kind: str):
    rng = np.random.default_rng(seed)
    if key_kind == "i64":
        key, phys, dt = rng.integers(-2**40, 2**40, n, dtype=np.int64), N.P_I64, "<i8"
    elif key_kind == "i32":
        key, phys, dt = rng.integers(-50, 50, n).astype(np.int32), N.P_I32, "<i4"   # few distinct keys: long partitions
    else:
        key, phys, dt = rng.integers(0, 200, n).astype(np.uint8), N.P_U8, "<u1"
    cols = [key, np.arange(n, dtype=np.int64), rng.random(n), rng.integers(0, 2**31, n).astype(np.uint32), rng.integers(0, 255, n).astype(np.uint8)]
    return cols, [phys, N.P_I64, N.P_F64, N.P_U32, N.P_U8], [dt, "<i8", "<f8", "<u4", "<u1"]


@pytest.mark.parametrize("nparts", [1, 2, 8, 16, 64])
@pytest.mark.parametrize(("nrows", "key_kind"), [(0, "i64"), (1, "i64"), (255, "i32"), (4096, "i64"), (4097, "u8"), (100_003, "i64"), (1_000_000, "i32")])
def test_partition_is_exact_and_stable(ctx, nparts, nrows, key_kind):
    cols, physes, dtypes = _table(nrows, 17 * nparts + nrows, key_kind)
    rel = _upload(ctx, cols, physes)
    counts = (C.c_uint64 * nparts)()
    out = C.c_void_p()
    ctx.call("msc_partition", C.c_void_p(rel), 0, nparts, counts, C.byref(out))
    got = _download(ctx, out.value, dtypes)
    part = _parts(cols[0], nparts)
    want_counts = np.bincount(part, minlength=nparts)
    assert [int(c) for c in counts] == want_counts.tolist()
    order = np.argsort(part, kind="stable")        # partition-contiguous, input order inside each partition
    for g, c in zip(got, cols):
        assert np.array_equal(g, c[order])
    ctx.lib.msc_rel_free(out)
    ctx.lib.msc_rel_free(C.c_void_p(rel))


@pytest.mark.parametrize("nparts", [1, 2, 8])
@pytest.mark.parametrize("sorted_keys", [True, False])
def test_partition_by_key_range(ctx, nparts, sorted_keys):
    """msc_partition_range: partition p = keys in [lower_bounds[p], lower_bounds[p + 1]); stable, so a sorted input stays
    sorted inside every partition (what the final aggregate after a shuffle of sorted partial results relies on)."""
    cols, physes, dtypes = _table(50_001, 3 * nparts, "i64")
    if sorted_keys:
        order = np.argsort(cols[0], kind="stable")
        cols = [c[order] for c in cols]
    qs = np.quantile(cols[0], [i / nparts for i in range(nparts)]).astype(np.int64)
    qs[0] = 12345  # (entry 0 is ignored: the first partition has no lower bound)
    if nparts > 2:
        qs[2] = qs[1]  # an empty range
    rel = _upload(ctx, cols, physes)
    counts = (C.c_uint64 * nparts)()
    out = C.c_void_p()
    ctx.call("msc_partition_range", C.c_void_p(rel), 0, nparts, (C.c_int64 * nparts)(*qs.tolist()), counts, C.byref(out))
    got = _download(ctx, out.value, dtypes)
    part = np.zeros(len(cols[0]), dtype=np.int64)
    for p in range(1, nparts):
        part += cols[0] >= qs[p]
    assert [int(c) for c in counts] == np.bincount(part, minlength=nparts).tolist()
    order = np.argsort(part, kind="stable")
    for g, c in zip(got, cols):
        assert np.array_equal(g, c[order])
    if sorted_keys:
        assert np.all(np.diff(got[0]) >= 0)
    ctx.lib.msc_rel_free(out)
    ctx.lib.msc_rel_free(C.c_void_p(rel))
    with pytest.raises(N.NativeError):  # bounds must not decrease
        ctx.call("msc_partition_range", C.c_void_p(0), 0, 3, (C.c_int64 * 3)(0, 5, 4), counts, C.byref(out))


def test_partition_rejects_bad_arguments(ctx):
    cols, physes, _ = _table(10, 1, "i64")
    rel = _upload(ctx, cols, physes)
    counts = (C.c_uint64 * 65)()
    out = C.c_void_p()
    for key_col, nparts in ((0, 0), (0, 65), (5, 2), (-1, 2)):
        with pytest.raises(N.NativeError):
            ctx.call("msc_partition", C.c_void_p(rel), key_col, nparts, counts, C.byref(out))
    ctx.lib.msc_rel_free(C.c_void_p(rel))


def test_exchange_through_the_ranks_own_slot(ctx):
    """world = 1: the whole begin / finish protocol (matrix through the control block, push kernel into the slot, flags)
    runs against the rank's own memory; rows must come back unchanged, a later exchange reuses the slot."""
    sh = PeerShuffle(ctx, Comm())
    assert sh.ok
    for nrows in (5000, 0, 70_000, 3):
        cols, physes, dtypes = _table(nrows, nrows + 5, "i64")
        rel = _upload(ctx, cols, physes)
        got_h, n = sh.exchange(rel, 0)
        assert n == nrows and sh.last_matrix == [[nrows]]
        ms = C.c_double()
        ctx.check(ctx.lib.msc_shuffle_wait(sh.handle, C.byref(ms)))
        for g, c in zip(_download(ctx, got_h, dtypes), cols):
            assert np.array_equal(g, c)
        ctx.lib.msc_rel_free(C.c_void_p(got_h))
        got_h, n = sh.exchange(rel, None)  # all rows to all ranks (a second slot: the first is still in use)
        assert n == nrows and sh.busy == {0, 1}
        for g, c in zip(_download(ctx, got_h, dtypes), cols):
            assert np.array_equal(g, c)
        ctx.lib.msc_rel_free(C.c_void_p(got_h))
        ctx.lib.msc_rel_free(C.c_void_p(rel))
        sh.release_all()
    assert set(sh.slot_bytes) == {0, 1}
    # the table all-gather
    src = np.arange(96, dtype=np.int64)
    a, b = C.c_void_p(), C.c_void_p()
    ctx.call("msc_dev_alloc", src.nbytes, C.byref(a))
    ctx.call("msc_dev_alloc", src.nbytes, C.byref(b))
    ctx.call("msc_memcpy_h2d", a, src.ctypes.data_as(C.c_void_p), src.nbytes)
    for _ in range(3):
        sh.allgather_table(a.value, src.nbytes, b.value)
    back = np.zeros_like(src)
    ctx.call("msc_memcpy_d2h", back.ctypes.data_as(C.c_void_p), b, src.nbytes)
    assert np.array_equal(back, src)
    with pytest.raises(N.NativeError):
        sh.allgather_table(a.value, N.K["MSC_SHUFFLE_TABLE_BYTES"] + 8, b.value)
    ctx.call("msc_dev_free", a)
    ctx.call("msc_dev_free", b)
    sh.close()
