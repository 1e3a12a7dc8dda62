"""World-size-2 tests of the multi-GPU host plumbing under gloo on CPU (no GPU needed)."""

from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minispark_b200.distributed import Comm, exchange_plan, range_bounds, shard_blocks, unify_keys


def test_shard_blocks_cover_every_block_once():
    for world in (1, 2, 3, 8):
        seen = sorted(b for r in range(world) for b in shard_blocks(43, r, world))
        assert seen == list(range(43))
    assert shard_blocks(5, 1, 2) == [2, 3, 4] and shard_blocks(5, 0, 2) == [0, 1]  # contiguous ranges: rank order = row order
    assert [len(shard_blocks(43, r, 8)) for r in range(8)] == [5, 5, 6, 5, 5, 6, 5, 6]
    with pytest.raises(ValueError):
        shard_blocks(5, 2, 2)


def test_unify_keys_and_exchange_plan():
    universe, maps = unify_keys([["N", "R"], ["A", "N"], []])
    assert universe == ["A", "N", "R"]
    assert maps == [[1, 2], [0, 1], []]
    send, recv = exchange_plan([[1, 2], [3, 4]], rank=1)
    assert send == [3, 4] and recv == [2, 4]


def test_range_bounds_for_sorted_partial_results():
    """(sorted?, rows, first key, last key) per rank -> the key range every rank receives, or None (hash instead)."""
    # a clustered table sharded in row order: rank r keeps its keys; key 50 straddles ranks 0 and 1 and ends up on rank 1
    assert range_bounds([(True, 10, 1, 50), (True, 10, 50, 90), (True, 0, 0, 0), (True, 5, 95, 99)]) == [1, 50, 95, 95]
    assert range_bounds([(True, 10, 1, 60), (True, 10, 50, 90)]) is None      # ranges interleave
    assert range_bounds([(True, 10, 1, 40), (False, 10, 50, 90)]) is None     # a rank whose rows do not ascend
    assert range_bounds([(True, 0, 0, 0), (True, 0, 0, 0)]) == [(1 << 63) - 1] * 2
    assert range_bounds([(True, 3, -9, -2)]) == [-9]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = Comm.from_env()
        assert (comm.rank, comm.world) == (rank, world)
        # dictionary unification of a low-cardinality group key
        local = [["N", "R"], ["A", "N", "R"]][rank]
        universe, maps = unify_keys(comm.all_gather_object(local))
        assert universe == ["A", "N", "R"] and [universe[i] for i in maps[rank]] == local
        assert comm.max_int(10 * (rank + 1)) == 20
        # ragged all-gather of partial aggregate tables
        n = 2 + rank
        keys = torch.tensor(maps[rank], dtype=torch.int64)
        sums = torch.arange(n, dtype=torch.float64) + 100 * rank
        (gk, gs), counts = comm.all_gather_rows([keys, sums], n)
        assert counts == [2, 3]
        assert gk.tolist() == [1, 2, 0, 1, 2] and gs.tolist() == [0.0, 1.0, 100.0, 101.0, 102.0]
        # all-to-all of partition-contiguous rows: rank r sends (r+1) rows to rank 0 and 2 rows to rank 1
        send_counts = [rank + 1, 2]
        payload = torch.arange(sum(send_counts), dtype=torch.int64) + 1000 * rank
        (got,), recv = comm.all_to_all_rows([payload], send_counts)
        if rank == 0:
            assert recv == [1, 2] and got.tolist() == [0, 1000, 1001]
        else:
            assert recv == [2, 2] and got.tolist() == [1, 2, 1002, 1003]
        comm.barrier()
    finally:
        dist.destroy_process_group()


def test_comm_world_size_2_gloo():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_invert_code_maps_for_the_in_kernel_merge():
    """Host side of msc_dense_fused_peer: per-rank local -> merged code maps, inverted into the [world][32] table the
    scan kernel's last CTA indexes by merged group."""
    from minispark_b200.distributed import invert_code_maps, unify_keys

    universe, maps = unify_keys([["A", "N"], ["N", "R"], []])
    assert universe == ["A", "N", "R"] and maps == [[0, 1], [1, 2], []]
    padded = [m + [-1] * (2 - len(m)) for m in maps]         # every rank's slot has room for gmax = 2 groups
    inv = invert_code_maps(padded)
    assert len(inv) == 3 * 32
    assert inv[0:3] == [0, 1, -1]                               # rank 0: A -> local 0, N -> local 1, no R
    assert inv[32:35] == [-1, 0, 1]                             # rank 1: no A, N -> local 0, R -> local 1
    assert inv[64:67] == [-1, -1, -1]                           # rank 2 holds no group at all
    import pytest

    with pytest.raises(ValueError):
        invert_code_maps([[40]])
