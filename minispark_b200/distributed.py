"""Multi-GPU plumbing: one process (rank) per GPU.

The reference's only parallelism is data-parallel row-blocks plus a hash-partitioned shuffle through
files (``src/mini_spark/plan.py:90-109``: one ``ScanJob`` per block, partitions re-read per
``LoadShuffleFilesJob``; ``tasks.py:347-375`` writes ``hash(key) % 10`` buckets).  Here:

* row-blocks are dealt to ranks as contiguous ranges (:func:`shard_blocks`), so that the rank-ordered union of
  rank-local filter / project results is the table's input order (tests/test_execution.py:40-46 of the reference);
* a low-cardinality GROUP BY merges tiny per-rank partial tables: string keys are unified through
  their dictionary *entries* (:func:`unify_keys`); no row-level data crosses the fabric;
* a high-cardinality GROUP BY / JOIN routes rows by ``hash(key) % world`` (``msc_partition``) and exchanges them
  over NVLink peer memory (:class:`PeerShuffle`: the library's push kernel stores every column segment straight into
  the receiving rank's buffer, csrc/shuffle.cu) -- the GPU analogue of the reference's shuffle files.  Where CUDA IPC
  is not available the same rows travel through ONE group of NCCL sends / receives (:meth:`Comm.all_to_all_rows`).

``torch.distributed`` carries the rendezvous and the small host-side metadata (IPC handles, dictionary entries, row
counts under gloo); the host logic runs under ``gloo`` on CPU (tests) and under ``nccl`` on GPUs unchanged.
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Optional, Sequence


def shard_blocks(nblocks: int, rank: int, world: int) -> list[int]:
    """Row-blocks owned by ``rank``: one job per block as in the reference (plan.py:90-93), dealt as balanced CONTIGUOUS
    ranges -- rank r owns blocks [r * n // world, (r + 1) * n // world) -- so that concatenating rank-local results in rank
    order keeps the table's row order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank * nblocks // world, (rank + 1) * nblocks // world))


def unify_keys(per_rank_keys: Sequence[Sequence[str]]) -> tuple[list[str], list[list[int]]]:
    """Global key list (sorted, unique) and, per rank, the map local code -> global code.

    Works on dictionary entries (one string per *group*, not per row)."""
    universe = sorted({k for keys in per_rank_keys for k in keys})
    index = {k: i for i, k in enumerate(universe)}
    return universe, [[index[k] for k in keys] for keys in per_rank_keys]


def invert_code_maps(maps: Sequence[Sequence[int]], width: int = 32) -> list[int]:
    """For the in-kernel merge of partial aggregate tables (msc_dense_fused_peer): ``maps[r][g]`` is the merged group of
    rank r's local group g (unify_keys); the kernel wants the opposite direction as a flat [world][width] table --
    entry r * width + G = rank r's local group of merged group G, or -1 when rank r never saw it."""
    inv = [-1] * (len(maps) * width)
    for r, m in enumerate(maps):
        for g, merged in enumerate(m):
            if merged < 0:
                continue
            if merged >= width:
                raise ValueError(f"merged group {merged} does not fit the {width}-wide exchange table")
            inv[r * width + merged] = g
    return inv


def range_bounds(per_rank: Sequence[tuple[bool, int, int, int]]) -> Optional[list[int]]:
    """Key ranges for a shuffle of SORTED partial results.  per_rank[r] = (sorted?, rows, first key, last key) of rank r's
    rows.  When every rank's keys ascend and the ranks' ranges follow each other (last[r] <= first[r + 1], as for a table
    clustered by the key and sharded in row order), rank r keeps the keys from its own first key on: only a key that
    straddles two ranks moves, and what a rank receives is still sorted.  Returns lower_bounds[world] (entry 0 unused), or
    None when the ranges interleave -- then hashing spreads the keys evenly."""
    world = len(per_rank)
    if not all(p[0] for p in per_rank):
        return None
    nonempty = [p for p in per_rank if p[1] > 0]
    for a, b in zip(nonempty, nonempty[1:]):
        if a[3] > b[2]:
            return None
    bounds = [0] * world
    nxt = None  # an empty rank owns an empty range: its bound is the next non-empty rank's
    for r in range(world - 1, -1, -1):
        if per_rank[r][1] > 0:
            nxt = per_rank[r][2]
        bounds[r] = nxt if nxt is not None else (1 << 63) - 1
    return bounds


def boundary_plan(per_rank: Sequence[tuple[int, int, int]], bounds: Sequence[int], rank: int) -> tuple[bool, list[int]]:
    """The shuffle of sorted, rank-ordered partial aggregates (range_bounds) without moving rows: per_rank[r] = (rows, first
    key, last key) of rank r's partial result, whose keys are unique on the rank.  A group can only be shared by a rank's
    LAST row and the first row of the rank that owns the key (the last rank whose range starts at or below it).  Returns for
    `rank`: (its last row belongs to a later rank: drop it, the ranks whose last row must be folded into its first row --
    in rank order)."""
    world = len(per_rank)

    def owner(key: int) -> int:
        r = 0
        for i in range(1, world):
            if key >= bounds[i]:
                r = i
        return r

    moving = [per_rank[s][0] > 0 and owner(per_rank[s][2]) != s for s in range(world)]
    incoming = [s for s in range(world) if moving[s] and owner(per_rank[s][2]) == rank]
    return moving[rank], incoming


def exchange_plan(counts_matrix: Sequence[Sequence[int]], rank: int) -> tuple[list[int], list[int]]:
    """(send_counts, recv_counts) of ``rank`` given counts_matrix[src][dst] rows routed src -> dst."""
    world = len(counts_matrix)
    send = [int(counts_matrix[rank][dst]) for dst in range(world)]
    recv = [int(counts_matrix[src][rank]) for src in range(world)]
    return send, recv


@dataclass
class Comm:
    """Thin wrapper over ``torch.distributed`` (or a single-rank no-op)."""

    rank: int = 0
    world: int = 1
    device: Optional[Any] = None  # torch.device for collective buffers

    @classmethod
    def from_env(cls, device_index: Optional[int] = None) -> "Comm":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return cls()
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("WORLD_SIZE > 1 but torch.distributed is not initialised (call init_process_group first)")
        device = torch.device("cuda", device_index) if dist.get_backend() == "nccl" else torch.device("cpu")
        return cls(dist.get_rank(), dist.get_world_size(), device)

    # -- small helpers ------------------------------------------------------------------------------
    def all_gather_object(self, obj: Any) -> list[Any]:
        if self.world == 1:
            return [obj]
        import torch.distributed as dist

        out: list[Any] = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def max_int(self, value: int) -> int:
        return max(self.all_gather_object(int(value)))

    def barrier(self) -> None:
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()

    def all_gather_counts(self, values: Sequence[int]) -> list[list[int]]:
        """[rank][i] = ``values[i]`` of that rank: a tensor all-gather (on the device under NCCL: no pickling)."""
        if self.world == 1:
            return [[int(v) for v in values]]
        import torch
        import torch.distributed as dist

        mine = torch.tensor([int(v) for v in values], dtype=torch.int64, device=self.device)
        out = torch.empty(self.world * len(values), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, mine)
        return out.view(self.world, len(values)).cpu().tolist()

    def all_gather_rows(self, columns: Sequence[Any], nrows: int) -> tuple[list[Any], list[int]]:
        """All-gather a small relation given as 1-D tensors of ``nrows`` elements each (rank order).

        Returns (concatenated columns, rows per rank)."""
        counts = [c[0] for c in self.all_gather_counts([nrows])]
        if self.world == 1:
            return [c[:nrows].clone() for c in columns], counts
        out, _ = self.all_to_all_rows([c[:nrows].repeat(self.world) for c in columns], [nrows] * self.world, matrix=[[n] * self.world for n in counts])
        return out, counts

    def all_to_all_rows(self, columns: Sequence[Any], send_counts: Sequence[int], pad_rows: int = 0,
                        matrix: Optional[Sequence[Sequence[int]]] = None) -> tuple[list[Any], list[int]]:
        """Exchange partition-contiguous rows: the first ``send_counts[0]`` rows of every column go to
        rank 0, the next ``send_counts[1]`` to rank 1, ...  Returns (received columns, recv_counts); rows arrive
        ordered by sending rank.  All columns and peers travel in ONE group of point-to-point operations (one NCCL
        launch).  ``pad_rows``: zero rows appended to every received column (tile padding of the scan kernels)."""
        import torch

        if self.world == 1:
            return [c[:send_counts[0]].clone() for c in columns], [int(send_counts[0])]
        import torch.distributed as dist

        if matrix is None:
            matrix = self.all_gather_counts(send_counts)
        send, recv = exchange_plan(matrix, self.rank)
        total = sum(recv)
        send_off = [sum(send[:d]) for d in range(self.world)]
        recv_off = [sum(recv[:s]) for s in range(self.world)]
        out, ops = [], []
        for col in columns:
            dst = torch.zeros(total + pad_rows, dtype=col.dtype, device=col.device)
            out.append(dst)
            me = self.rank
            dst[recv_off[me]:recv_off[me] + recv[me]] = col[send_off[me]:send_off[me] + send[me]]
            for peer in range(self.world):
                if peer == me:
                    continue
                if send[peer]:
                    ops.append(dist.P2POp(dist.isend, col[send_off[peer]:send_off[peer] + send[peer]], peer))
                if recv[peer]:
                    ops.append(dist.P2POp(dist.irecv, dst[recv_off[peer]:recv_off[peer] + recv[peer]], peer))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return [c[:total] if pad_rows == 0 else c for c in out], recv


class PeerShuffle:
    """Host side of the library's row exchange over NVLink peer memory (csrc/shuffle.cu, msc_shuffle_*).

    Created collectively by all ranks.  ``exchange`` = one shuffle of a device relation: rows travel to rank
    ``hash(key) % world`` (or, with ``key_col=None``, every row to every rank: the all-gather of small partial results).
    Receive buffers ("slots") live as long as the engine; they only grow, and because every rank sees the same
    rows[src][dst] matrix all ranks take the same sizing decisions without another message.  A slot is busy until the
    query that received into it releases it (:meth:`release_all`)."""

    GRANULE = 2 << 20

    def __init__(self, ctx: Any, comm: "Comm") -> None:
        import ctypes as C

        self.ctx, self.comm = ctx, comm
        self.handle = C.c_void_p()
        self.epoch = 0
        self.table_epoch = 0
        self.slot_bytes: dict[int, int] = {}
        self.busy: set[int] = set()
        self.last_matrix: list[list[int]] = []
        self._ints: Optional[tuple[int, int, int]] = None  # (capacity, device source, device destination) of allgather_ints
        ipc = C.create_string_buffer(64)
        ok = True
        try:
            ctx.call("msc_shuffle_create", comm.rank, comm.world, C.byref(self.handle), ipc)
        except Exception:  # noqa: BLE001  (no IPC on this system: every rank must learn it)
            ok = False
        everyone = comm.all_gather_object((ok, bytes(ipc.raw)))
        if ok and all(o for o, _ in everyone):
            try:
                ctx.check(ctx.lib.msc_shuffle_attach(self.handle, b"".join(h for _, h in everyone)))
            except Exception:  # noqa: BLE001
                ok = False
        else:
            ok = False
        self.ok = all(comm.all_gather_object(ok))
        if not self.ok:
            self.close()

    def close(self) -> None:
        if self.handle:
            self.ctx.lib.msc_shuffle_free(self.handle)
            self.handle = None

    def release_all(self) -> None:
        self.busy.clear()

    def _ensure_slot(self, slot: int, want: int) -> None:
        """Collective: every rank calls it with the same arguments."""
        import ctypes as C

        if self.slot_bytes.get(slot, 0) >= want:
            return
        size = -(-max(want + want // 4, 16 * self.GRANULE) // self.GRANULE) * self.GRANULE
        if slot in self.slot_bytes:  # peers must unmap the old buffer before its owner frees it
            self.ctx.check(self.ctx.lib.msc_shuffle_slot_detach(self.handle, slot))
            self.comm.barrier()
        ipc = C.create_string_buffer(64)
        self.ctx.check(self.ctx.lib.msc_shuffle_slot_alloc(self.handle, slot, size, ipc))
        handles = self.comm.all_gather_object(bytes(ipc.raw))
        self.ctx.check(self.ctx.lib.msc_shuffle_slot_attach(self.handle, slot, b"".join(handles)))
        self.slot_bytes[slot] = size

    def exchange(self, rel_handle: int, key_col: Optional[int], lower_bounds: Optional[Sequence[int]] = None) -> tuple[int, int]:
        """-> (handle of the received relation -- it wraps the slot, free it with msc_rel_free --, rows received).
        ``lower_bounds``: route by key range instead of by hash -- rank r receives lower_bounds[r] <= key < lower_bounds[r + 1]."""
        import ctypes as C

        world = self.comm.world
        self.epoch += 1
        matrix = (C.c_uint64 * (world * world))()
        need = (C.c_uint64 * world)()
        if lower_bounds is not None:
            bounds = (C.c_int64 * world)(*[int(b) for b in lower_bounds])
            self.ctx.check(self.ctx.lib.msc_shuffle_begin_range(self.handle, C.c_void_p(rel_handle), key_col, bounds, self.epoch, matrix, need))
        else:
            self.ctx.check(self.ctx.lib.msc_shuffle_begin(self.handle, C.c_void_p(rel_handle), -1 if key_col is None else key_col, self.epoch,
                                                          matrix, need))
        slot = next(i for i in range(64) if i not in self.busy)
        self._ensure_slot(slot, max(need))
        out = C.c_void_p()
        self.ctx.check(self.ctx.lib.msc_shuffle_finish(self.handle, slot, self.epoch, C.byref(out)))
        self.busy.add(slot)
        self.last_matrix = [[int(matrix[s * world + d]) for d in range(world)] for s in range(world)]
        return out.value, sum(row[self.comm.rank] for row in self.last_matrix)

    def allgather_ints(self, values: Sequence[int]) -> list[list[int]]:
        """[rank][i] = values[i] of that rank (64-bit integers), through the control blocks: one small kernel and one read-back
        instead of a host collective (a torch.distributed all-gather of a few numbers costs 0.1-0.2 ms)."""
        import ctypes as C

        import numpy as np

        n = len(values)
        world = self.comm.world
        if self._ints is None or self._ints[0] < n:
            src, dst = C.c_void_p(), C.c_void_p()
            self.ctx.call("msc_dev_alloc", 8 * max(n, 16), C.byref(src))
            self.ctx.call("msc_dev_alloc", 8 * max(n, 16) * world, C.byref(dst))
            self._ints = (max(n, 16), src.value, dst.value)
        _, src_ptr, dst_ptr = self._ints
        mine = np.asarray([int(v) for v in values], dtype=np.int64)
        self.ctx.call("msc_memcpy_h2d", C.c_void_p(src_ptr), mine.ctypes.data_as(C.c_void_p), mine.nbytes)
        self.allgather_table(src_ptr, 8 * n, dst_ptr)
        out = np.zeros(world * n, dtype=np.int64)
        self.ctx.call("msc_memcpy_d2h", out.ctypes.data_as(C.c_void_p), C.c_void_p(dst_ptr), out.nbytes)
        return out.reshape(world, n).tolist()

    def allgather_table(self, src_ptr: int, nbytes: int, dst_ptr: int) -> None:
        """`nbytes` from every rank into dst[world][nbytes] (device), stream-ordered, no host wait."""
        import ctypes as C

        self.table_epoch += 1
        self.ctx.check(self.ctx.lib.msc_shuffle_allgather(self.handle, C.c_void_p(src_ptr), nbytes, self.table_epoch, C.c_void_p(dst_ptr)))
