"""Multi-GPU plumbing: one process (rank) per GPU, ``torch.distributed`` for the exchange.

The reference's only parallelism is data-parallel row-blocks plus a hash-partitioned shuffle through
files (``src/mini_spark/plan.py:90-109``: one ``ScanJob`` per block, partitions re-read per
``LoadShuffleFilesJob``; ``tasks.py:347-375`` writes ``hash(key) % 10`` buckets).  Here:

* row-blocks are dealt round-robin to ranks (:func:`shard_blocks`); every rank scans only its blocks;
* a low-cardinality GROUP BY merges tiny per-rank partial tables: string keys are unified through
  their dictionary *entries* (:func:`unify_keys`), the partial rows are all-gathered and re-aggregated
  on every GPU -- no row-level data crosses the fabric;
* a high-cardinality GROUP BY / JOIN routes rows by ``hash(key) % world`` (``msc_partition``) and
  exchanges them with one all-to-all (:meth:`Comm.all_to_all_rows`), the GPU analogue of the
  reference's shuffle files.

This module holds only host logic on ``torch`` tensors, so it runs under ``gloo`` on CPU (tests) and
under ``nccl`` on GPUs unchanged.  PyTorch is plumbing here (buffers + collectives), not compute.
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Optional, Sequence


def shard_blocks(nblocks: int, rank: int, world: int) -> list[int]:
    """Row-blocks owned by ``rank``: block ``b`` goes to rank ``b % world`` (cf. plan.py:90-93)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return [b for b in range(nblocks) if b % world == rank]


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

    def all_gather_rows(self, columns: Sequence[Any], nrows: int) -> tuple[list[Any], list[int]]:
        """All-gather a small relation given as 1-D tensors of ``nrows`` elements each.

        Returns (concatenated columns, rows per rank).  Ragged row counts are padded to the maximum."""
        import torch

        counts = self.all_gather_object(int(nrows))
        if self.world == 1:
            return [c[:nrows].clone() for c in columns], counts
        import torch.distributed as dist

        width = max(max(counts), 1)
        out = []
        for col in columns:
            padded = torch.zeros(width, dtype=col.dtype, device=col.device)
            padded[:nrows] = col[:nrows]
            parts = [torch.empty_like(padded) for _ in range(self.world)]
            dist.all_gather(parts, padded)
            out.append(torch.cat([p[:n] for p, n in zip(parts, counts)]))
        return out, counts

    def all_to_all_rows(self, columns: Sequence[Any], send_counts: Sequence[int]) -> tuple[list[Any], list[int]]:
        """Exchange partition-contiguous rows: the first ``send_counts[0]`` rows of every column go to
        rank 0, the next ``send_counts[1]`` to rank 1, ...  Returns (received columns, recv_counts)."""
        import torch

        if self.world == 1:
            return [c[:send_counts[0]].clone() for c in columns], [int(send_counts[0])]
        import torch.distributed as dist

        matrix = self.all_gather_object([int(c) for c in send_counts])
        send, recv = exchange_plan(matrix, self.rank)
        total = sum(recv)
        out = []
        for col in columns:
            dst = torch.empty(total, dtype=col.dtype, device=col.device)
            src = col[:sum(send)].contiguous()
            if dist.get_backend() == "gloo":  # gloo has no all_to_all_single: emulate with per-peer broadcasts of slices
                pieces = self.all_gather_object(src.cpu())
                offs = 0
                for s in range(self.world):
                    lo = sum(matrix[s][:self.rank])
                    n = matrix[s][self.rank]
                    dst[offs:offs + n] = pieces[s][lo:lo + n].to(dst.device)
                    offs += n
            else:
                dist.all_to_all_single(dst, src, output_split_sizes=recv, input_split_sizes=send)
            out.append(dst)
        return out, recv
