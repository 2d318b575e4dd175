"""Sharding helpers: one process per GPU, torch.distributed for the plumbing.

The path shards only over independent units (particles / restarts / prediction points):
rows are block-partitioned, every rank runs the same kernels on its block, and the only collective is
the all-gather of the per-row results (NCCL on GPUs; gloo in the CPU tests of the partitioning logic).
"""
from __future__ import annotations

import torch
import torch.distributed as td


def world(group=None):
    if td.is_available() and td.is_initialized():
        return td.get_rank(group), td.get_world_size(group)
    return 0, 1


def block_bounds(N, rank, size):
    """[lo, hi) of `rank` in a balanced block partition of N rows (first N % size ranks get one more)."""
    base, rem = divmod(N, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_bounds(N, group=None):
    rank, size = world(group)
    return block_bounds(N, rank, size)


def all_gather_rows(local, N, group=None):
    """Concatenate the ranks' blocks (block_bounds layout) of a 1-D tensor into the full length-N tensor."""
    rank, size = world(group)
    if size == 1:
        return local
    base, rem = divmod(N, size)
    width = base + (1 if rem else 0)
    pad = torch.zeros(width, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    parts = [torch.empty_like(pad) for _ in range(size)]
    td.all_gather(parts, pad, group=group)
    chunks = []
    for r in range(size):
        lo, hi = block_bounds(N, r, size)
        chunks.append(parts[r][: hi - lo])
    return torch.cat(chunks)
