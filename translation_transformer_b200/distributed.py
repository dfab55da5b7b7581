"""Data-parallel sharding of query batches (one process per GPU, torch.distributed plumbing).

The hot path partitions by query: every query is decoded independently, so ranks never exchange
activations.  The only collective is the all-gather that collects the predictions (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  Reference context: the reference runs single-GPU
(`--trainer.devices [0]`, scripts/product_prediction.sh); Lightning's DDP predict would shard the
dataloader the same way (contiguous, order-preserving shards here so the CSV order is kept).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batches(n_batches: int, rank: int, world_size: int) -> range:
    lo, hi = shard_bounds(n_batches, rank, world_size)
    return range(lo, hi)


def gather_predictions(local: torch.Tensor, counts: list[int] | None = None, group=None) -> torch.Tensor:
    """All-gather (n_local, n_best, max_len) int64 predictions from every rank, rank order.

    Shards may be ragged (`counts[r]` rows on rank r); they are padded to the longest shard for the
    collective and trimmed afterwards.  With world_size 1 (or no process group) this is the identity.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if counts is None:
        c = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_c = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(all_c, c, group=group)
        counts = [int(x.item()) for x in all_c]
    longest = max(counts)
    padded = local
    if local.shape[0] < longest:
        pad = torch.zeros((longest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if out.is_cuda:
        dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
        chunks = out.view(world, longest, *local.shape[1:])
    else:
        lst = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(lst, padded.contiguous(), group=group)
        chunks = torch.stack(lst, dim=0)
    return torch.cat([chunks[r, :counts[r]] for r in range(world)], dim=0)
