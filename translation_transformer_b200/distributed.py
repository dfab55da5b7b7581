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


class BatchQueue:
    """Dynamic batch queue shared by all ranks of a job: `next()` hands out the indices 0 .. n_items-1 exactly once across
    every rank and host thread, in order, whoever asks first (the counter lives in the process group's store on rank 0;
    without a process group it is a local counter).  Replaces the static contiguous split of `shard_batches` where finish
    times are ragged: a rank that drew long queries simply draws fewer batches."""

    _serial = 0

    def __init__(self, n_items: int, name: str | None = None, store=None, local: bool = False):
        import threading
        self.n_items = n_items
        self._lock = threading.Lock()
        self._local = 0
        self._store = None if local else store
        if not local and store is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._store = dist.distributed_c10d._get_default_store()
        # every rank constructs its queues in the same order, so the serial number names the same counter everywhere
        BatchQueue._serial += 1
        self._key = name or f"ttb_batch_queue_{BatchQueue._serial}"

    def next(self) -> int | None:
        if self._store is not None:
            i = int(self._store.add(self._key, 1)) - 1
        else:
            with self._lock:
                i = self._local
                self._local += 1
        return i if i < self.n_items else None


def gather_indexed_predictions(indices: list[int], preds: list[torch.Tensor], n_total: int, device=None, group=None) -> list[torch.Tensor | None]:
    """Collect `(batch index, prediction)` pairs decoded by any rank into a list of `n_total` predictions in batch order on
    every rank (ONE all-gather of the padded predictions + one of the indices; identity without a process group).  All
    predictions must share their trailing dimensions (batch, n_best, max_len); a rank may hold any number of them."""
    out: list[torch.Tensor | None] = [None] * n_total
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        for i, p in zip(indices, preds):
            out[i] = p
        return out
    world = dist.get_world_size(group)
    backend_cuda = dist.get_backend(group) == "nccl"
    dev = device if device is not None else (preds[0].device if preds else torch.device("cuda" if backend_cuda else "cpu"))
    shape = torch.tensor(list(preds[0].shape) if preds else [0, 0, 0], dtype=torch.int64, device=dev)
    meta = torch.cat([torch.tensor([len(preds)], dtype=torch.int64, device=dev), shape])
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    counts = [int(m[0]) for m in metas]
    tail = next((tuple(int(v) for v in m[1:]) for m in metas if int(m[0]) > 0), None)
    if tail is None:
        return out
    longest = max(counts)
    local = torch.zeros((longest,) + tail, dtype=torch.int64, device=dev)
    idx = torch.full((longest,), -1, dtype=torch.int64, device=dev)
    for k, (i, p) in enumerate(zip(indices, preds)):
        local[k] = p.to(dev)
        idx[k] = i
    all_p = torch.empty((world * longest,) + tail, dtype=torch.int64, device=dev)
    all_i = torch.empty((world * longest,), dtype=torch.int64, device=dev)
    if backend_cuda:
        dist.all_gather_into_tensor(all_p, local, group=group)
        dist.all_gather_into_tensor(all_i, idx, group=group)
    else:
        lp = [torch.empty_like(local) for _ in range(world)]
        li = [torch.empty_like(idx) for _ in range(world)]
        dist.all_gather(lp, local, group=group)
        dist.all_gather(li, idx, group=group)
        all_p, all_i = torch.cat(lp), torch.cat(li)
    for k, i in enumerate(all_i.tolist()):
        if i >= 0:
            out[i] = all_p[k]
    return out


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
