"""Several query batches in flight on one GPU.

One decoding iteration of a batch is a chain of ~23 one-wave kernels bounded by dependent latencies
(DESIGN.md §8): while a kernel of one batch drains its stores or waits for its first TMA tile, the SMs
have nothing else to run.  A second batch decoded at the same time by a second engine (own stream,
workspace, KV caches and CUDA graphs) fills those gaps.  Batches are independent in the reference (one
`predict_step` per batch, lightning_model.py:236-239), so the predictions are exactly those of the
sequential loop; only their completion order changes, and `InFlightDecoder` hands them back in
submission order.

The generators are driven from host threads: the C ABI call blocks until its batch is decoded and
ctypes releases the GIL for its duration.
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Iterable, Iterator, Sequence

import torch


class InFlightDecoder:
    """`generators`: decoding strategies (anything with `generate(src)`), each bound to its OWN engine.
    `submit(src)` returns a Future of `post(generate(src))`; at most `len(generators)` batches run at once."""

    def __init__(self, generators: Sequence, device: int | torch.device | None = None) -> None:
        assert len(generators) >= 1
        engines = [id(getattr(g, "model", g)) for g in generators]
        assert len(set(engines)) == len(engines), "every generator in flight needs its own engine"
        self.generators = list(generators)
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        self._free: queue.SimpleQueue = queue.SimpleQueue()
        for g in self.generators:
            s = torch.cuda.Stream(device=self.device) if self.device is not None and self.device.type == "cuda" else None
            self._free.put((g, s))
        self._pool = ThreadPoolExecutor(max_workers=len(self.generators), thread_name_prefix="ttb-inflight")
        self._lock = threading.Lock()

    def __len__(self) -> int:
        return len(self.generators)

    def _job(self, src, pre: Callable | None, post: Callable | None):
        g, s = self._free.get()
        try:
            if s is None:
                x = pre(src) if pre is not None else src
                out = g.generate(x)
                return post(out) if post is not None else out
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(s):
                x = pre(src) if pre is not None else src
                out = g.generate(x)
                if post is not None:
                    out = post(out)
                s.synchronize()
            # the result was allocated on this worker's side stream but is consumed on the caller's stream: tell the
            # caching allocator, so that the block is not handed to the next generate() of this worker while kernels
            # of the consumer that read it are still queued
            for t in (out if isinstance(out, (tuple, list)) else (out,)):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(torch.cuda.current_stream(self.device))
            return out
        finally:
            self._free.put((g, s))

    def submit(self, src, pre: Callable | None = None, post: Callable | None = None) -> Future:
        """`pre` (e.g. the host->device copy) and `post` (e.g. the device->host copy) run on the worker's stream."""
        return self._pool.submit(self._job, src, pre, post)

    def map(self, sources: Iterable, pre: Callable | None = None, post: Callable | None = None,
            on_error: Callable | None = None) -> Iterator:
        """Predictions in submission order.  A `RuntimeError` of a batch (the reference's own failure modes) is passed
        to `on_error(index, exception)`, whose return value stands in for the batch; without it the error is raised."""
        futures = [self.submit(s, pre, post) for s in sources]
        for i, f in enumerate(futures):
            try:
                yield f.result()
            except RuntimeError as ex:
                if on_error is None:
                    raise
                yield on_error(i, ex)

    def drain(self, next_item: Callable, pre: Callable | None = None, post: Callable | None = None,
              on_error: Callable | None = None) -> list:
        """Dynamic batch queue: every worker keeps pulling `(key, src)` pairs from `next_item()` (which must be thread-safe and
        return None once the queue is empty — e.g. a counter shared by all ranks of a job, distributed.BatchQueue) and
        decodes them on its own engine; returns the `(key, result)` pairs this decoder processed, in completion order.
        No static assignment of batches to workers or ranks: a slow batch never holds back the rest of a shard."""
        done: list = []

        def worker():
            while True:
                item = next_item()
                if item is None:
                    return
                key, src = item
                try:
                    res = self._job(src, pre, post)
                except RuntimeError as ex:
                    if on_error is None:
                        raise
                    res = on_error(key, ex)
                with self._lock:
                    done.append((key, res))

        futures = [self._pool.submit(worker) for _ in self.generators]
        for f in futures:
            f.result()
        return done

    def counter(self, name: str):
        """Sum of a per-generator counter (`model_calls_num`, `accepted_tokens_num`, ...)."""
        return sum(getattr(g, name) for g in self.generators)

    def close(self) -> None:
        self._pool.shutdown(wait=True)
