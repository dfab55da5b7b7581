#!/usr/bin/env python
"""Several batches in flight on ONE GPU: K engines (each with its private stream, workspace and CUDA graphs) decode
independent bs=32 batches of the bench.py workload from K host threads (the C ABI releases the GIL).

One decoding iteration is a chain of 23 one-wave kernels (128 CTAs on 148 SMs) bounded by dependent latencies
(DESIGN.md §8); a second batch in flight fills the SMs a kernel of the first leaves idle while it drains / starts.
Outputs are checked to be identical to the single-engine run.

    python scripts/bench_pipelined.py --in-flight 1 2 3 --batches 8
"""
from __future__ import annotations

import argparse
import json
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

import bench  # noqa: E402
from translation_transformer_b200.decoding import TranslationInferenceGreedySpeculative  # noqa: E402
from translation_transformer_b200.model import B200Transformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-flight", type=int, nargs="+", default=[1, 2, 3])
    ap.add_argument("--batches", type=int, default=12, help="batches decoded per measurement (shared by the engines)")
    ap.add_argument("--warmup", type=int, default=3)
    cli = ap.parse_args()
    sys.argv = sys.argv[:1]
    args = bench.parse()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg, sd = bench.build_weights(args)
    host = [bench.batch_for(args, 0, i).pin_memory() for i in range(cli.batches)]
    kmax = max(cli.in_flight)
    engines, gens, streams = [], [], []
    for _ in range(kmax):
        e = B200Transformer(cfg, sd, precision=args.precision, device=0)
        engines.append(e)
        gens.append(TranslationInferenceGreedySpeculative(e, args.max_len, args.draft_len, args.n_drafts, bench.PAD, bench.BOS,
                                                          bench.EOS, bench.REPLACE))
        streams.append(torch.cuda.Stream(device=dev))
    outs_ref = None
    for k in cli.in_flight:
        outs = [None] * cli.batches
        errs = []

        def worker(j, first, count):
            torch.cuda.set_device(0)
            with torch.cuda.stream(streams[j]):
                for i in range(first, first + count):
                    idx = i % cli.batches
                    try:
                        o = gens[j].generate(host[idx].to(dev, non_blocking=True))
                        outs[idx] = o.cpu()
                    except RuntimeError as ex:
                        errs.append(str(ex)[:80])
                        outs[idx] = None

        def run(total):
            per = [total // k + (1 if j < total % k else 0) for j in range(k)]
            first, th = 0, []
            for j in range(k):
                th.append(threading.Thread(target=worker, args=(j, first, per[j])))
                first += per[j]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            return time.perf_counter() - t0

        run(cli.warmup * k)
        dt = run(cli.batches)
        same = None
        if outs_ref is None:
            outs_ref = [o.clone() if o is not None else None for o in outs]
        else:
            same = all((a is None and b is None) or (a is not None and b is not None and torch.equal(a, b))
                       for a, b in zip(outs, outs_ref))
        print(json.dumps({"in_flight": k, "batches": cli.batches, "batch_size": args.batch_size,
                          "smiles_per_s": round(cli.batches * args.batch_size / dt, 1), "ms_per_batch": round(1000 * dt / cli.batches, 2),
                          "identical_to_first_run": same, "reference_failures": errs}), flush=True)
    for e in engines:
        e.close()


if __name__ == "__main__":
    main()
