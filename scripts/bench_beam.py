#!/usr/bin/env python
"""Timing of the speculative beam search on BASELINE.json configs[2] / configs[3] (synthetic sources, random-init
weights).  Not the bench line (bench.py measures configs[1]); reported in DESIGN.md."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from translation_transformer_b200.decoding import TranslationInferenceBeamSearchSpeculative  # noqa: E402
from translation_transformer_b200.model import B200Transformer  # noqa: E402
from translation_transformer_b200.synthetic import synthetic_sources  # noqa: E402
from translation_transformer_b200.weights import ModelConfig, random_init_state_dict  # noqa: E402

CONFIGS = {
    "product_bs4_nbest5": dict(layers=4, bs=4, nbest=5, draft_len=10, n_drafts=23),
    "retro_bs8_nbest10": dict(layers=6, bs=8, nbest=10, draft_len=10, n_drafts=23),
    "retro_bs8_nbest20": dict(layers=6, bs=8, nbest=20, draft_len=10, n_drafts=23),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="product_bs4_nbest5", choices=list(CONFIGS))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--max-len", type=int, default=200)
    ap.add_argument("--vocab", type=int, default=288)
    ap.add_argument("--eos-bias", type=float, default=2.0, help="random-init models never stop on their own; bias EOS so "
                                                                  "that hypotheses finish at USPTO-like lengths")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--profile", action="store_true", help="per kernel class CUDA-event times of one batch")
    ap.add_argument("--in-flight", type=int, nargs="+", default=[1],
                    help="batches decoded concurrently (one engine + host thread each, pipeline.py); several values are run in turn")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    cfg = ModelConfig(src_vocab_size=a.vocab, tgt_vocab_size=a.vocab, embedding_dim=256, feedforward_dim=2048,
                      num_encoder_layers=c["layers"], num_decoder_layers=c["layers"], num_heads=8)
    sd = {k: v.clone() for k, v in random_init_state_dict(cfg, 1234).items()}
    sd["tgt_token_featurizer.embedding.weight"] = sd["src_token_featurizer.embedding.weight"]
    sd["next_token_classifier.bias"][2] += a.eos_bias
    sd["next_token_classifier.bias"][0] -= 5.0
    eng = B200Transformer(cfg, sd, precision=a.precision, device=0)
    gen = TranslationInferenceBeamSearchSpeculative(eng, a.max_len, c["nbest"], c["draft_len"], c["n_drafts"], a.vocab, False, 0, 1, 2, 7)
    dev = torch.device("cuda", 0)
    times, calls, errs = [], [], 0
    for i in range(a.warmup + a.steps):
        src = synthetic_sources(c["bs"], a.vocab, seed=7000 + i).to(dev)
        c0 = gen.model_calls_num
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            out = gen.generate(src)
        except RuntimeError:
            errs += 1
        torch.cuda.synchronize()
        if i >= a.warmup:
            times.append(time.perf_counter() - t0)
            calls.append(gen.model_calls_num - c0)
    print(json.dumps({"config": a.config, "precision": a.precision, "smiles_per_s": c["bs"] * len(times) / sum(times),
                      "ms_per_batch": 1000 * sum(times) / len(times), "decoder_calls_per_batch": sum(calls) / len(calls),
                      "accepted_tokens": gen.accepted_tokens_num, "reference_failures": errs}), flush=True)
    if a.profile:   # CUDA-event time of every kernel class in one more (eagerly launched, serialised) batch
        import ctypes as C
        lib = eng.lib
        n_cls = lib.ttb_kernel_class_count()
        names = [lib.ttb_kernel_class_name(i).decode() for i in range(n_cls)]
        lib.ttb_engine_set_profiling(eng._h, (1 << n_cls) - 1)
        gen.generate(synthetic_sources(c["bs"], a.vocab, seed=7000 + a.warmup).to(dev))
        ms_arr, n_arr = (C.c_double * n_cls)(), (C.c_int64 * n_cls)()
        lib.ttb_engine_get_profile(eng._h, n_cls, ms_arr, n_arr)
        lib.ttb_engine_set_profiling(eng._h, 0)
        tot = sum(ms_arr) or 1.0
        for i in range(n_cls):
            if n_arr[i]:
                print(f"  {names[i]:18s} {ms_arr[i]:8.3f} ms {n_arr[i]:6d} launches {1000 * ms_arr[i] / n_arr[i]:7.2f} us  {ms_arr[i] / tot:5.3f}", flush=True)
    # ---- several batches in flight: these searches keep only a few hundred token rows busy (2-8 row blocks of the GEMM
    # kernels on 148 SMs), so independent batches overlap almost freely
    from translation_transformer_b200.pipeline import InFlightDecoder
    kmax = max(a.in_flight)
    if kmax > 1:
        gens = [gen] + [TranslationInferenceBeamSearchSpeculative(B200Transformer(cfg, sd, precision=a.precision, device=0), a.max_len, c["nbest"],
                                                                  c["draft_len"], c["n_drafts"], a.vocab, False, 0, 1, 2, 7) for _ in range(kmax - 1)]
        ref = None
        for k in a.in_flight:
            if k < 2:
                continue
            fly = InFlightDecoder(gens[:k], device=0)
            n = max(a.steps, 2) * k
            srcs = [synthetic_sources(c["bs"], a.vocab, seed=7000 + a.warmup + (i % a.steps)) for i in range(n)]
            list(fly.map(srcs[:k], pre=lambda s: s.to(dev), on_error=lambda i, ex: None))     # warm-up of every engine
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            outs = [o.cpu() if o is not None else None for o in fly.map(srcs, pre=lambda s: s.to(dev), on_error=lambda i, ex: None)]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            same = all((x is None and y is None) or (x is not None and y is not None and torch.equal(x, y))
                       for x, y in zip(outs[:a.steps], outs[a.steps:2 * a.steps]))   # the same sources decoded by different engines
            print(json.dumps({"config": a.config, "in_flight": k, "smiles_per_s": c["bs"] * n / dt, "ms_per_batch": 1000 * dt / n,
                              "batches": n, "repeat_identical": same}), flush=True)
            fly.close()


if __name__ == "__main__":
    main()
