"""Golden fixtures AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs[1..3]), from the UNMODIFIED reference on CPU.

    python tests/golden/make_golden_bench.py [--only cfg1_copy,cfg1_random,cfg2_copy,cfg3_copy,cfg3_copy20]

Same method as make_golden.py (the reference is imported from /root/reference/src with `pytorch_lightning`
stubbed, observation hooks installed from the outside); the cases are the exact workloads `bench.py` times:

  cfg1_copy    greedy speculative, product-prediction arch 256/2048/4+4/8, vocab 288, the bench's first timed batch
               (synthetic_sources(32, 288, seed=100003)), draft_len 10, n_drafts 23, max_len 200, weights
               `copy_task_state_dict(seed 1234)` (trained-like: every query finishes at its own length)
  cfg1_random  the same batch with plain random-init weights (round-1 headline: nothing finishes, outputs all PAD;
               pinned through the per-iteration decoder inputs / accepted lengths / picks)
  cfg2_copy    speculative beam search bs 4, n_best 5, draft_len 10, n_drafts 23 (scripts/product_prediction.sh defaults)
  cfg3_copy    single-step retrosynthesis arch 6+6, bs 8, n_best 10, draft_len 10, n_drafts 2
               (scripts/single_step_retrosynthesis.sh:166-174), USPTO-50k-shape sources
  cfg3_copy20  the same arch, bs 8, n_best 20, draft_len 14, n_drafts 5 (the script's n_best 20 setting)

Writes tests/golden/bench_configs.{npz,json}; a case that already exists is kept unless named in --only.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import Hooks, import_reference  # noqa: E402
from translation_transformer_b200.synthetic import synthetic_sources  # noqa: E402
from translation_transformer_b200.weights import (ModelConfig, PRODUCT_PREDICTION, SINGLE_STEP_RETRO, copy_task_state_dict,  # noqa: E402
                                                  random_init_state_dict, state_dict_checksum)

VOCAB, SEED = 288, 1234
RETRO_SRC = dict(mean_len=45.0, std_len=15.0, min_len=12, max_len=150)   # USPTO-50k products are shorter than USPTO-MIT reactant sets

CASES = {
    "cfg1_copy": dict(kind="greedy", arch="product", weights="copy", B=32, src_seed=100003, max_len=200, draft_len=10, n_drafts=23),
    "cfg1_random": dict(kind="greedy", arch="product", weights="random", B=32, src_seed=100003, max_len=200, draft_len=10, n_drafts=23),
    "cfg2_copy": dict(kind="beam", arch="product", weights="copy", B=4, src_seed=100003, max_len=200, n_best=5, draft_len=10, n_drafts=23),
    "cfg3_copy": dict(kind="beam", arch="retro", weights="copy", B=8, src_seed=200003, max_len=200, n_best=10, draft_len=10, n_drafts=2),
    "cfg3_copy20": dict(kind="beam", arch="retro", weights="copy", B=8, src_seed=200003, max_len=200, n_best=20, draft_len=14, n_drafts=5),
    # smart_drafts_mode=True at the configs[2] shape (speculative_decoding.py:600-845)
    "cfg2_smart_copy": dict(kind="beam", arch="product", weights="copy", B=4, src_seed=100003, max_len=200, n_best=5, draft_len=10, n_drafts=23, smart=True),
    # BASELINE.json configs[0]: bs 1, draft_len 10, the reference's own test sources (tokenizer trained on the test file), one case per line
    **{f"cfg0_copy_line{i}": dict(kind="greedy", arch="product", weights="copy", B=1, src_seed=-1, max_len=200, draft_len=10, n_drafts=23, file_line=i)
       for i in range(4)},
}
ARCH = {"product": PRODUCT_PREDICTION, "retro": SINGLE_STEP_RETRO}


def case_inputs(c):
    if "file_line" in c:     # the reference's test file, tokenized with the reference's tokenizer trained on that file
        from make_golden import load_test_sources
        tk, src, _, _ = load_test_sources(import_reference()[4])
        cfg = ModelConfig(src_vocab_size=tk.n_tokens, tgt_vocab_size=tk.n_tokens, **ARCH[c["arch"]])
        row = src[c["file_line"]:c["file_line"] + 1]
        row = row[:, :int((row != 0).sum())]
        return cfg, copy_task_state_dict(cfg, SEED), row
    cfg = ModelConfig(src_vocab_size=VOCAB, tgt_vocab_size=VOCAB, **ARCH[c["arch"]])
    sd = copy_task_state_dict(cfg, SEED) if c["weights"] == "copy" else random_init_state_dict(cfg, SEED)
    kw = RETRO_SRC if c["arch"] == "retro" else {}
    src = synthetic_sources(32 if c["arch"] == "product" else c["B"], VOCAB, seed=c["src_seed"], **kw)[:c["B"]]
    return cfg, sd, src


def run_case(ref, name, c, arrays):
    VanillaTransformer, _, spec, _, _ = ref
    cfg, sd, src = case_inputs(c)
    m = VanillaTransformer(cfg.src_vocab_size, cfg.tgt_vocab_size, cfg.num_encoder_layers, cfg.num_decoder_layers, cfg.embedding_dim,
                           cfg.num_heads, cfg.feedforward_dim, 0.1, "relu", cfg.share_embeddings, 0, 0)
    m.load_state_dict(sd, strict=True)
    m.eval()
    if c["kind"] == "greedy":
        g = spec.TranslationInferenceGreedySpeculative(m, max_len=c["max_len"], draft_len=c["draft_len"], n_drafts=c["n_drafts"],
                                                       pad_token=0, bos_token=1, eos_token=2, replace_token=7)
        # (the replace token of the test-file vocabulary is whatever id 7 is there: any non-service token serves)
    else:
        g = spec.TranslationInferenceBeamSearchSpeculative(m, max_len=c["max_len"], n_best=c["n_best"], draft_len=c["draft_len"],
                                                           n_drafts=c["n_drafts"], vocab_size=cfg.tgt_vocab_size, smart_drafts_mode=bool(c.get("smart", False)),
                                                           pad_token=0, bos_token=1, eos_token=2, C_token=7)
    rec = dict(c, id=name, vocab=cfg.tgt_vocab_size, seed=SEED, checksum=state_dict_checksum(sd), src_kw=RETRO_SRC if c["arch"] == "retro" else {})
    t0 = time.time()
    with Hooks(m) as h, torch.inference_mode():
        try:
            out = g.generate(src)
            rec["error"] = None
            arrays[f"{name}_out"] = out.numpy().astype(np.int16)
        except Exception as e:  # reference failure modes are part of the behaviour
            rec["error"] = type(e).__name__
            rec["error_msg"] = str(e)[:200]
    rec["seconds_reference_cpu"] = round(time.time() - t0, 1)
    rec["cpu_threads"] = torch.get_num_threads()
    rec["model_calls"] = int(g.model_calls_num)
    if c["kind"] == "beam":
        rec["accepted_tokens"] = int(g.accepted_tokens_num)
        rec["produced_non_pad_tokens"] = int(g.produced_non_pad_tokens)
        rec["topk1_shapes"] = [list(t[0].shape) for t in h.topk1]
    else:
        rec["rows_per_iter"] = [int(t[0].shape[0]) for t in h.topk1]
    rec["decoder_input_sha1"] = h.calls
    arrays[f"{name}_nacc"] = np.concatenate([t[0].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
    arrays[f"{name}_pick"] = np.concatenate([t[1].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
    arrays[f"{name}_src"] = src.numpy().astype(np.int16)
    print(name, "calls", rec["model_calls"], "error", rec["error"], f"{rec['seconds_reference_cpu']} s", flush=True)
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(CASES))
    ap.add_argument("--threads", type=int, default=16)
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    ref = import_reference()
    npz, js = HERE / "bench_configs.npz", HERE / "bench_configs.json"
    arrays = dict(np.load(npz)) if npz.exists() else {}
    cases = {c["id"]: c for c in json.load(open(js))} if js.exists() else {}
    for name in args.only.split(","):
        for k in [k for k in arrays if k.startswith(name + "_") and k[len(name) + 1:] in ("out", "nacc", "pick", "src")]:
            del arrays[k]
        cases[name] = run_case(ref, name, CASES[name], arrays)
        np.savez_compressed(npz, **arrays)
        json.dump([cases[k] for k in sorted(cases)], open(js, "w"))
    print("bench_configs written:", sorted(cases))
