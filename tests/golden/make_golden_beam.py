"""Golden fixtures for the speculative beam search (run through make_golden.py --only beam)."""
from __future__ import annotations

import json

import numpy as np
import torch

from make_golden import HERE, SMALL, Hooks, ModelConfig, load_test_sources, ref_model, state_dict_checksum, synthetic_sources


def gen_beam(ref):
    VanillaTransformer, _, spec, _, Tok = ref
    tk, src, _, _ = load_test_sources(Tok)
    cases, arrays = [], {}

    def run(cid, model, s, max_len, n_best, D, N, c_token, vocab):
        g = spec.TranslationInferenceBeamSearchSpeculative(model, max_len=max_len, n_best=n_best, draft_len=D, n_drafts=N,
                                                           vocab_size=vocab, smart_drafts_mode=False, pad_token=0, bos_token=1,
                                                           eos_token=2, C_token=c_token)
        rec = {"id": cid, "max_len": max_len, "n_best": n_best, "draft_len": D, "n_drafts": N, "C_token": c_token,
               "B": int(s.shape[0]), "vocab": vocab}
        with Hooks(model) as h, torch.inference_mode():
            try:
                out = g.generate(s)
                rec["error"] = None
                arrays[f"{cid}_out"] = out.numpy().astype(np.int16)
            except Exception as e:
                rec["error"] = type(e).__name__
                rec["error_msg"] = str(e)[:200]
        rec["model_calls"] = g.model_calls_num
        rec["accepted_tokens"] = int(g.accepted_tokens_num)
        rec["produced_non_pad_tokens"] = int(g.produced_non_pad_tokens)
        rec["decoder_input_sha1"] = h.calls
        rec["topk1_shapes"] = [list(t[0].shape) for t in h.topk1]
        arrays[f"{cid}_nacc"] = np.concatenate([t[0].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
        arrays[f"{cid}_pick"] = np.concatenate([t[1].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
        arrays[f"{cid}_src"] = s.numpy().astype(np.int16)
        cases.append(rec)
        print(cid, "calls", g.model_calls_num, "error", rec["error"], rec.get("error_msg", "")[:60])

    idx = 0
    cfg = ModelConfig(src_vocab_size=tk.n_tokens, tgt_vocab_size=tk.n_tokens, **SMALL)
    for (seed, eos_bias, B, max_len, n_best, D, N) in ((33, 0.5, 1, 60, 5, 10, 23), (33, 0.5, 4, 60, 5, 10, 7), (46, 0.5, 2, 80, 3, 14, 10),
                                                       (46, 0.7, 3, 50, 5, 9, 10), (37, 1.1, 2, 100, 10, 10, 10), (47, 0.9, 4, 40, 5, 5, 3),
                                                       (33, 0.9, 8, 70, 10, 10, 2), (21, 0.0, 2, 30, 5, 10, 7)):
        m, sd = ref_model(VanillaTransformer, cfg, seed)
        with torch.no_grad():
            m.next_token_classifier.bias[2] += eos_bias
            m.next_token_classifier.bias[0] -= 5.0
        cid = f"beam{idx}"
        run(cid, m, src[:B], max_len, n_best, D, N, tk.encode("c")[1], tk.n_tokens)
        cases[-1].update(arch="small", seed=seed, eos_bias=eos_bias, pad_bias=-5.0, checksum=state_dict_checksum(sd), source="test_file")
        idx += 1
    cfg300 = ModelConfig(src_vocab_size=300, tgt_vocab_size=300, **SMALL)
    syn = synthetic_sources(300, 8, 20, 90, seed=4)
    for (seed, eos_bias, B, max_len, n_best, D, N) in ((32, 0.9, 4, 80, 5, 10, 7), (33, 0.9, 8, 60, 10, 10, 2)):
        m, sd = ref_model(VanillaTransformer, cfg300, seed)
        with torch.no_grad():
            m.next_token_classifier.bias[2] += eos_bias
            m.next_token_classifier.bias[0] -= 5.0
        cid = f"beamsyn{idx}"
        run(cid, m, syn[:B], max_len, n_best, D, N, 7, 300)
        cases[-1].update(arch="small", seed=seed, eos_bias=eos_bias, pad_bias=-5.0, checksum=state_dict_checksum(sd), source="synthetic")
        idx += 1
    np.savez_compressed(HERE / "beam_speculative.npz", **arrays)
    json.dump(cases, open(HERE / "beam_speculative.json", "w"))
    print("beam written")


def gen_beam_smart(ref):
    """`smart_drafts_mode=True` (speculative_decoding.py:600-845): fixtures in beam_smart.{npz,json}."""
    VanillaTransformer, _, spec, _, Tok = ref
    tk, src, _, _ = load_test_sources(Tok)
    cases, arrays = [], {}

    def run(cid, model, s, max_len, n_best, D, N, c_token, vocab):
        g = spec.TranslationInferenceBeamSearchSpeculative(model, max_len=max_len, n_best=n_best, draft_len=D, n_drafts=N,
                                                           vocab_size=vocab, smart_drafts_mode=True, pad_token=0, bos_token=1,
                                                           eos_token=2, C_token=c_token)
        rec = {"id": cid, "max_len": max_len, "n_best": n_best, "draft_len": D, "n_drafts": N, "C_token": c_token,
               "B": int(s.shape[0]), "vocab": vocab}
        with Hooks(model) as h, torch.inference_mode():
            try:
                out = g.generate(s)
                rec["error"] = None
                arrays[f"{cid}_out"] = out.numpy().astype(np.int16)
            except Exception as e:
                rec["error"] = type(e).__name__
                rec["error_msg"] = str(e)[:200]
        rec["model_calls"] = g.model_calls_num
        rec["accepted_tokens"] = int(g.accepted_tokens_num)
        rec["produced_non_pad_tokens"] = int(g.produced_non_pad_tokens)
        rec["model_input_lines_num"] = int(g.model_input_lines_num)
        rec["decoder_input_sha1"] = h.calls
        rec["topk1_shapes"] = [list(t[0].shape) for t in h.topk1]
        arrays[f"{cid}_nacc"] = np.concatenate([t[0].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
        arrays[f"{cid}_pick"] = np.concatenate([t[1].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
        arrays[f"{cid}_src"] = s.numpy().astype(np.int16)
        cases.append(rec)
        print(cid, "calls", g.model_calls_num, "lines", g.model_input_lines_num, "error", rec["error"], rec.get("error_msg", "")[:60])

    idx = 0
    cfg = ModelConfig(src_vocab_size=tk.n_tokens, tgt_vocab_size=tk.n_tokens, **SMALL)
    for (seed, eos_bias, B, max_len, n_best, D, N) in ((33, 0.5, 1, 60, 5, 10, 23), (33, 0.5, 4, 60, 5, 10, 7), (46, 0.5, 2, 80, 3, 14, 10),
                                                       (46, 0.7, 3, 50, 5, 9, 10), (37, 1.1, 2, 100, 10, 10, 10), (47, 0.9, 4, 40, 5, 5, 3),
                                                       (33, 0.9, 8, 70, 10, 10, 2), (21, 0.0, 2, 30, 5, 10, 7)):
        m, sd = ref_model(VanillaTransformer, cfg, seed)
        with torch.no_grad():
            m.next_token_classifier.bias[2] += eos_bias
            m.next_token_classifier.bias[0] -= 5.0
        cid = f"smart{idx}"
        run(cid, m, src[:B], max_len, n_best, D, N, tk.encode("c")[1], tk.n_tokens)
        cases[-1].update(arch="small", seed=seed, eos_bias=eos_bias, pad_bias=-5.0, checksum=state_dict_checksum(sd), source="test_file")
        idx += 1
    cfg300 = ModelConfig(src_vocab_size=300, tgt_vocab_size=300, **SMALL)
    syn = synthetic_sources(300, 8, 20, 90, seed=4)
    for (seed, eos_bias, B, max_len, n_best, D, N) in ((32, 0.9, 4, 80, 5, 10, 7), (33, 0.9, 8, 60, 10, 10, 2)):
        m, sd = ref_model(VanillaTransformer, cfg300, seed)
        with torch.no_grad():
            m.next_token_classifier.bias[2] += eos_bias
            m.next_token_classifier.bias[0] -= 5.0
        cid = f"smartsyn{idx}"
        run(cid, m, syn[:B], max_len, n_best, D, N, 7, 300)
        cases[-1].update(arch="small", seed=seed, eos_bias=eos_bias, pad_bias=-5.0, checksum=state_dict_checksum(sd), source="synthetic")
        idx += 1
    np.savez_compressed(HERE / "beam_smart.npz", **arrays)
    json.dump(cases, open(HERE / "beam_smart.json", "w"))
    print("beam smart written")
