"""Golden fixtures for the standard (non-speculative) greedy and beam-search decoding of the reference
(src/decoding/standard_decoding.py), run through `make_golden.py --only standard`."""
from __future__ import annotations

import json

import numpy as np
import torch

from make_golden import HERE, SMALL, Hooks, ModelConfig, load_test_sources, ref_model, state_dict_checksum, synthetic_sources


def gen_standard(ref):
    VanillaTransformer, _, _, std, Tok = ref
    tk, src, _, _ = load_test_sources(Tok)
    cases, arrays = [], {}

    def run(cid, kind, model, s, max_len, beam):
        if kind == "greedy":
            g = std.TranslationInferenceGreedy(model, max_len=max_len, pad_token=0, bos_token=1, eos_token=2)
        else:
            g = std.TranslationInferenceBeamSearch(model, beam_size=beam, max_len=max_len, pad_token=0, bos_token=1, eos_token=2)
        rec = {"id": cid, "kind": kind, "max_len": max_len, "beam_size": beam, "B": int(s.shape[0])}
        with Hooks(model) as h, torch.inference_mode():
            out = g.generate(s)
        arrays[f"{cid}_out"] = out.numpy().astype(np.int16)
        arrays[f"{cid}_src"] = s.numpy().astype(np.int16)
        rec["model_calls"] = int(g.model_calls_num)
        rec["given_tokens"] = int(g.given_tokens)
        rec["decoder_input_sha1"] = h.calls
        rec["out_shape"] = list(out.shape)
        cases.append(rec)
        print(cid, kind, "calls", g.model_calls_num, "out", tuple(out.shape))

    idx = 0
    cfg = ModelConfig(src_vocab_size=tk.n_tokens, tgt_vocab_size=tk.n_tokens, **SMALL)
    greedy_cfgs = ((21, None, 3, 40), (33, 0.5, 10, 100), (46, 0.7, 6, 60), (37, 1.1, 1, 150), (47, 0.9, 10, 33))
    for (seed, eos_bias, B, max_len) in greedy_cfgs:
        m, sd = ref_model(VanillaTransformer, cfg, seed)
        if eos_bias is not None:
            with torch.no_grad():
                m.next_token_classifier.bias[2] += eos_bias
                m.next_token_classifier.bias[0] -= 5.0
        cid = f"greedy{idx}"
        run(cid, "greedy", m, src[:B], max_len, 0)
        cases[-1].update(arch="small", seed=seed, checksum=state_dict_checksum(sd), vocab=tk.n_tokens, source="test_file")
        if eos_bias is not None:
            cases[-1].update(eos_bias=eos_bias, pad_bias=-5.0)
        idx += 1
    beam_cfgs = ((33, 0.5, 1, 60, 5), (33, 0.5, 4, 60, 5), (46, 0.7, 3, 50, 3), (37, 1.1, 2, 100, 10), (47, 0.9, 8, 40, 5),
                 (21, None, 2, 24, 4))
    for (seed, eos_bias, B, max_len, beam) in beam_cfgs:
        m, sd = ref_model(VanillaTransformer, cfg, seed)
        if eos_bias is not None:
            with torch.no_grad():
                m.next_token_classifier.bias[2] += eos_bias
                m.next_token_classifier.bias[0] -= 5.0
        cid = f"beam{idx}"
        run(cid, "beam", m, src[:B], max_len, beam)
        cases[-1].update(arch="small", seed=seed, checksum=state_dict_checksum(sd), vocab=tk.n_tokens, source="test_file")
        if eos_bias is not None:
            cases[-1].update(eos_bias=eos_bias, pad_bias=-5.0)
        idx += 1
    cfg300 = ModelConfig(src_vocab_size=300, tgt_vocab_size=300, **SMALL)
    syn = synthetic_sources(300, 8, 20, 90, seed=4)
    for kind, (seed, eos_bias, B, max_len, beam) in (("greedy", (32, 0.9, 8, 100, 0)), ("beam", (32, 0.9, 4, 80, 5)), ("beam", (33, 0.9, 8, 60, 10))):
        m, sd = ref_model(VanillaTransformer, cfg300, seed)
        with torch.no_grad():
            m.next_token_classifier.bias[2] += eos_bias
            m.next_token_classifier.bias[0] -= 5.0
        cid = f"{kind}syn{idx}"
        run(cid, kind, m, syn[:B], max_len, beam)
        cases[-1].update(arch="small", seed=seed, eos_bias=eos_bias, pad_bias=-5.0, checksum=state_dict_checksum(sd), vocab=300,
                         source="synthetic", syn_seed=4)
        idx += 1
    np.savez_compressed(HERE / "standard_decoding.npz", **arrays)
    json.dump(cases, open(HERE / "standard_decoding.json", "w"))
    print("standard written")
