"""Generate the golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [--only model,drafts,greedy,beam,beam_smart,standard]

What it does
  * stubs `pytorch_lightning` (not installed here; only class bases are needed to import
    the reference packages), puts /root/reference/src on sys.path and imports the
    reference's `VanillaTransformer`, `make_drafts` and decoding classes untouched;
  * builds random-init weights with `translation_transformer_b200.weights.random_init_state_dict`
    (deterministic CPU generator) and loads them into the reference model with
    `load_state_dict(strict=True)`, so fixtures only need to store seeds + a checksum;
  * records inputs and reference outputs into small compressed `.npz` / `.json` files.

The reference's decoding loops expose no per-iteration state, so two observation
hooks are installed from the outside (no reference file is edited):
  * `model.decode_tgt` is wrapped to log a SHA-1 of every decoder input token matrix;
  * `torch.Tensor.topk` is wrapped to log the (input, chosen index) of every `topk(1)`
    the decoding loops issue (accepted-token counts per draft and the draft picked).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))

from translation_transformer_b200.weights import ModelConfig, random_init_state_dict, state_dict_checksum  # noqa: E402


def import_reference():
    pl = types.ModuleType("pytorch_lightning")

    class _LM(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    pl.LightningModule = _LM
    pl.LightningDataModule = object
    pl.Callback = object
    pl.Trainer = object
    u = types.ModuleType("pytorch_lightning.utilities")
    ut = types.ModuleType("pytorch_lightning.utilities.types")
    ut.STEP_OUTPUT = object
    cb = types.ModuleType("pytorch_lightning.callbacks")
    cb.BasePredictionWriter = object
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": u,
                        "pytorch_lightning.utilities.types": ut, "pytorch_lightning.callbacks": cb})
    sys.path.insert(0, "/root/reference/src")
    import warnings
    warnings.filterwarnings("ignore")
    from model.modules import VanillaTransformer
    from utils.drafting import make_drafts
    from decoding import speculative_decoding, standard_decoding
    from data_handling.tokenizer_smiles import ChemSMILESTokenizer
    return VanillaTransformer, make_drafts, speculative_decoding, standard_decoding, ChemSMILESTokenizer


def ref_model(VanillaTransformer, cfg: ModelConfig, seed: int):
    sd = random_init_state_dict(cfg, seed)
    m = VanillaTransformer(cfg.src_vocab_size, cfg.tgt_vocab_size, cfg.num_encoder_layers,
                           cfg.num_decoder_layers, cfg.embedding_dim, cfg.num_heads, cfg.feedforward_dim,
                           0.1, "relu", cfg.share_embeddings, cfg.src_pad_token_idx, cfg.tgt_pad_token_idx)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a.astype(np.int64))
    return hashlib.sha1(str(a.shape).encode() + a.tobytes()).hexdigest()


SMALL = dict(embedding_dim=64, feedforward_dim=128, num_encoder_layers=2, num_decoder_layers=2, num_heads=4)
FULL = dict(embedding_dim=256, feedforward_dim=2048, num_encoder_layers=4, num_decoder_layers=4, num_heads=8)


def load_test_sources(Tok):
    src_lines = [l.strip() for l in open(HERE / "product_prediction_src_test.txt") if l.strip()]
    tgt_lines = [l.strip() for l in open(HERE / "product_prediction_tgt_test.txt") if l.strip()]
    tk = Tok()
    tk.train_tokenizer(src_lines + tgt_lines)
    from torch.nn.utils.rnn import pad_sequence
    src = pad_sequence([torch.tensor(tk.encode(l)) for l in src_lines], batch_first=True, padding_value=0).long()
    return tk, src, src_lines, tgt_lines


def synthetic_sources(vocab: int, B: int, lo: int, hi: int, seed: int) -> torch.Tensor:
    """BOS + random body tokens (ids >= 4) + EOS, right-padded with 0."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, hi + 1, (B,), generator=g)
    L = int(lens.max()) + 2
    src = torch.zeros(B, L, dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        src[b, 0] = 1
        src[b, 1:1 + n] = torch.randint(4, vocab, (n,), generator=g)
        src[b, 1 + n] = 2
    return src


# ---------------------------------------------------------------------------------------------
def gen_model(ref):
    VanillaTransformer, _, _, _, Tok = ref
    tk, src, _, _ = load_test_sources(Tok)
    out = {}
    meta = {}
    for name, arch, seed in (("small", SMALL, 11), ("full", FULL, 12)):
        cfg = ModelConfig(src_vocab_size=tk.n_tokens, tgt_vocab_size=tk.n_tokens, **arch)
        m, sd = ref_model(VanillaTransformer, cfg, seed)
        s = src[:4]
        pad = s == 0
        g = torch.Generator().manual_seed(5)
        tgt = torch.randint(4, tk.n_tokens, (4, 19), generator=g)
        tgt[:, 0] = 1
        tgt[1, 12:] = 0
        tgt[3, 7:] = 0
        tgt[2, 5] = 0  # a PAD in the middle: masked as key for later positions
        with torch.inference_mode():
            mem = m.encode_src(s, pad)
            logits = m.decode_tgt(tgt, mem, memory_pad_mask=pad)
            full = m(s, tgt)
        out[f"{name}_src"] = s.numpy()
        out[f"{name}_tgt"] = tgt.numpy()
        out[f"{name}_memory"] = mem.numpy()
        out[f"{name}_logits"] = logits.numpy()
        out[f"{name}_forward_logits"] = full.numpy()
        meta[name] = {"config": cfg.as_dict(), "seed": seed, "checksum": state_dict_checksum(sd)}
    np.savez_compressed(HERE / "model_forward.npz", **out)
    json.dump({"meta": meta, "vocab": tk.encoder_dict}, open(HERE / "model_forward.json", "w"), indent=1)
    print("model_forward written")


def gen_drafts(ref):
    _, make_drafts, _, _, Tok = ref
    tk, src, _, _ = load_test_sources(Tok)
    cases = []
    arrays = {}
    i = 0
    for B in (1, 3, 10):
        for D in (1, 2, 3, 5, 10, 17, 50, 126, 200):
            for N in (1, 2, 3, 10, 23, 25, 100, 200):
                for with_bos in (False, True):
                    s = src[:B] if with_bos else src[:B, 1:]
                    d = make_drafts(s, draft_len=D, n_drafts=N, min_draft_len=1, max_draft_len=200,
                                    eos_token_idx=2, pad_token_idx=0, replace_token_idx=tk.encoder_dict["c"])
                    arrays[f"d{i}"] = d.numpy().astype(np.int16)
                    cases.append(dict(id=i, B=B, D=D, N=N, with_bos=with_bos, min_draft_len=1, max_draft_len=200,
                                      eos=2, pad=0, replace=tk.encoder_dict["c"]))
                    i += 1
    # clamped draft lengths as the beam search uses them (min 5, max 200) and synthetic ragged sources
    syn = synthetic_sources(300, 6, 3, 150, seed=3)
    arrays["syn_src"] = syn.numpy().astype(np.int16)
    for D in (1, 4, 11, 40):
        for N in (1, 2, 7, 23, 60):
            d = make_drafts(syn[:, 1:], D, N, 5, 200, 2, 0, 7)
            arrays[f"d{i}"] = d.numpy().astype(np.int16)
            cases.append(dict(id=i, B=6, D=D, N=N, with_bos=False, min_draft_len=5, max_draft_len=200,
                              eos=2, pad=0, replace=7, synthetic=True))
            i += 1
    np.savez_compressed(HERE / "drafts.npz", **arrays)
    json.dump(cases, open(HERE / "drafts.json", "w"))
    print("drafts written:", i, "cases")


class Hooks:
    """Observation hooks around the untouched reference (see module docstring)."""

    def __init__(self, model):
        self.model = model
        self.calls = []
        self.topk1 = []
        self._orig_decode = model.decode_tgt
        self._orig_topk = torch.Tensor.topk

    def __enter__(self):
        hooks = self

        def decode(tgt, memory, memory_pad_mask, *a, **k):
            hooks.calls.append(sha(tgt.numpy()))
            return hooks._orig_decode(tgt, memory, memory_pad_mask=memory_pad_mask, *a, **k)

        def topk(t, k, *a, **kw):
            r = hooks._orig_topk(t, k, *a, **kw)
            if k == 1 and t.dtype in (torch.int64, torch.int32) and t.dim() == 2:
                hooks.topk1.append((t.numpy().copy(), r.indices.numpy().copy()))
            return r

        self.model.decode_tgt = decode
        torch.Tensor.topk = topk
        return self

    def __exit__(self, *exc):
        self.model.decode_tgt = self._orig_decode
        torch.Tensor.topk = self._orig_topk


def gen_greedy(ref):
    VanillaTransformer, _, spec, _, Tok = ref
    tk, src, _, _ = load_test_sources(Tok)
    cases = []
    arrays = {}

    def run(case_id, model, s, max_len, D, N, replace):
        g = spec.TranslationInferenceGreedySpeculative(model, max_len=max_len, draft_len=D, n_drafts=N,
                                                       pad_token=0, bos_token=1, eos_token=2, replace_token=replace)
        rec = {"id": case_id, "max_len": max_len, "draft_len": D, "n_drafts": N, "replace": replace, "B": int(s.shape[0])}
        with Hooks(model) as h, torch.inference_mode():
            try:
                out = g.generate(s)
                rec["error"] = None
                arrays[f"{case_id}_out"] = out.numpy().astype(np.int16)
            except Exception as e:  # reference failure modes are part of the behaviour
                rec["error"] = type(e).__name__
                rec["error_msg"] = str(e)[:200]
        rec["model_calls"] = g.model_calls_num
        rec["decoder_input_sha1"] = h.calls
        arrays[f"{case_id}_nacc"] = np.concatenate([t[0].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
        arrays[f"{case_id}_pick"] = np.concatenate([t[1].reshape(-1) for t in h.topk1]).astype(np.int16) if h.topk1 else np.zeros(0, np.int16)
        rec["rows_per_iter"] = [int(t[0].shape[0]) for t in h.topk1]
        arrays[f"{case_id}_src"] = s.numpy().astype(np.int16)
        cases.append(rec)
        print(case_id, "calls", g.model_calls_num, "error", rec["error"])

    # (a) small model on the reference's own test sources, several batch / draft settings
    cfg = ModelConfig(src_vocab_size=tk.n_tokens, tgt_vocab_size=tk.n_tokens, **SMALL)
    idx = 0
    for seed in (21, 22):
        m, sd = ref_model(VanillaTransformer, cfg, seed)
        for (B, max_len, D, N) in ((1, 60, 10, 23), (2, 48, 5, 3), (5, 64, 7, 7), (10, 80, 4, 2), (3, 200, 17, 23),
                                   (4, 33, 1, 1), (10, 40, 10, 1)):
            cid = f"small{idx}"
            run(cid, m, src[:B], max_len, D, N, tk.encoder_dict["c"])
            cases[-1].update(arch="small", seed=seed, checksum=state_dict_checksum(sd), vocab=tk.n_tokens, source="test_file")
            idx += 1
    # (b) small models whose classifier bias favours EOS (and suppresses PAD) so that queries
    #     finish at different iterations and leave the batch one by one
    for (seed, eos_bias, B, max_len, D, N) in ((33, 0.5, 10, 100, 6, 5), (46, 0.5, 10, 100, 6, 5), (37, 1.1, 7, 150, 10, 23),
                                               (47, 0.9, 10, 60, 3, 4), (46, 0.7, 10, 100, 10, 23)):
        m, sd = ref_model(VanillaTransformer, cfg, seed)
        with torch.no_grad():
            m.next_token_classifier.bias[2] += eos_bias
            m.next_token_classifier.bias[0] -= 5.0
        cid = f"eos{idx}"
        run(cid, m, src[:B], max_len, D, N, tk.encoder_dict["c"])
        cases[-1].update(arch="small", seed=seed, eos_bias=eos_bias, pad_bias=-5.0, checksum=state_dict_checksum(sd),
                         vocab=tk.n_tokens, source="test_file")
        idx += 1
    # (c) synthetic USPTO-shape sources, vocab 300
    cfg300 = ModelConfig(src_vocab_size=300, tgt_vocab_size=300, **SMALL)
    syn = synthetic_sources(300, 8, 20, 90, seed=4)
    for (seed, eos_bias, B, max_len, D, N) in ((32, 0.9, 8, 100, 6, 5), (33, 0.9, 8, 120, 10, 23), (32, 1.1, 8, 200, 5, 3)):
        m, sd = ref_model(VanillaTransformer, cfg300, seed)
        with torch.no_grad():
            m.next_token_classifier.bias[2] += eos_bias
            m.next_token_classifier.bias[0] -= 5.0
        cid = f"syn{idx}"
        run(cid, m, syn[:B], max_len, D, N, 7)
        cases[-1].update(arch="small", seed=seed, eos_bias=eos_bias, pad_bias=-5.0, checksum=state_dict_checksum(sd),
                         vocab=300, source="synthetic", syn_seed=4)
        idx += 1
    # (d) BASELINE.json configs[0]: full-size random-init model, bs=1, draft_len=10 on the test file
    cfgF = ModelConfig(src_vocab_size=tk.n_tokens, tgt_vocab_size=tk.n_tokens, **FULL)
    m, sd = ref_model(VanillaTransformer, cfgF, 12)
    for b in (0, 1):
        cid = f"full{idx}"
        run(cid, m, src[b:b + 1], 200, 10, 23, tk.encoder_dict["c"])
        cases[-1].update(arch="full", seed=12, checksum=state_dict_checksum(sd), vocab=tk.n_tokens, source="test_file", row=b)
        idx += 1
    np.savez_compressed(HERE / "greedy_speculative.npz", **arrays)
    json.dump(cases, open(HERE / "greedy_speculative.json", "w"))
    print("greedy written")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="model,drafts,greedy")
    args = ap.parse_args()
    torch.set_num_threads(8)
    ref = import_reference()
    todo = args.only.split(",")
    if "model" in todo:
        gen_model(ref)
    if "drafts" in todo:
        gen_drafts(ref)
    if "greedy" in todo:
        gen_greedy(ref)
    if "beam" in todo:
        from make_golden_beam import gen_beam
        gen_beam(ref)
    if "beam_smart" in todo:
        from make_golden_beam import gen_beam_smart
        gen_beam_smart(ref)
    if "standard" in todo:
        from make_golden_standard import gen_standard
        gen_standard(ref)
