"""Loader of the UNMODIFIED reference from oracle/_ref (put there by oracle/make_ref.sh) — test / bench infrastructure.

`bench.py --impl reference` and the `cpu_baseline` leg of bench.py time the reference's own PyTorch implementation
through this module when oracle/_ref exists (kind "reference"); otherwise they fall back to the oracle port
(kind "port").  Nothing under translation_transformer_b200/ imports this file.
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_DIR / "src" / "decoding" / "speculative_decoding.py").exists() and (REF_DIR / "stubs" / "pytorch_lightning").exists()


_loaded = None


def load():
    """(VanillaTransformer, speculative_decoding module, standard_decoding module) of the reference copy."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is missing: run oracle/make_ref.sh in the build container")
    try:
        import pytorch_lightning  # noqa: F401  (a real Lightning wins over the stub)
    except Exception:
        sys.path.insert(0, str(REF_DIR / "stubs"))
    sys.path.insert(0, str(REF_DIR / "src"))
    warnings.filterwarnings("ignore")
    from decoding import speculative_decoding, standard_decoding
    from model.modules import VanillaTransformer
    _loaded = (VanillaTransformer, speculative_decoding, standard_decoding)
    return _loaded


def build_model(cfg, state_dict, device="cpu"):
    """The reference's `VanillaTransformer` with `state_dict` loaded (strict), in eval mode on `device`."""
    VanillaTransformer, _, _ = load()
    m = VanillaTransformer(cfg.src_vocab_size, cfg.tgt_vocab_size, cfg.num_encoder_layers, cfg.num_decoder_layers, cfg.embedding_dim,
                           cfg.num_heads, cfg.feedforward_dim, 0.1, "relu", cfg.share_embeddings, cfg.src_pad_token_idx, cfg.tgt_pad_token_idx)
    m.load_state_dict(state_dict, strict=True)
    return m.eval().to(device)


def greedy_speculative(model, max_len, draft_len, n_drafts, pad, bos, eos, replace):
    _, spec, _ = load()
    return spec.TranslationInferenceGreedySpeculative(model, max_len=max_len, draft_len=draft_len, n_drafts=n_drafts, pad_token=pad,
                                                      bos_token=bos, eos_token=eos, replace_token=replace)


def beam_speculative(model, max_len, n_best, draft_len, n_drafts, vocab, smart, pad, bos, eos, c_token):
    _, spec, _ = load()
    return spec.TranslationInferenceBeamSearchSpeculative(model, max_len=max_len, n_best=n_best, draft_len=draft_len, n_drafts=n_drafts,
                                                          vocab_size=vocab, smart_drafts_mode=smart, pad_token=pad, bos_token=bos,
                                                          eos_token=eos, C_token=c_token)
