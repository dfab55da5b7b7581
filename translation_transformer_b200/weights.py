"""Model configuration and reference-compatible weight containers (host side, plumbing only).

The engine consumes exactly the parameter names of the reference's
``VanillaTransformer.state_dict()`` (/root/reference/src/model/modules.py:39-83);
an optional ``model.`` prefix (Lightning checkpoint, lightning_model.py:73) is
stripped.  ``random_init_state_dict`` draws a random-init model with the same
per-tensor distributions as the reference constructor so that benchmarks and
parity tests can run without checkpoints (there is no network).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import torch


@dataclass(frozen=True)
class ModelConfig:
    """Architecture hyper-parameters (lightning_model.py:28-35 defaults are the small model;
    configs/cfg_standard_product_prediction.yaml is 256/2048/4/4/8)."""
    src_vocab_size: int
    tgt_vocab_size: int
    embedding_dim: int = 256
    feedforward_dim: int = 2048
    num_encoder_layers: int = 4
    num_decoder_layers: int = 4
    num_heads: int = 8
    share_embeddings: bool = True
    src_pad_token_idx: int = 0
    tgt_pad_token_idx: int = 0

    def as_dict(self):
        return asdict(self)


PRODUCT_PREDICTION = dict(embedding_dim=256, feedforward_dim=2048, num_encoder_layers=4,
                          num_decoder_layers=4, num_heads=8, share_embeddings=True)
SINGLE_STEP_RETRO = dict(embedding_dim=256, feedforward_dim=2048, num_encoder_layers=6,
                         num_decoder_layers=6, num_heads=8, share_embeddings=True)


def strip_prefix(state_dict: dict) -> dict:
    out = {}
    for k, v in state_dict.items():
        out[k[6:] if k.startswith("model.") else k] = v
    return out


def infer_config(state_dict: dict, num_heads: int, src_pad: int = 0, tgt_pad: int = 0) -> ModelConfig:
    sd = strip_prefix(state_dict)
    src_e = sd["src_token_featurizer.embedding.weight"]
    tgt_e = sd["tgt_token_featurizer.embedding.weight"]
    n_enc = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("transformer.encoder.layers."))
    n_dec = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("transformer.decoder.layers."))
    ff = sd["transformer.encoder.layers.0.linear1.weight"].shape[0]
    return ModelConfig(src_vocab_size=src_e.shape[0], tgt_vocab_size=tgt_e.shape[0],
                       embedding_dim=src_e.shape[1], feedforward_dim=ff,
                       num_encoder_layers=n_enc, num_decoder_layers=n_dec, num_heads=num_heads,
                       share_embeddings=src_e.data_ptr() == tgt_e.data_ptr() or torch.equal(src_e, tgt_e),
                       src_pad_token_idx=src_pad, tgt_pad_token_idx=tgt_pad)


def random_init_state_dict(cfg: ModelConfig, seed: int) -> dict:
    """Random-init weights: xavier-uniform for every matrix inside nn.Transformer
    (torch's ``Transformer._reset_parameters``), zero attention biases, N(0,1)
    embeddings with a zero pad row, default nn.Linear init for the FFN biases and the
    classifier.  Deterministic for a given torch version (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    E, Fd = cfg.embedding_dim, cfg.feedforward_dim

    def xavier(rows, cols):
        a = math.sqrt(6.0 / (rows + cols))
        return (torch.rand(rows, cols, generator=g) * 2 - 1) * a

    def unif(n, fan_in):
        return (torch.rand(n, generator=g) * 2 - 1) / math.sqrt(fan_in)

    sd = {}
    emb = torch.randn(cfg.src_vocab_size, E, generator=g)
    emb[cfg.src_pad_token_idx] = 0
    sd["src_token_featurizer.embedding.weight"] = emb
    if cfg.share_embeddings:
        assert cfg.src_vocab_size == cfg.tgt_vocab_size
        sd["tgt_token_featurizer.embedding.weight"] = emb
    else:
        e2 = torch.randn(cfg.tgt_vocab_size, E, generator=g)
        e2[cfg.tgt_pad_token_idx] = 0
        sd["tgt_token_featurizer.embedding.weight"] = e2

    def attn(prefix):
        sd[prefix + ".in_proj_weight"] = xavier(3 * E, E)
        sd[prefix + ".in_proj_bias"] = torch.zeros(3 * E)
        sd[prefix + ".out_proj.weight"] = xavier(E, E)
        sd[prefix + ".out_proj.bias"] = torch.zeros(E)

    def ffn_norms(prefix, n_norm):
        sd[prefix + ".linear1.weight"] = xavier(Fd, E)
        sd[prefix + ".linear1.bias"] = unif(Fd, E)
        sd[prefix + ".linear2.weight"] = xavier(E, Fd)
        sd[prefix + ".linear2.bias"] = unif(E, Fd)
        for j in range(1, n_norm + 1):
            sd[f"{prefix}.norm{j}.weight"] = torch.ones(E)
            sd[f"{prefix}.norm{j}.bias"] = torch.zeros(E)

    for i in range(cfg.num_encoder_layers):
        p = f"transformer.encoder.layers.{i}"
        attn(p + ".self_attn")
        ffn_norms(p, 2)
    sd["transformer.encoder.norm.weight"] = torch.ones(E)
    sd["transformer.encoder.norm.bias"] = torch.zeros(E)
    for i in range(cfg.num_decoder_layers):
        p = f"transformer.decoder.layers.{i}"
        attn(p + ".self_attn")
        attn(p + ".multihead_attn")
        ffn_norms(p, 3)
    sd["transformer.decoder.norm.weight"] = torch.ones(E)
    sd["transformer.decoder.norm.bias"] = torch.zeros(E)
    sd["next_token_classifier.weight"] = (torch.rand(cfg.tgt_vocab_size, E, generator=g) * 2 - 1) / math.sqrt(E)
    sd["next_token_classifier.bias"] = unif(cfg.tgt_vocab_size, E)
    return sd


def state_dict_checksum(sd: dict) -> float:
    """Order-independent fingerprint used by the golden fixtures to detect RNG drift."""
    tot = 0.0
    for k in sorted(sd):
        t = sd[k].double()
        w = (torch.arange(t.numel(), dtype=torch.float64) % 7 + 1).reshape(t.shape)
        tot += float((t * w).sum())
    return tot
