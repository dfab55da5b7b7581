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


def copy_task_state_dict(cfg: ModelConfig, seed: int, mismatch: float = 0.08, damp: float = 0.05, sharpness: float = 60.0,
                         gain: float = 8.0, alpha: float = 0.05, pad_bias: float = -5.0) -> dict:
    """Synthetic stand-in for a TRAINED reaction model (checkpoints are not available offline): the random-init weights
    of `random_init_state_dict(cfg, seed)` with a deterministic "copy circuit" laid over them, so that the model behaves
    like the trained Molecular Transformer does on USPTO data in the respects that drive the decoding loops:

    * the prediction is (mostly) a copy of the source: position p of the target attends to source position p + 1 through
      a purely positional cross-attention in the last decoder layer (queries = the sinusoidal code rotated by one
      position, keys = the code itself, read from the first `head_dim` embedding dimensions, which the token embedding
      leaves free) and the classifier reads the copied token embedding.  Source windows therefore make good drafts
      (several accepted tokens per decoder call) and a query ends when the copy reaches the source's EOS, i.e. target
      lengths follow the source lengths (ragged finish times, retirement and compaction of the batch);
    * it is imperfect in a deterministic way: a fraction `mismatch` of the vocabulary is predicted as a different token
      (a fixed permutation) and the positional attention is off by one for a few percent of the positions, so drafts
      are rejected at realistic rates; the logit margins stay healthy (like a trained model's), so a bf16 forward pass
      reproduces the fp32 token sequence except at genuine near-ties;
    * every other matrix keeps its random-init values (the sub-layer output projections scaled by `damp`, so that the
      residual stream carries token + position through the stack); the arithmetic per decoder call is exactly that of
      any other weight set of the architecture.

    The state dict loads into the reference's `VanillaTransformer` unchanged (tests/golden/make_golden_bench.py)."""
    sd = {k: v.clone() for k, v in random_init_state_dict(cfg, seed).items()}
    g = torch.Generator().manual_seed(seed + 77)
    E, H, V = cfg.embedding_dim, cfg.num_heads, cfg.tgt_vocab_size
    HD = E // H
    assert cfg.share_embeddings and cfg.src_vocab_size == V and H >= 2 and HD % 2 == 0
    emb = sd["src_token_featurizer.embedding.weight"]
    emb[:, :HD] = 0                                   # dimensions [0, HD) carry the positional code alone
    emb[:, E - 2] = 0                                 # ... and so does dimension E - 2 (see below)
    sd["tgt_token_featurizer.embedding.weight"] = emb
    for k in sd:
        if k.endswith("out_proj.weight") or k.endswith("linear2.weight") or k.endswith("linear2.bias"):
            sd[k] *= damp
    perm = torch.arange(V)                            # token the model emits for a copied source token
    body = torch.arange(4, V)
    n_mis = int(round(mismatch * len(body)))
    if n_mis >= 2:
        pick = body[torch.randperm(len(body), generator=g)[:n_mis]]
        perm[pick] = pick.roll(1)
    Wc = torch.zeros(V, E)
    Wc[perm] = alpha * emb
    sd["next_token_classifier.weight"] = Wc
    sd["next_token_classifier.bias"][cfg.tgt_pad_token_idx] += pad_bias
    sd["next_token_classifier.bias"][1] += 8.0 * pad_bias    # BOS is never predicted (a trained model does not either)
    p = f"transformer.decoder.layers.{cfg.num_decoder_layers - 1}.multihead_attn"
    Win = torch.zeros(3 * E, E)
    omega = torch.exp(torch.arange(0, E, 2).float() * (-math.log(10000.0) / E))[:HD // 2]
    R = torch.zeros(HD, HD)                           # code(p + 1) = R code(p)
    for i in range(HD // 2):
        c, s_ = math.cos(float(omega[i])), math.sin(float(omega[i]))
        R[2 * i, 2 * i], R[2 * i, 2 * i + 1] = c, s_
        R[2 * i + 1, 2 * i], R[2 * i + 1, 2 * i + 1] = -s_, c
    # LayerNorm subtracts the mean mu over all E dimensions from the code (and divides by sigma).  Dimension E - 2 holds
    # sin(position x slowest frequency) ~ 0 and no token embedding, so after LayerNorm it reads -mu / sigma: subtracting
    # it from every code dimension inside the projections restores the plain sinusoidal code (scaled by 1 / sigma)
    Bin = torch.zeros(3 * E)
    Wo = torch.zeros(E, E)
    ones = torch.ones(HD)
    for h in range(H):
        Win[h * HD:(h + 1) * HD, :HD] = sharpness * R
        Win[h * HD:(h + 1) * HD, E - 2] = -sharpness * (R @ ones)
        Win[E + h * HD:E + (h + 1) * HD, :HD] = torch.eye(HD)
        Win[E + h * HD:E + (h + 1) * HD, E - 2] = -ones
        if h < H - 1:                                 # head h carries embedding dimensions [HD (h + 1), HD (h + 2))
            Win[2 * E + h * HD:2 * E + (h + 1) * HD, HD * (h + 1):HD * (h + 2)] = torch.eye(HD)
            Wo[HD * (h + 1):HD * (h + 2), h * HD:(h + 1) * HD] = gain * torch.eye(HD)
    sd[p + ".in_proj_weight"] = Win
    sd[p + ".in_proj_bias"] = Bin
    sd[p + ".out_proj.weight"] = Wo
    return sd


def state_dict_checksum(sd: dict) -> float:
    """Order-independent fingerprint used by the golden fixtures to detect RNG drift."""
    tot = 0.0
    for k in sorted(sd):
        t = sd[k].double()
        w = (torch.arange(t.numel(), dtype=torch.float64) % 7 + 1).reshape(t.shape)
        tot += float((t * w).sum())
    return tot
