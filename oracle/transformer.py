"""Oracle (test infrastructure): fp32 CPU restatement of the Molecular Transformer forward.

Follows, op for op, the math that the reference obtains from
`torch.nn.Transformer` in `/root/reference/src/model/modules.py`:

  * token embedding + shifted sinusoidal table  -> embeddings.py:7-15, 32-64
    (row 0 of the table is all zeros, real positions start at row 1; no
    sqrt(d_model) scaling of the embedding)
  * post-norm encoder layer (norm_first=False)  -> modules.py:56-68
  * post-norm decoder layer, causal + key-padding masks -> modules.py:69-80, 121-133
  * final LayerNorm of encoder / decoder stacks  -> modules.py:67, 79
  * `next_token_classifier` Linear               -> modules.py:83, 136

Weights are taken from a plain ``{name: tensor}`` dict that uses the key names
of ``VanillaTransformer.state_dict()`` (an optional leading ``model.`` as in a
Lightning checkpoint is stripped).

``gemm_dtype="bf16"`` emulates the precision contract of the CUDA bf16 path
(DESIGN.md "precision contract"): both operands of every Linear are rounded to
bf16, products are accumulated in fp32, Q/K/V, attention outputs and the FFN
hidden activations are stored as bf16, everything else stays fp32.
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

LN_EPS = 1e-5  # modules.py:53


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def sinusoid_table(emb: int, max_len: int = 5000) -> torch.Tensor:
    """(max_len + 1, emb) table; row 0 is zeros (embeddings.py:41-47)."""
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    freq = torch.exp(torch.arange(0, emb, 2).float() * (-math.log(10000.0) / emb))
    tab = torch.zeros(max_len + 1, emb)
    tab[1:, 0::2] = torch.sin(pos * freq)
    tab[1:, 1::2] = torch.cos(pos * freq)
    return tab


class OracleTransformer:
    """Stateless-forward restatement of VanillaTransformer (inference only)."""

    def __init__(self, state_dict: dict, num_heads: int, src_pad: int = 0, tgt_pad: int = 0,
                 gemm_dtype: str = "fp32"):
        sd = {}
        for k, v in state_dict.items():
            k = k[6:] if k.startswith("model.") else k
            sd[k] = v.detach().to(torch.float32).cpu()
        self.sd = sd
        self.heads = num_heads
        self.emb = sd["src_token_featurizer.embedding.weight"].shape[1]
        self.n_enc = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("transformer.encoder.layers."))
        self.n_dec = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("transformer.decoder.layers."))
        self.src_pad_token_i = src_pad
        self.tgt_pad_token_i = tgt_pad
        self.pe = sinusoid_table(self.emb)
        assert gemm_dtype in ("fp32", "bf16")
        self.lowp = gemm_dtype == "bf16"

    # -- primitives ---------------------------------------------------------
    def _lin(self, x, prefix):
        w, b = self.sd[prefix + ".weight"], self.sd[prefix + ".bias"]
        if self.lowp:
            return _bf16_round(x) @ _bf16_round(w).t() + b
        return x @ w.t() + b

    def _lin_slice(self, x, wname, bname, lo, hi):
        w, b = self.sd[wname][lo:hi], self.sd[bname][lo:hi]
        if self.lowp:
            return _bf16_round(x) @ _bf16_round(w).t() + b
        return x @ w.t() + b

    def _store(self, t):
        """Tensors that the bf16 CUDA path keeps in HBM as bf16."""
        return _bf16_round(t) if self.lowp else t

    def _ln(self, x, prefix):
        return F.layer_norm(x, (self.emb,), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], LN_EPS)

    def _attend(self, q, k, v, bias):
        """q (B,Lq,E), k/v (B,Lk,E); bias broadcastable to (B,1,Lq,Lk) additive (0 / -inf)."""
        B, Lq, E = q.shape
        Lk = k.shape[1]
        H, dh = self.heads, E // self.heads
        qh = q.view(B, Lq, H, dh).transpose(1, 2)
        kh = k.view(B, Lk, H, dh).transpose(1, 2)
        vh = v.view(B, Lk, H, dh).transpose(1, 2)
        s = (qh @ kh.transpose(-1, -2)) * (1.0 / math.sqrt(dh)) + bias
        if self.lowp:
            # tensor-core attention of the bf16 path: un-normalised probabilities are rounded to
            # bf16 for the P.V product, the normaliser is summed in fp32 from the unrounded values
            pu = torch.exp(s - s.amax(dim=-1, keepdim=True))
            o = (_bf16_round(pu) @ vh) / pu.sum(dim=-1, keepdim=True)
        else:
            o = torch.softmax(s, dim=-1) @ vh
        o = o.transpose(1, 2).reshape(B, Lq, E)
        return self._store(o)

    def _mha(self, prefix, xq, xkv, bias):
        E = self.emb
        wn, bn = prefix + ".in_proj_weight", prefix + ".in_proj_bias"
        q = self._store(self._lin_slice(xq, wn, bn, 0, E))
        k = self._store(self._lin_slice(xkv, wn, bn, E, 2 * E))
        v = self._store(self._lin_slice(xkv, wn, bn, 2 * E, 3 * E))
        return self._lin(self._attend(q, k, v, bias), prefix + ".out_proj")

    def _ffn(self, prefix, x):
        h = self._store(torch.relu(self._lin(x, prefix + ".linear1")))
        return self._lin(h, prefix + ".linear2")

    def _embed(self, which, tokens):
        L = tokens.shape[1]
        return self.sd[which + ".embedding.weight"][tokens] + self.pe[1:L + 1]

    # -- public API (same names / argument meaning as modules.py:108-137) ----
    def encode_src(self, src: torch.Tensor, src_pad_mask: torch.Tensor) -> torch.Tensor:
        x = self._embed("src_token_featurizer", src)
        bias = torch.zeros(src.shape, dtype=torch.float32).masked_fill(src_pad_mask, float("-inf"))[:, None, None, :]
        for i in range(self.n_enc):
            p = f"transformer.encoder.layers.{i}"
            x = self._ln(x + self._mha(p + ".self_attn", x, x, bias), p + ".norm1")
            x = self._ln(x + self._ffn(p, x), p + ".norm2")
        return self._ln(x, "transformer.encoder.norm")

    def decode_tgt(self, tgt: torch.Tensor, memory: torch.Tensor, memory_pad_mask: torch.Tensor) -> torch.Tensor:
        B, L = tgt.shape
        x = self._embed("tgt_token_featurizer", tgt)
        causal = torch.full((L, L), float("-inf")).triu(1)
        self_bias = causal[None, None] + torch.zeros(B, L).masked_fill(tgt == self.tgt_pad_token_i, float("-inf"))[:, None, None, :]
        mem_bias = torch.zeros(memory_pad_mask.shape, dtype=torch.float32).masked_fill(memory_pad_mask, float("-inf"))[:, None, None, :]
        for i in range(self.n_dec):
            p = f"transformer.decoder.layers.{i}"
            x = self._ln(x + self._mha(p + ".self_attn", x, x, self_bias), p + ".norm1")
            x = self._ln(x + self._mha(p + ".multihead_attn", x, memory, mem_bias), p + ".norm2")
            x = self._ln(x + self._ffn(p, x), p + ".norm3")
        x = self._ln(x, "transformer.decoder.norm")
        return self._lin(x, "next_token_classifier")

    def __call__(self, src, tgt):
        """Full encoder+decoder pass (modules.py:85-106)."""
        pad = src == self.src_pad_token_i
        return self.decode_tgt(tgt, self.encode_src(src, pad), pad)
