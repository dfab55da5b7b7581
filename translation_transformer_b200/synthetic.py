"""Synthetic SMILES-token batches of USPTO-MIT / USPTO-50k shape (there is no dataset offline).

A source row is BOS, `n` body tokens drawn uniformly from the non-service ids [4, vocab), EOS,
right-padded with PAD — exactly what `Seq2SeqDM.collate_fn` (seq2seq_wrappers.py:122-128) hands to
`predict_step`.  Lengths follow a clipped normal that matches the token-length statistics of the
mixed USPTO-MIT sources (mean ~80 tokens, max 200 with service tokens)."""
from __future__ import annotations

import torch


def synthetic_sources(batch_size: int, vocab: int, seed: int, mean_len: float = 80.0, std_len: float = 30.0,
                      min_len: int = 20, max_len: int = 198, pad: int = 0, bos: int = 1, eos: int = 2) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    lens = (torch.randn(batch_size, generator=g) * std_len + mean_len).round().clamp(min_len, max_len).long()
    L = int(lens.max()) + 2
    src = torch.full((batch_size, L), pad, dtype=torch.int64)
    for b in range(batch_size):
        n = int(lens[b])
        src[b, 0] = bos
        src[b, 1:1 + n] = torch.randint(4, vocab, (n,), generator=g)
        src[b, 1 + n] = eos
    return src
