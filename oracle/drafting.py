"""Oracle (test infrastructure): numpy restatement of `make_drafts`.

Reference: /root/reference/src/utils/drafting.py:5-65.

A draft is a window of `D` consecutive source tokens.  Per source row the
reference keeps `N` windows whose start offsets are spread evenly over the
windows that contain no EOS / PAD token (or over the first `N` windows when
there are fewer clean ones), then overwrites any EOS / PAD left inside the kept
windows with `replace_token_idx`.

The start offsets are computed by the reference in float32
(`steps * ((take_from - 1) / max(N - 1, 1))` followed by `.long()`,
drafting.py:58-59); the float32 rounding is part of the observable behaviour
and is reproduced here with explicit np.float32 arithmetic.
"""
from __future__ import annotations

import numpy as np


def draft_start_offsets(n_clean: np.ndarray, n_drafts: int) -> np.ndarray:
    """(B,) clean-window counts -> (B, N) window start offsets (drafting.py:57-59)."""
    take_from = np.maximum(n_clean.astype(np.int64), n_drafts)
    step = (take_from - 1).astype(np.float32) / np.float32(max(n_drafts - 1, 1))
    k = np.arange(n_drafts, dtype=np.int64).astype(np.float32)
    prod = (k[None, :] * step[:, None]).astype(np.float32)
    return prod.astype(np.int64)  # truncation toward zero, values are >= 0


def make_drafts(src, draft_len: int, n_drafts: int, min_draft_len: int, max_draft_len: int,
                eos_token_idx: int, pad_token_idx: int, replace_token_idx: int) -> np.ndarray:
    if not n_drafts > 0:
        raise AssertionError("The number of drafts must be greater than 0")
    if not min_draft_len <= max_draft_len:
        raise AssertionError("The minimum draft length must not be greater than the maximum draft length")
    if pad_token_idx == replace_token_idx:
        raise AssertionError("The pad token and the replace token must be different")
    if eos_token_idx == replace_token_idx:
        raise AssertionError("The eos token and the replace token must be different")
    if eos_token_idx == pad_token_idx:
        raise AssertionError("The eos token and the pad token must be different")

    s = np.asarray(src, dtype=np.int64)
    B, L = s.shape
    N = n_drafts
    D = min(max(min_draft_len, draft_len), max_draft_len)
    extra = N + D - L - 1
    if extra > 0:
        s = np.concatenate([s, np.full((B, extra), pad_token_idx, dtype=np.int64)], axis=1)
    n_win = s.shape[1] - D + 1
    service = (s == eos_token_idx) | (s == pad_token_idx)
    # number of service tokens inside every window via a prefix sum
    csum = np.concatenate([np.zeros((B, 1), dtype=np.int64), np.cumsum(service, axis=1)], axis=1)
    per_window = csum[:, D:D + n_win] - csum[:, 0:n_win]
    n_clean = (per_window == 0).sum(axis=1)
    starts = draft_start_offsets(n_clean, N)
    out = np.empty((B, N, D), dtype=np.int64)
    for b in range(B):
        for n in range(N):
            o = int(starts[b, n])
            out[b, n] = s[b, o:o + D]
    out[(out == eos_token_idx) | (out == pad_token_idx)] = replace_token_idx
    return out
