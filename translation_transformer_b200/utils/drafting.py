"""`make_drafts` with the reference's signature (utils/drafting.py:5-65), executed by the
draft-construction kernel of libttb200 (csrc/drafting.cu)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def make_drafts(src: torch.Tensor, draft_len: int, n_drafts: int, min_draft_len: int, max_draft_len: int,
                eos_token_idx: int, pad_token_idx: int, replace_token_idx: int) -> torch.Tensor:
    """(B, L) source token ids -> (B, N, D) drafts; same assertions as the reference."""
    assert n_drafts > 0, "The number of drafts must be greater than 0"
    assert min_draft_len <= max_draft_len, "The minimum draft length must not be greater than the maximum draft length"
    assert pad_token_idx != replace_token_idx, "The pad token and the replace token must be different"
    assert eos_token_idx != replace_token_idx, "The eos token and the replace token must be different"
    assert eos_token_idx != pad_token_idx, "The eos token and the pad token must be different"
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("make_drafts runs on the GPU only (no CPU path)")
    dev = src.device if src.is_cuda else torch.device("cuda", torch.cuda.current_device())
    s = src.to(device=dev, dtype=torch.int64)
    if s.stride(1) != 1:
        s = s.contiguous()
    B, L = s.shape
    D = min(max(min_draft_len, draft_len), max_draft_len)
    out = torch.empty(B, n_drafts, D, dtype=torch.int64, device=dev)
    d_out = C.c_int32(0)
    with torch.cuda.device(dev):
        _lib.check(lib.ttb_make_drafts(s.data_ptr(), s.stride(0), B, L, draft_len, n_drafts, min_draft_len, max_draft_len,
                                       eos_token_idx, pad_token_idx, replace_token_idx, out.data_ptr(), C.byref(d_out),
                                       torch.cuda.current_stream(dev).cuda_stream), "ttb_make_drafts")
    assert d_out.value == D
    return out if src.is_cuda else out.to(src.device)
