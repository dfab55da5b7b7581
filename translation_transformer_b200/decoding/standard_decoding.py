"""Standard (non-speculative) decoding strategies with the reference's constructor / `generate` interface.

Reference: /root/reference/src/decoding/standard_decoding.py.  `TranslationInferenceGreedy.generate` hands the whole
loop to libttb200: it is the KV-cached device loop of the speculative greedy decoding run with one row per query and
no draft tokens (one CUDA-graph replay per generated token, no host round-trip per step).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..model import B200Transformer


class TranslationInferenceGreedy:
    """Mirror of standard_decoding.py:4-56 (same arguments, counters and output)."""

    def __init__(self, model: B200Transformer, max_len: int, pad_token: int, bos_token: int, eos_token: int) -> None:
        self.model = model
        self.max_len = max_len
        self.pad_token, self.bos_token, self.eos_token = pad_token, bos_token, eos_token
        self.model_calls_num = 0
        self.given_tokens = 0
        self.gpu_launches = 0
        self.gpu_ms = 0.0
        self.last_stats = None

    def __str__(self):
        return f"Greedy decoding (max_len={self.max_len})"

    def generate(self, src: torch.Tensor) -> torch.Tensor:
        """(B, Ls) source token ids -> (B, 1, max_len) tokens.  Like the reference, decoding stops at the first step in
        which every row predicts EOS or PAD; rows that finished earlier keep their later predictions."""
        m = self.model
        src_d = src.to(device=m.device, dtype=torch.int64, non_blocking=True).contiguous()
        B, Ls = src_d.shape
        out = torch.empty(B, self.max_len, dtype=torch.int64, device=m.device)
        stats = _lib.GenerateStats()
        with torch.cuda.device(m.device):
            rc = m.lib.ttb_greedy_generate(m._h, src_d.data_ptr(), B, Ls, self.max_len, self.pad_token, self.bos_token,
                                           self.eos_token, out.data_ptr(), C.byref(stats),
                                           torch.cuda.current_stream(m.device).cuda_stream)
        _lib.check(rc, "ttb_greedy_generate")
        self.last_stats = stats
        self.model_calls_num += stats.model_calls
        self.given_tokens += int((src_d != m.src_pad_token_i).sum().item())
        self.gpu_launches += stats.gpu_launches
        self.gpu_ms += stats.gpu_ms
        out = out.unsqueeze(1)
        return out if src.is_cuda else out.to(src.device)


class TranslationInferenceBeamSearch:
    """Mirror of standard_decoding.py:59-174 (same arguments, counters and output): `generate(src)` returns the
    (B, beam_size, width) hypotheses best-first; width is the number of generated columns."""

    def __init__(self, model: B200Transformer, beam_size: int, max_len: int, pad_token: int, bos_token: int, eos_token: int):
        assert max_len > 1
        assert beam_size > 0
        self.model = model
        self.beam_size, self.max_len = beam_size, max_len
        self.pad_token, self.bos_token, self.eos_token = pad_token, bos_token, eos_token
        self.model_calls_num = 0
        self.given_tokens = 0
        self.b_sz = 0
        self.gpu_launches = 0
        self.gpu_ms = 0.0
        self.last_stats = None

    def __str__(self):
        return f"Beam search decoding (beam_size={self.beam_size}, max_len={self.max_len})"

    def generate(self, src: torch.Tensor) -> torch.Tensor:
        m = self.model
        src_d = src.to(device=m.device, dtype=torch.int64, non_blocking=True).contiguous()
        B, Ls = src_d.shape
        out = torch.empty(B * self.beam_size * self.max_len, dtype=torch.int64, device=m.device)
        stats = _lib.GenerateStats()
        width = C.c_int32(0)
        with torch.cuda.device(m.device):
            rc = m.lib.ttb_beam_search_generate(m._h, src_d.data_ptr(), B, Ls, self.max_len, self.beam_size, self.pad_token,
                                                self.bos_token, self.eos_token, out.data_ptr(), C.byref(width), C.byref(stats),
                                                torch.cuda.current_stream(m.device).cuda_stream)
        _lib.check(rc, "ttb_beam_search_generate")
        self.last_stats = stats
        self.model_calls_num += stats.model_calls
        self.given_tokens += int((src_d != m.src_pad_token_i).sum().item())
        self.gpu_launches += stats.gpu_launches
        self.gpu_ms += stats.gpu_ms
        W = width.value
        res = out[:B * self.beam_size * W].view(B, self.beam_size, W)
        return res if src.is_cuda else res.to(src.device)
