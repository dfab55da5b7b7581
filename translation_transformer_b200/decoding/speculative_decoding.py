"""Speculative decoding strategies with the reference's constructor / `generate` interface.

Reference: /root/reference/src/decoding/speculative_decoding.py.  `generate(src)` hands the whole
decoding loop to libttb200: the encoder, draft construction, KV-cached decoder steps, draft
verification, token append, retirement of finished queries and the stop test all run on the
GPU without a host round-trip per step; the host only polls a completion word.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..model import B200Transformer


class TranslationInferenceGreedySpeculative:
    """Mirror of speculative_decoding.py:8-174 (same arguments, same counters, same output)."""

    def __init__(self, model: B200Transformer, max_len: int, draft_len: int, n_drafts: int,
                 pad_token: int, bos_token: int, eos_token: int, replace_token: int,
                 tie_break: str = "torch_cpu", keep_trace: bool = False) -> None:
        self.model = model
        self.max_len = max_len
        self.pad_token, self.bos_token, self.eos_token = pad_token, bos_token, eos_token
        self.replace_token = replace_token
        self.draft_len = draft_len
        self.n_drafts = n_drafts
        self.accepted_tokens_num = 0
        self.produced_tokens_num = 0
        self.model_calls_num = 0
        self.gpu_launches = 0
        self.gpu_ms = 0.0
        self.tie_break = {"torch_cpu": 0, "lowest_index": 1}[tie_break]
        self.keep_trace = keep_trace
        self.trace = []
        self.last_stats = None

    def __str__(self):
        return f"Greedy speculative decoding (draft_len={self.draft_len}, n_drafts={self.n_drafts}, max_len={self.max_len})"

    def generate(self, src: torch.Tensor) -> torch.Tensor:
        """(B, Ls) source token ids -> (B, 1, max_len) predictions, PAD-filled; a query that did not
        reach EOS within the width limit comes back as an all-PAD row, like in the reference."""
        m = self.model
        src_d = src.to(device=m.device, dtype=torch.int64, non_blocking=True).contiguous()
        B, Ls = src_d.shape
        out = torch.empty(B, self.max_len, dtype=torch.int64, device=m.device)
        trace = None
        if self.keep_trace:
            trace = torch.full((self.max_len + 1, B, 4), -1, dtype=torch.int32, device=m.device)
        stats = _lib.GenerateStats()
        with torch.cuda.device(m.device):
            rc = m.lib.ttb_greedy_speculative_generate(
                m._h, src_d.data_ptr(), B, Ls, self.max_len, self.draft_len, self.n_drafts, self.pad_token,
                self.bos_token, self.eos_token, self.replace_token, self.tie_break, out.data_ptr(),
                trace.data_ptr() if trace is not None else None, C.byref(stats),
                torch.cuda.current_stream(m.device).cuda_stream)
        self.last_stats = stats
        self.model_calls_num += stats.model_calls
        self.accepted_tokens_num += stats.accepted_tokens
        self.produced_tokens_num += stats.produced_tokens
        self.gpu_launches += stats.gpu_launches
        self.gpu_ms += stats.gpu_ms
        if trace is not None:
            t = trace[:stats.model_calls].cpu()
            for it in range(t.shape[0]):
                rows = t[it][t[it, :, 0] >= 0]
                self.trace.append({"rows": rows[:, 0].tolist(), "n_accepted": rows[:, 1].tolist(),
                                   "draft_index": rows[:, 2].tolist(), "width": int(rows[0, 3]) if len(rows) else 0})
        _lib.check(rc, "ttb_greedy_speculative_generate")
        out = out.unsqueeze(1)
        return out if src.is_cuda else out.to(src.device)


class TranslationInferenceBeamSearchSpeculative:
    """Mirror of speculative_decoding.py:241-845, both draft modes (`smart_drafts_mode=False`: every candidate tries
    the n_drafts source windows, :428-598; `True`: windows of a draft library keyed by the candidate's last token,
    :600-845).  Same constructor
    arguments and counters (`model_calls_num`, `accepted_tokens_num`, `produced_non_pad_tokens`);
    `generate(src)` returns the (B, n_best, width) hypotheses best-first, like the reference."""

    def __init__(self, model: B200Transformer, max_len: int, n_best: int, draft_len: int, n_drafts: int,
                 vocab_size: int, smart_drafts_mode: bool, pad_token: int, bos_token: int, eos_token: int,
                 C_token: int, tie_break: str = "torch_cpu", keep_trace: bool = False) -> None:
        self.smart_drafts_mode = bool(smart_drafts_mode)
        self.model = model
        self.max_len = max_len
        self.vocab_size = vocab_size
        self.pad_token_idx, self.bos_token_idx, self.eos_token_idx, self.C_token_idx = pad_token, bos_token, eos_token, C_token
        self.n_best = n_best
        self.accepted_tokens_num = 0
        self.model_calls_num = 0
        self.produced_non_pad_tokens = 0
        self.requested_drafts_num = n_drafts
        self.max_draft_len, self.min_draft_len = 200, 5
        drafts_len = min(max(self.min_draft_len, draft_len), self.max_draft_len)
        if drafts_len != draft_len:
            print(f"The draft length should be in range [{self.min_draft_len}: {self.max_draft_len}], so it was changed to {drafts_len}")
        self.draft_len = drafts_len
        self.n_drafts = 0
        self.gpu_launches = 0
        self.gpu_ms = 0.0
        self.tie_break = {"torch_cpu": 0, "lowest_index": 1}[tie_break]
        self.keep_trace = keep_trace
        self.trace = []
        self.last_stats = None

    def __str__(self):
        return (f"SpeculativeSampling decoding (n_best={self.n_best}, max_len={self.max_len}, "
                f"max_num_of_drafts={self.requested_drafts_num}, draft_len={self.draft_len})")

    def generate(self, src: torch.Tensor) -> torch.Tensor:
        m = self.model
        src_d = src.to(device=m.device, dtype=torch.int64, non_blocking=True).contiguous()
        B, Ls = src_d.shape
        K, N = self.n_best, self.requested_drafts_num
        cap = self.max_len + self.draft_len + 4
        out = torch.empty(B * K * cap, dtype=torch.int64, device=m.device)
        t_nacc = t_pick = None
        if self.keep_trace:
            t_nacc = torch.full((self.max_len, B * K, N), -1, dtype=torch.int32, device=m.device)
            t_pick = torch.full((self.max_len, B * K), -1, dtype=torch.int32, device=m.device)
        stats = _lib.GenerateStats()
        width = C.c_int32(0)
        with torch.cuda.device(m.device):
            rc = m.lib.ttb_beam_speculative_generate(
                m._h, src_d.data_ptr(), B, Ls, self.max_len, K, self.draft_len, N, int(self.smart_drafts_mode), self.pad_token_idx, self.bos_token_idx,
                self.eos_token_idx, self.C_token_idx, self.tie_break, out.data_ptr(), C.byref(width),
                t_nacc.data_ptr() if t_nacc is not None else None, t_pick.data_ptr() if t_pick is not None else None,
                C.byref(stats), torch.cuda.current_stream(m.device).cuda_stream)
        self.last_stats = stats
        self.n_drafts += B * N
        self.model_calls_num += stats.model_calls
        self.accepted_tokens_num += stats.accepted_tokens
        self.produced_non_pad_tokens += stats.produced_tokens
        self.gpu_launches += stats.gpu_launches
        self.gpu_ms += stats.gpu_ms
        if t_nacc is not None:
            na, pk = t_nacc[:stats.model_calls].cpu(), t_pick[:stats.model_calls].cpu()
            for it in range(na.shape[0]):
                live = pk[it] >= 0
                self.trace.append({"n_accepted": na[it][live].numpy().astype("int64"), "pick": pk[it][live].numpy().astype("int64")})
        _lib.check(rc, "ttb_beam_speculative_generate")
        W = width.value
        res = out[:B * K * W].view(B, K, W)
        return res if src.is_cuda else res.to(src.device)
