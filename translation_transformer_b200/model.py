"""Engine-backed mirror of the reference's `VanillaTransformer` inference interface.

Reference: /root/reference/src/model/modules.py:10-137.  Same method names and argument
meaning (`encode_src(src, src_pad_mask)`, `decode_tgt(tgt, memory, memory_pad_mask)`,
`forward(src, tgt)`, attributes `src_pad_token_i`, `tgt_pad_token_i`, `emb_dim`, ...), so the
decoding strategies and tests read like the reference's.  All compute happens in libttb200
(`csrc/`); torch tensors are only the containers of device memory.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .weights import ModelConfig, strip_prefix


def sinusoid_table(emb: int, max_len: int) -> torch.Tensor:
    """Positional table exactly as embeddings.py:41-47 builds it (row 0 = zeros)."""
    pe = torch.zeros(max_len, emb)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, emb, 2).float() * (-math.log(10000.0) / emb))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return torch.vstack((torch.zeros(1, emb), pe))


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class B200Transformer:
    """Inference-only Molecular Transformer running on libttb200."""

    MAX_POSITIONS = 1024

    def __init__(self, cfg: ModelConfig, state_dict: dict, precision: str = "bf16", device: int | str | torch.device = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("translation_transformer_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.cfg = cfg
        self.precision = precision
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.src_vocab_size, self.tgt_vocab_size = cfg.src_vocab_size, cfg.tgt_vocab_size
        self.src_pad_token_i, self.tgt_pad_token_i = cfg.src_pad_token_idx, cfg.tgt_pad_token_idx
        self.emb_dim, self.num_heads, self.ff_dim = cfg.embedding_dim, cfg.num_heads, cfg.feedforward_dim
        self.num_enc_layers, self.num_dec_layers = cfg.num_encoder_layers, cfg.num_decoder_layers
        desc = _lib.ModelDesc(cfg.src_vocab_size, cfg.tgt_vocab_size, cfg.embedding_dim, cfg.feedforward_dim,
                              cfg.num_encoder_layers, cfg.num_decoder_layers, cfg.num_heads,
                              cfg.src_pad_token_idx, cfg.tgt_pad_token_idx, _lib.PRECISION[precision], self.MAX_POSITIONS)
        handle = C.c_void_p()
        _lib.check(self.lib.ttb_engine_create(C.byref(desc), self.device.index or 0, C.byref(handle)), "ttb_engine_create")
        self._h = handle
        self.load_state_dict(state_dict)

    # -- weights ------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: dict) -> None:
        sd = dict(strip_prefix(state_dict))
        sd["positional_encoding.pe"] = sinusoid_table(self.emb_dim, self.MAX_POSITIONS)
        for name, t in sd.items():
            t = t.detach().to(dtype=torch.float32, device="cpu").contiguous()
            _lib.check(self.lib.ttb_engine_set_param(self._h, name.encode(), t.data_ptr(), t.numel()),
                       f"ttb_engine_set_param({name})")
        _lib.check(self.lib.ttb_engine_finalize(self._h), "ttb_engine_finalize")

    @classmethod
    def from_state_dict(cls, state_dict: dict, num_heads: int, precision: str = "bf16", device=0,
                        src_pad: int = 0, tgt_pad: int = 0) -> "B200Transformer":
        from .weights import infer_config
        return cls(infer_config(state_dict, num_heads, src_pad, tgt_pad), state_dict, precision, device)

    def eval(self):
        return self

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ttb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- forward (modules.py:85-137) -------------------------------------------------------------
    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        return t.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    def encode_src(self, src: torch.Tensor, src_pad_mask: torch.Tensor) -> torch.Tensor:
        src_d = self._dev(src, torch.int64)
        mask_d = self._dev(src_pad_mask, torch.uint8)
        B, Ls = src_d.shape
        mem = torch.empty(B, Ls, self.emb_dim, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.ttb_encode_src(self._h, src_d.data_ptr(), mask_d.data_ptr(), B, Ls, mem.data_ptr(),
                                           _stream_ptr(self.device)), "ttb_encode_src")
        return mem

    def decode_tgt(self, tgt: torch.Tensor, memory: torch.Tensor, memory_pad_mask: torch.Tensor) -> torch.Tensor:
        tgt_d = self._dev(tgt, torch.int64)
        mem_d = self._dev(memory, torch.float32)
        mask_d = self._dev(memory_pad_mask, torch.uint8)
        B, Lt = tgt_d.shape
        Ls = mem_d.shape[1]
        logits = torch.empty(B, Lt, self.tgt_vocab_size, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.ttb_decode_tgt(self._h, tgt_d.data_ptr(), B, Lt, mem_d.data_ptr(), mask_d.data_ptr(), Ls,
                                           logits.data_ptr(), _stream_ptr(self.device)), "ttb_decode_tgt")
        return logits

    def forward(self, src: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
        pad = src == self.src_pad_token_i
        return self.decode_tgt(tgt, self.encode_src(src, pad), pad)

    __call__ = forward
