"""Shared helpers for the parity tests: rebuild the exact weights / inputs of a golden case."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np
import torch

from translation_transformer_b200.weights import ModelConfig, random_init_state_dict, state_dict_checksum

GOLDEN = Path(__file__).resolve().parent / "golden"
SMALL = dict(embedding_dim=64, feedforward_dim=128, num_encoder_layers=2, num_decoder_layers=2, num_heads=4)
FULL = dict(embedding_dim=256, feedforward_dim=2048, num_encoder_layers=4, num_decoder_layers=4, num_heads=8)
ARCH = {"small": SMALL, "full": FULL}


def sha_tokens(a) -> str:
    a = np.ascontiguousarray(np.asarray(a).astype(np.int64))
    return hashlib.sha1(str(a.shape).encode() + a.tobytes()).hexdigest()


def case_weights(case: dict):
    """(cfg, state_dict) of a decoding golden case; verifies the RNG fingerprint."""
    cfg = ModelConfig(src_vocab_size=case["vocab"], tgt_vocab_size=case["vocab"], **ARCH[case["arch"]])
    sd = random_init_state_dict(cfg, case["seed"])
    assert state_dict_checksum(sd) == case["checksum"], "torch CPU RNG drifted: regenerate the golden fixtures"
    sd = {k: v.clone() for k, v in sd.items()}
    if cfg.share_embeddings:
        sd["tgt_token_featurizer.embedding.weight"] = sd["src_token_featurizer.embedding.weight"]
    if "eos_bias" in case:
        sd["next_token_classifier.bias"][2] += case["eos_bias"]
        sd["next_token_classifier.bias"][0] += case.get("pad_bias", 0.0)
    return cfg, sd


def load_json(name):
    return json.load(open(GOLDEN / name))


def load_npz(name):
    return np.load(GOLDEN / name)


def test_file_sources(vocab: dict):
    from torch.nn.utils.rnn import pad_sequence
    from translation_transformer_b200.data_handling import ChemSMILESTokenizer
    tk = ChemSMILESTokenizer()
    tk.assign_vocab(dict(vocab))
    lines = [l.strip() for l in open(GOLDEN / "product_prediction_src_test.txt") if l.strip()]
    src = pad_sequence([torch.tensor(tk.encode(l)) for l in lines], batch_first=True, padding_value=0).long()
    return tk, src, lines


test_file_sources.__test__ = False
