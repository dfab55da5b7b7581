"""SMILES tokenizer (host side) with the reference's vocabulary file format.

Behavioural mirror of /root/reference/src/data_handling/tokenizer_base.py:16-94 and
tokenizer_smiles.py:8-39: fixed service tokens (PAD=0, BOS=1, EOS=2, UNK=3), a
regular-expression atom-level splitter, a frequency-ordered vocabulary, a JSON
vocabulary file that maps *index -> string*, BOS/EOS framing in `encode`, and a
`decode` that drops service tokens and stops at the first EOS.
"""
from __future__ import annotations

import json
import re
from collections import Counter
from pathlib import Path
from typing import Iterable

BOS_TOKEN, EOS_TOKEN, PAD_TOKEN, UNK_TOKEN = "<BOS>", "<EOS>", "<PAD>", "?"

# bracket atoms | two-letter halogens | organic subset | aromatic | bonds, branches, ring closures
_SMILES_ATOMS = re.compile(
    r"(\[[^\]]+]|Br?|Cl?|N|O|S|P|F|I|b|c|n|o|s|p|\(|\)|\.|=|#|-|\+|\\|\/|:|~|@|\?|>|\*|\$|\%[0-9]{2}|[0-9])"
)


def split_smiles(smi: str, check_reconstruction: bool = False) -> list[str]:
    pieces = _SMILES_ATOMS.findall(smi)
    if check_reconstruction and "".join(pieces) != smi:
        raise AssertionError(f"SMILES {smi!r} is not covered by the tokenizer pattern")
    return pieces


class GenericTokenizer:
    def __init__(self, bos_token=BOS_TOKEN, eos_token=EOS_TOKEN, pad_token=PAD_TOKEN, unk_token=UNK_TOKEN):
        self.bos_token, self.eos_token, self.pad_token, self.unk_token = bos_token, eos_token, pad_token, unk_token
        self.pad_token_idx, self.bos_token_idx, self.eos_token_idx, self.unk_token_idx = 0, 1, 2, 3
        self.encoder_dict = {pad_token: 0, bos_token: 1, eos_token: 2, unk_token: 3}
        self.decoder_dict = {i: s for s, i in self.encoder_dict.items()}

    @property
    def n_tokens(self) -> int:
        return len(self.encoder_dict)

    def train_tokenizer(self, train_data: Iterable[str]) -> None:
        raise NotImplementedError

    def encode(self, seq: str) -> list[int]:
        raise NotImplementedError

    def save_vocab(self, voc_save_path) -> None:
        p = Path(voc_save_path).resolve()
        p.parent.mkdir(parents=True, exist_ok=True)
        with p.open("w") as f:
            json.dump(self.decoder_dict, f, sort_keys=True)

    def load_vocab(self, voc_load_path) -> None:
        p = Path(voc_load_path).resolve()
        if not p.exists():
            raise FileNotFoundError
        with p.open() as f:
            self.decoder_dict = {int(k): v for k, v in json.load(f).items()}
        self.encoder_dict = {s: i for i, s in self.decoder_dict.items()}

    def assign_vocab(self, vocab: dict[str, int]) -> None:
        self.encoder_dict = vocab
        self.decoder_dict = {i: s for s, i in vocab.items()}

    def decode(self, tokens: Iterable[int], skip_service_tokens: bool = True) -> str:
        if not skip_service_tokens:
            return "".join(self.decoder_dict[int(i)] for i in tokens)
        service = (self.bos_token_idx, self.eos_token_idx, self.pad_token_idx)
        out = []
        for i in tokens:
            i = int(i)
            if i not in service:
                out.append(self.decoder_dict[i])
            if i == self.eos_token_idx:
                break
        return "".join(out)

    def decode_batch(self, tokens) -> list[str]:
        return [self.decode(row) for row in tokens]


class ChemSMILESTokenizer(GenericTokenizer):
    def train_tokenizer(self, train_data: Iterable[str]) -> None:
        counts = Counter()
        for line in train_data:
            counts.update(split_smiles(line.strip(), check_reconstruction=True))
        for tok, _ in counts.most_common():
            # like the reference, a string that is already present (e.g. "?") is re-assigned
            self.encoder_dict[tok] = len(self.encoder_dict)
        self.decoder_dict = {i: s for s, i in self.encoder_dict.items()}

    def encode(self, seq: str) -> list[int]:
        unk = self.encoder_dict[self.unk_token]
        ids = [self.encoder_dict.get(t, unk) for t in split_smiles(seq)]
        return [self.bos_token_idx, *ids, self.eos_token_idx]
