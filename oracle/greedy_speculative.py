"""Oracle (test infrastructure): speculative greedy decoding, restated.

Reference: /root/reference/src/decoding/speculative_decoding.py:8-174
(`TranslationInferenceGreedySpeculative.generate`).

Algorithm, per decoding iteration, for every still-active query `b`:

  1. every one of its `N` source-derived drafts (D tokens) is appended to the
     tokens generated so far and the decoder is run on all `Bc*N` rows
     (speculative_decoding.py:97-122);
  2. the argmax predictions at the last generated position and the D draft
     positions are read back (:124-126);
  3. a draft token is accepted while it equals the prediction made one position
     earlier; the draft with the longest accepted prefix wins (:129-133);
  4. the accepted tokens plus one "bonus" prediction are appended (:136-146);
  5. rows that now contain EOS are written to the output and leave the batch
     (:149-168).

Reference quirks that are observable and therefore reproduced:

  * the width of the token matrix is shared by the whole batch and the loop runs
    `while width < max_len` (:93); a query that has not produced EOS when the
    loop ends is returned as an all-PAD row (:87, :174);
  * the width grows by `D + 1 - (#columns that are PAD in every active row)`
    (:97-102; a negative amount truncates, as `torch.nn.functional.pad` does);
  * if a query finishes in an iteration whose width exceeds `max_len`, the
    reference fails with a shape-mismatch RuntimeError (:158); so does the oracle.

Tie-breaking between drafts with the same accepted length: the reference uses
`topk(1)` whose choice among equal values is backend-specific.  The oracle
delegates to `oracle.topk_emulation.topk1_index` which reproduces the torch CPU
behaviour; tied drafts always carry identical accepted tokens, so the returned
sequences do not depend on it.
"""
from __future__ import annotations

import numpy as np
import torch

from .drafting import make_drafts
from .topk_emulation import topk1_index


class GreedySpeculativeOracle:
    def __init__(self, model, max_len: int, draft_len: int, n_drafts: int,
                 pad_token: int, bos_token: int, eos_token: int, replace_token: int,
                 keep_trace: bool = False):
        self.model = model
        self.max_len = max_len
        self.draft_len = draft_len
        self.n_drafts = n_drafts
        self.pad_token, self.bos_token, self.eos_token = pad_token, bos_token, eos_token
        self.replace_token = replace_token
        self.accepted_tokens_num = 0
        self.model_calls_num = 0
        self.keep_trace = keep_trace
        self.trace = []  # one dict per iteration when keep_trace

    def __str__(self):
        return (f"Greedy speculative decoding (draft_len={self.draft_len}, "
                f"n_drafts={self.n_drafts}, max_len={self.max_len})")

    @torch.inference_mode()
    def generate(self, src: torch.Tensor) -> torch.Tensor:
        PAD, EOS = self.pad_token, self.eos_token
        N = self.n_drafts
        B = src.shape[0]
        src_pad = src == self.model.src_pad_token_i
        memory = self.model.encode_src(src, src_pad)
        drafts = make_drafts(src[:, 1:].numpy(), self.draft_len, N, 1, self.max_len, EOS, PAD, self.replace_token)
        D = drafts.shape[2]

        out = np.full((B, self.max_len), PAD, dtype=np.int64)
        active = np.arange(B)                      # original query index of every live row
        G = np.full((B, 1), self.bos_token, dtype=np.int64)
        front = np.zeros(B, dtype=np.int64)        # index of the last generated token per live row

        while G.shape[1] < self.max_len:
            Bc, W = G.shape
            dead_cols = int(((G == PAD).sum(axis=0) == Bc).sum())
            grow = D + 1 - dead_cols
            if grow >= 0:
                Gp = np.concatenate([G, np.full((Bc, grow), PAD, dtype=np.int64)], axis=1)
            else:
                Gp = G[:, :W + grow]
            Wn = Gp.shape[1]

            # decoder input: every live row repeated N times with one draft spliced in
            X = np.repeat(Gp[:, :Wn - 1], N, axis=0)
            for r in range(Bc):
                lo = front[r] + 1
                if lo + D > Wn - 1:
                    raise RuntimeError("index out of bounds while splicing drafts (reference scatter, :111)")
                X[r * N:(r + 1) * N, lo:lo + D] = drafts[active[r]]
            rows = np.repeat(active, N)
            logits = self.model.decode_tgt(torch.from_numpy(X), memory[rows], src_pad[rows])
            self.model_calls_num += 1
            pred = torch.argmax(logits, dim=2).numpy()

            Gn = Gp.copy()
            it_acc, it_pick = [], []
            for r in range(Bc):
                f = int(front[r])
                window = pred[r * N:(r + 1) * N, f:f + D + 1]           # (N, D+1)
                hits = window[:, :D] == drafts[active[r]]
                n_acc = np.where(hits.all(axis=1), D, np.argmin(hits, axis=1))  # leading matches
                pick = topk1_index(n_acc)
                a = int(n_acc[pick])
                Gn[r, f + 1:f + a + 2] = window[pick, :a + 1]
                Gn[r, f + a + 2:f + D + 2] = PAD
                front[r] = f + a + 1
                it_acc.append(a)
                it_pick.append(int(pick))
            self.accepted_tokens_num += int(sum(it_acc))
            if self.keep_trace:
                self.trace.append({"rows": active.tolist(), "n_accepted": it_acc, "draft_index": it_pick, "width": Wn})

            done = (Gn == EOS).any(axis=1)
            if done.any():
                if Wn > self.max_len:
                    raise RuntimeError(
                        f"shape mismatch: finished rows of width {Wn} do not fit max_len {self.max_len} "
                        "(reference speculative_decoding.py:158)")
                out[active[done], :Wn] = Gn[done]
                keep = ~done
                active, Gn, front = active[keep], Gn[keep], front[keep]
            G = Gn
            if active.size == 0:
                break
        return torch.from_numpy(out).unsqueeze(1)
