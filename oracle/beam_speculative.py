"""Oracle (test infrastructure): speculative beam search, both draft modes, restated.

Reference: /root/reference/src/decoding/speculative_decoding.py:241-598
(`TranslationInferenceBeamSearchSpeculative.generate_trying_all_the_drafts`, `sample`,
`calculate_n_accepted_in_drafts`, `topk_in_each_group`, `mask_with_num_logits_according_nucleus`) and :600-845
(`generate_with_smart_drafts`, `get_vocab_tokens_bool_lib` :402-420).

`smart_drafts_mode=True` differs from "try all the drafts" in where a candidate's drafts come from: a library of
`Ls - 5` windows of `draft_len + 1` source tokens (BOS included) is built once per query; a candidate only tries the
windows whose FIRST token equals its own last token (at most `n_drafts` of them, in library order; window 0 when
there is none) and the remaining `draft_len` tokens of the window are the draft (:690-735).  Candidates therefore
own a different number of decoder rows; the best draft of a candidate is picked by `topk(1)` over its accepted
lengths padded with -1 to the longest group of the iteration (`topk_in_each_group`, :206-223).

Per iteration, for every candidate (beam) of every query:

  1. each of the N source drafts is written into the first `dl` PAD slots of the candidate and the
     decoder is run on the unfinished (candidate, draft) rows (:497-531); finished candidates get
     the artificial "PAD with logit 35" distribution (:466-469);
  2. at each of the dl+1 positions the distribution is truncated to its nucleus (exclusive cumulative
     probability < 0.9975, at most n_best tokens, the best one always kept, :539-541, :871-904);
     a draft token is *accepted* while it lies inside the truncated support (:847-869); the draft with
     the most accepted tokens is chosen per candidate (:553, torch CPU `topk(1)` tie order);
  3. leaves of the candidate's continuation tree: at every position p <= n_accepted each of the
     n_best most probable tokens (except the draft token that continues the accepted path, and except
     logits that are exactly 0.0 because the reference marks pruned entries with 0.0) ends a new
     sequence `accepted draft tokens[0:p] + token` with log-probability
     root + sum(log softmax of the taken tokens) (:294-400);
  4. per query the n_best highest-scoring leaves become the next candidates (:573-576).

The loop ends when every candidate contains EOS or the length budget is exhausted (:464, :586).
Restated per candidate with explicit Python loops; the numerically sensitive parts (softmax, log,
cumulative sums) use the same torch CPU ops as the reference so that scores agree to the last bit.
"""
from __future__ import annotations

import numpy as np
import torch

from .drafting import make_drafts
from .topk_emulation import topk_indices


def truncated_support(logits_row: torch.Tensor, nucleus: float, max_keep: int):
    """Token ids kept by `mask_with_num_logits_according_nucleus` for one distribution
    (speculative_decoding.py:886-899) in descending-logit order."""
    sorted_logits, sorted_idx = torch.sort(logits_row, descending=True)
    cum = torch.cumsum(sorted_logits.softmax(-1), dim=-1)
    keep = [int(sorted_idx[0])]
    for j in range(1, min(max_keep, logits_row.numel())):
        if float(cum[j - 1]) < nucleus:      # exclusive cumulative probability of the better tokens
            keep.append(int(sorted_idx[j]))
    return keep


class BeamSearchSpeculativeOracle:
    def __init__(self, model, max_len: int, n_best: int, draft_len: int, n_drafts: int, vocab_size: int,
                 pad_token: int, bos_token: int, eos_token: int, C_token: int, keep_trace: bool = False,
                 smart_drafts_mode: bool = False):
        self.smart_drafts_mode = smart_drafts_mode
        self.model_input_lines_num = 0
        self.model = model
        self.max_len = max_len
        self.vocab_size = vocab_size
        self.pad, self.bos, self.eos, self.C_token = pad_token, bos_token, eos_token, C_token
        self.n_best = n_best
        self.requested_drafts_num = n_drafts
        self.min_draft_len, self.max_draft_len = 5, 200            # :278-279
        self.draft_len = min(max(self.min_draft_len, draft_len), self.max_draft_len)
        self.accepted_tokens_num = 0
        self.produced_non_pad_tokens = 0
        self.model_calls_num = 0
        self.keep_trace = keep_trace
        self.trace = []

    # -- one decoder call on the unfinished (candidate, draft) rows --------------------------------
    def _logits(self, rows_tokens: np.ndarray, live: np.ndarray, memory, src_pad, row_query: np.ndarray, dl: int,
                first_slot: np.ndarray) -> torch.Tensor:
        """Returns (rows, dl+1, V) logits at positions first_slot-1 .. first_slot+dl-1; finished rows get
        the artificial distribution."""
        R = rows_tokens.shape[0]
        out = torch.zeros(R, dl + 1, self.vocab_size)
        out[:, :, self.pad] = 35.0
        if live.any():
            idx = np.nonzero(live)[0]
            lg = self.model.decode_tgt(torch.from_numpy(rows_tokens[idx]), memory[row_query[idx]], src_pad[row_query[idx]])
            for k, r in enumerate(idx):
                s = int(first_slot[r])
                out[r] = lg[k, s - 1:s + dl]
        return out

    @torch.inference_mode()
    def generate(self, src: torch.Tensor) -> torch.Tensor:
        PAD, EOS, K = self.pad, self.eos, self.n_best
        B = src.shape[0]
        smart = self.smart_drafts_mode
        if smart:
            # library of Ls - 5 windows of draft_len + 1 tokens, BOS column included (:603-615); the first token of a
            # window is its key, the rest the draft
            lib = make_drafts(src.numpy(), self.draft_len + 1, src.shape[1] - 5, self.min_draft_len, self.max_draft_len,
                              EOS, PAD, self.C_token)                                  # (B, n_lib, dl0 + 1)
            drafts_all = lib[:, :, 1:]
            dl = drafts_all.shape[2]
            # per (query, token): the library windows that start with the token, first n_drafts of them, window 0 if none
            by_token = []
            for b in range(B):
                d = {}
                for n in range(lib.shape[1]):
                    d.setdefault(int(lib[b, n, 0]), [])
                    if len(d[int(lib[b, n, 0])]) < self.requested_drafts_num:
                        d[int(lib[b, n, 0])].append(n)
                by_token.append(d)
        else:
            drafts_all = make_drafts(src[:, 1:].numpy(), self.draft_len, self.requested_drafts_num, self.min_draft_len,
                                     self.max_draft_len, EOS, PAD, self.C_token)       # (B, N, dl0)
            dl = drafts_all.shape[2]
        src_pad = src == self.model.src_pad_token_i
        memory = self.model.encode_src(src, src_pad)

        cand = np.full((B, 1), self.bos, dtype=np.int64)     # candidates, query-major
        cand_query = np.arange(B)
        logp = torch.zeros(B, 1)
        empty_cols = 0
        filled = 1                                           # position after the last meaningful token
        budget = self.max_len - filled - 1
        new_cand = cand
        while budget >= 1 and filled <= self.max_len:
            dl = min(budget, dl)
            C = cand.shape[0]
            grow = dl + 1 - empty_cols
            if grow > 0:
                cand = np.concatenate([cand, np.full((C, grow), PAD, dtype=np.int64)], axis=1)
            W = cand.shape[1]
            self.model_calls_num += 1
            # the first `dl` PAD slots of every candidate receive the draft (:497-508)
            slots = np.zeros((C, dl), dtype=np.int64)
            for c in range(C):
                pads = np.nonzero(cand[c] == PAD)[0]
                slots[c] = pads[:dl]
            finished = (cand == EOS).any(axis=1)
            # drafts tried by every candidate (indices into drafts_all[query])
            if smart:
                cand_drafts = []
                for c in range(C):
                    last = int(cand[c, int((cand[c] != PAD).sum()) - 1])      # last meaningful token (:693-697)
                    cand_drafts.append(by_token[int(cand_query[c])].get(last, [0]))
            else:
                cand_drafts = [list(range(drafts_all.shape[1]))] * C
            counts = np.array([len(x) for x in cand_drafts])
            row_base = np.concatenate([[0], np.cumsum(counts)])
            self.model_input_lines_num += int(counts.sum())
            row_cand = np.repeat(np.arange(C), counts)
            rows = cand[row_cand].copy()
            row_query = cand_query[row_cand]
            for c in range(C):
                for j, n in enumerate(cand_drafts[c]):
                    rows[row_base[c] + j, slots[c]] = drafts_all[cand_query[c], n, :dl]
            first_slot = slots[row_cand, 0]
            contiguous = (slots[:, -1] - slots[:, 0] == dl - 1).all()
            if not contiguous:
                raise NotImplementedError("PAD predicted inside a sequence: non-contiguous draft slots")
            logits = self._logits(rows, ~finished[row_cand], memory, src_pad, row_query, dl, first_slot)  # (rows, dl+1, V)

            # accepted length of every draft, best draft per candidate (:539-558 / :769-780)
            longest = int(counts.max())
            n_acc = np.full((C, longest), -1, dtype=np.int64)     # ragged groups are padded with -1 before topk(1)
            for c in range(C):
                for j, n in enumerate(cand_drafts[c]):
                    a = 0
                    while a < dl and int(drafts_all[cand_query[c], n, a]) in truncated_support(logits[row_base[c] + j, a], 0.9975, K):
                        a += 1
                    n_acc[c, j] = a
            pick = np.array([topk_indices(n_acc[c], 1)[0] for c in range(C)])

            # leaves of every candidate's tree (:294-400)
            leaves = []        # (query, score tensor, tokens(np), accepted_count or -1)
            per_query = [[] for _ in range(B)]
            for c in range(C):
                q = int(cand_query[c])
                a = int(n_acc[c, pick[c]])
                n = int(cand_drafts[c][pick[c]])
                lg = logits[row_base[c] + pick[c]]                    # (dl+1, V)
                logprob = lg.softmax(-1).log()
                draft = drafts_all[q, n, :dl].copy()
                if a != dl:
                    draft[a] = self.bos                                # :338-339
                root = logp[c].min()
                s0 = int(slots[c, 0])
                for p in range(a + 1):
                    keep = truncated_support(lg[p], 20.0, K)
                    for tok in sorted(keep):                           # nonzero() order: ascending token id
                        if p < dl and tok == int(draft[p]):
                            continue                                   # continues the accepted path / BOS slot (:341)
                        if float(lg[p, tok]) == 0.0:
                            continue                                   # pruned entries are marked with 0.0 (:320-322, :345)
                        seq_new = np.concatenate([draft[:p], [tok]]).astype(np.int64)
                        contrib = torch.zeros(dl + 1)
                        contrib[:p + 1] = logprob[torch.arange(p + 1), torch.from_numpy(seq_new)]
                        score = root + contrib.cumsum(-1)[-1]
                        toks = cand[c].copy()
                        toks[s0:s0 + dl + 1] = PAD
                        toks[s0:s0 + p + 1] = seq_new
                        per_query[q].append((score, toks, -1 if finished[c] else p))
            new_rows, new_scores, acc_stats = [], [], []
            if min(len(x) for x in per_query) < K:
                raise AssertionError("fewer candidate continuations than n_best (reference topk_in_each_group, :195)")
            longest = max(len(x) for x in per_query)
            for q in range(B):
                scores = torch.stack([s for s, _, _ in per_query[q]])
                # ragged groups are padded with -inf up to the longest group before topk (:206-223)
                padded = np.concatenate([scores.numpy(), np.full(longest - len(per_query[q]), -np.inf, dtype=np.float32)])
                order = topk_indices(padded, K)
                for j in order:
                    new_rows.append(per_query[q][j][1])
                    new_scores.append(per_query[q][j][0])
                    acc_stats.append(per_query[q][j][2])
            new_cand = np.stack(new_rows)
            real = [x for x in acc_stats if x >= 0]
            self.accepted_tokens_num += int(sum(real))
            self.produced_non_pad_tokens += int(sum(real)) + len(real)
            if self.keep_trace:
                self.trace.append({"n_accepted": n_acc.copy(), "pick": pick.copy(), "width": W})
            if (new_cand == EOS).any(axis=1).all():
                break
            cand = new_cand
            cand_query = np.repeat(np.arange(B), K)
            logp = torch.stack(new_scores).reshape(B * K, 1)
            empty_cols = int((cand == PAD).sum(axis=1).min())
            filled = cand.shape[1] - empty_cols
            budget = self.max_len - filled - 1
        return torch.from_numpy(new_cand).reshape(B, K, -1)
