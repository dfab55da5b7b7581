"""CPU tests: pin the oracle (`oracle/`) against outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which imports
the reference from /root/reference.  Here nothing of the reference is needed.
"""
import numpy as np
import pytest
import torch

from oracle.drafting import make_drafts
from oracle.greedy_speculative import GreedySpeculativeOracle
from oracle.topk_emulation import topk_indices
from oracle.transformer import OracleTransformer
from translation_transformer_b200.weights import ModelConfig, random_init_state_dict, state_dict_checksum

from helpers import case_weights, load_json, load_npz, sha_tokens, test_file_sources


@pytest.mark.parametrize("name", ["small", "full"])
def test_transformer_forward_matches_reference(name):
    z, meta = load_npz("model_forward.npz"), load_json("model_forward.json")
    m = meta["meta"][name]
    cfg = ModelConfig(**m["config"])
    sd = random_init_state_dict(cfg, m["seed"])
    assert state_dict_checksum(sd) == m["checksum"]
    o = OracleTransformer(sd, cfg.num_heads)
    src, tgt = torch.from_numpy(z[name + "_src"]), torch.from_numpy(z[name + "_tgt"])
    pad = src == 0
    ref_mem = torch.from_numpy(z[name + "_memory"])
    mem = o.encode_src(src, pad)
    # padded source positions are never read (masked keys); the reference's nested-tensor
    # fast path returns zeros there, so compare real positions only
    assert (mem - ref_mem)[~pad].abs().max() < 1e-5
    assert (o.decode_tgt(tgt, ref_mem, pad) - torch.from_numpy(z[name + "_logits"])).abs().max() < 1e-5
    assert (o(src, tgt) - torch.from_numpy(z[name + "_forward_logits"])).abs().max() < 1e-5
    # bf16-contract emulation stays within the bf16 tolerance of the fp32 reference
    ob = OracleTransformer(sd, cfg.num_heads, gemm_dtype="bf16")
    assert torch.allclose(ob(src, tgt), torch.from_numpy(z[name + "_forward_logits"]), atol=1e-2, rtol=1e-2)


def test_make_drafts_matches_reference():
    cases, dz, meta = load_json("drafts.json"), load_npz("drafts.npz"), load_json("model_forward.json")
    _, src, _ = test_file_sources(meta["vocab"])
    src = src.numpy()
    syn = dz["syn_src"].astype(np.int64)
    for c in cases:
        if c.get("synthetic"):
            s = syn[:, 1:]
        else:
            s = src[:c["B"]] if c["with_bos"] else src[:c["B"], 1:]
        d = make_drafts(s, c["D"], c["N"], c["min_draft_len"], c["max_draft_len"], c["eos"], c["pad"], c["replace"])
        assert d.shape == (s.shape[0], c["N"], min(max(c["min_draft_len"], c["D"]), c["max_draft_len"]))
        assert np.array_equal(d, dz[f"d{c['id']}"].astype(np.int64)), c


# the reference's own test of this function (tests/test_drafting.py:19-61): every combination of these sizes on the
# sources of tests/product_prediction_src_test.txt must give drafts of the requested shape
REF_GRID_LENGTHS = [1, 2, 3, 4, 5, 8, 10, 15, 25, 35, 50, 80, 100, 200]
REF_GRID_AMOUNTS = [1, 2, 3, 5, 10, 15, 25, 35, 50, 80, 100, 200]
REF_GRID_BATCHES = [1, 2, 3, 4, 5, 10]


def test_make_drafts_reference_test_grid():
    meta = load_json("model_forward.json")
    tk, src, _ = test_file_sources(meta["vocab"])
    src = src.numpy()
    n = 0
    for B in REF_GRID_BATCHES:
        for n_drafts in REF_GRID_LENGTHS:         # the reference unpacks its product as (batch, n_drafts, draft_len)
            for draft_len in REF_GRID_AMOUNTS:
                d = make_drafts(src[:B], draft_len, n_drafts, 1, 200, tk.eos_token_idx, tk.pad_token_idx, tk.encoder_dict["c"])
                assert d.shape == (min(B, src.shape[0]), n_drafts, draft_len)
                n += 1
    assert n == 1008


def test_make_drafts_argument_checks():
    s = np.array([[5, 6, 7, 2, 0]])
    with pytest.raises(AssertionError):
        make_drafts(s, 2, 0, 1, 10, 2, 0, 5)
    with pytest.raises(AssertionError):
        make_drafts(s, 2, 1, 5, 4, 2, 0, 5)
    with pytest.raises(AssertionError):
        make_drafts(s, 2, 1, 1, 10, 2, 0, 0)
    with pytest.raises(AssertionError):
        make_drafts(s, 2, 1, 1, 10, 2, 0, 2)
    with pytest.raises(AssertionError):
        make_drafts(s, 2, 1, 1, 10, 2, 2, 5)


def test_topk_tie_emulation_matches_torch_cpu():
    g = torch.Generator().manual_seed(0)
    for n in list(range(1, 70)) + [100, 128, 640]:
        for hi in (1, 2, 11):
            for _ in range(8):
                x = torch.randint(0, hi + 1, (n,), generator=g)
                for k in (1, 2, 5):
                    if k <= n:
                        assert x.topk(k).indices.tolist() == topk_indices(x.numpy(), k), (n, k, x.tolist())


def _greedy_cases():
    return [c for c in load_json("greedy_speculative.json")]


@pytest.mark.parametrize("case", _greedy_cases(), ids=lambda c: c["id"])
def test_greedy_speculative_oracle_matches_reference(case):
    if case["arch"] == "full" and case.get("row", 0) != 0:
        pytest.skip("one full-size case is enough for the CPU suite")
    z = load_npz("greedy_speculative.npz")
    cfg, sd = case_weights(case)
    model = OracleTransformer(sd, cfg.num_heads)
    seen = []
    inner = model.decode_tgt

    def spy(tgt, memory, mask):
        seen.append(sha_tokens(tgt.numpy()))
        return inner(tgt, memory, mask)

    model.decode_tgt = spy
    gen = GreedySpeculativeOracle(model, case["max_len"], case["draft_len"], case["n_drafts"], 0, 1, 2,
                                  case["replace"], keep_trace=True)
    src = torch.from_numpy(z[case["id"] + "_src"].astype(np.int64))
    err = None
    try:
        out = gen.generate(src)
    except RuntimeError as e:
        err = type(e).__name__
    assert err == case["error"]
    assert gen.model_calls_num == case["model_calls"]
    assert seen == case["decoder_input_sha1"]          # every decoder input, every iteration
    if err is None:
        assert np.array_equal(out.numpy(), z[case["id"] + "_out"].astype(np.int64))
    # accepted lengths and chosen draft indices of every completed iteration
    ref_nacc = z[case["id"] + "_nacc"].reshape(-1, case["n_drafts"])
    ref_pick = z[case["id"] + "_pick"]
    nacc = np.array([a for t in gen.trace for a in t["n_accepted"]], dtype=np.int64)
    pick = np.array([p for t in gen.trace for p in t["draft_index"]], dtype=np.int64)
    n = len(pick)
    assert n <= len(ref_pick)
    assert np.array_equal(pick, ref_pick[:n])
    assert np.array_equal(nacc, ref_nacc[np.arange(n), ref_pick[:n]])


# ---------------------------------------------------------------------------------------------
def _beam_cases():
    return load_json("beam_speculative.json")


@pytest.mark.parametrize("case", _beam_cases(), ids=lambda c: c["id"])
def test_beam_speculative_oracle_matches_reference(case):
    from oracle.beam_speculative import BeamSearchSpeculativeOracle
    z = load_npz("beam_speculative.npz")
    cfg, sd = case_weights(case)
    model = OracleTransformer(sd, cfg.num_heads)
    seen = []
    inner = model.decode_tgt

    def spy(tgt, memory, mask):
        seen.append(sha_tokens(tgt.numpy()))
        return inner(tgt, memory, mask)

    model.decode_tgt = spy
    gen = BeamSearchSpeculativeOracle(model, case["max_len"], case["n_best"], case["draft_len"], case["n_drafts"],
                                      case["vocab"], 0, 1, 2, case["C_token"], keep_trace=True)
    out = gen.generate(torch.from_numpy(z[case["id"] + "_src"].astype(np.int64)))
    assert case["error"] is None
    assert np.array_equal(out.numpy(), z[case["id"] + "_out"].astype(np.int64))        # all n_best hypotheses, in order
    assert seen == case["decoder_input_sha1"]                                         # every decoder input
    assert (gen.model_calls_num, gen.accepted_tokens_num, gen.produced_non_pad_tokens) == \
        (case["model_calls"], case["accepted_tokens"], case["produced_non_pad_tokens"])
    nacc = np.concatenate([t["n_accepted"].reshape(-1) for t in gen.trace])
    pick = np.concatenate([t["pick"] for t in gen.trace])
    assert np.array_equal(nacc, z[case["id"] + "_nacc"].astype(np.int64))               # accepted lengths of every draft
    assert np.array_equal(pick, z[case["id"] + "_pick"].astype(np.int64))               # chosen draft indices


def _beam_smart_cases():
    return load_json("beam_smart.json")


@pytest.mark.parametrize("case", _beam_smart_cases(), ids=lambda c: c["id"])
def test_beam_speculative_smart_drafts_oracle_matches_reference(case):
    """smart_drafts_mode=True (speculative_decoding.py:600-845)."""
    from oracle.beam_speculative import BeamSearchSpeculativeOracle
    z = load_npz("beam_smart.npz")
    cfg, sd = case_weights(case)
    model = OracleTransformer(sd, cfg.num_heads)
    seen = []
    inner = model.decode_tgt

    def spy(tgt, memory, mask):
        seen.append(sha_tokens(tgt.numpy()))
        return inner(tgt, memory, mask)

    model.decode_tgt = spy
    gen = BeamSearchSpeculativeOracle(model, case["max_len"], case["n_best"], case["draft_len"], case["n_drafts"],
                                      case["vocab"], 0, 1, 2, case["C_token"], keep_trace=True, smart_drafts_mode=True)
    src = torch.from_numpy(z[case["id"] + "_src"].astype(np.int64))
    if case["error"] is not None:
        with pytest.raises(AssertionError):
            gen.generate(src)
        return
    out = gen.generate(src)
    assert np.array_equal(out.numpy(), z[case["id"] + "_out"].astype(np.int64))        # all n_best hypotheses, in order
    assert seen == case["decoder_input_sha1"]                                         # every decoder input
    assert (gen.model_calls_num, gen.accepted_tokens_num, gen.produced_non_pad_tokens, gen.model_input_lines_num) == \
        (case["model_calls"], case["accepted_tokens"], case["produced_non_pad_tokens"], case["model_input_lines_num"])
    nacc = np.concatenate([t["n_accepted"].reshape(-1) for t in gen.trace])
    pick = np.concatenate([t["pick"] for t in gen.trace])
    assert np.array_equal(nacc, z[case["id"] + "_nacc"].astype(np.int64))               # accepted lengths (padded with -1)
    assert np.array_equal(pick, z[case["id"] + "_pick"].astype(np.int64))               # chosen draft of every candidate


# ---------------------------------------------------------------------------------------------
# standard (non-speculative) decoding, standard_decoding.py
def _standard_cases():
    return load_json("standard_decoding.json")


@pytest.mark.parametrize("case", _standard_cases(), ids=lambda c: c["id"])
def test_standard_decoding_oracle_matches_reference_golden(case):
    from oracle.standard_decoding import BeamSearchOracle, GreedyOracle
    z = load_npz("standard_decoding.npz")
    cfg, sd = case_weights(case)
    model = OracleTransformer(sd, cfg.num_heads)
    src = torch.from_numpy(z[case["id"] + "_src"].astype(np.int64))
    if case["kind"] == "greedy":
        o = GreedyOracle(model, case["max_len"], 0, 1, 2)
    else:
        o = BeamSearchOracle(model, case["beam_size"], case["max_len"], 0, 1, 2)
    out = o.generate(src)
    assert np.array_equal(out.numpy(), z[case["id"] + "_out"].astype(np.int64))
    assert o.model_calls_num == case["model_calls"] and o.given_tokens == case["given_tokens"]
    shas = [sha_tokens(t.numpy()) for t in o.decoder_inputs]
    # the reference's first beam-search step goes through model.forward, which the recording hook does not see
    assert (shas if case["kind"] == "greedy" else shas[1:]) == case["decoder_input_sha1"]


def test_oracle_reproduces_the_reference_at_the_benchmarked_greedy_config():
    """tests/golden/bench_configs.*: BASELINE.json configs[1] at full size (256/2048/4+4/8, vocab 288, draft_len 10, n_drafts 23,
    trained-like weights).  Greedy queries are independent of their batch mates (unless the width limit is reached, which it is
    not here), so the oracle decodes the first three queries and must reproduce the reference's rows, accepted lengths and picks
    of those queries out of the 32-query run."""
    import json
    from pathlib import Path
    import numpy as np
    from oracle.greedy_speculative import GreedySpeculativeOracle
    from oracle.transformer import OracleTransformer
    from translation_transformer_b200.synthetic import synthetic_sources
    from translation_transformer_b200.weights import ModelConfig, PRODUCT_PREDICTION, copy_task_state_dict, state_dict_checksum
    golden = Path(__file__).resolve().parent / "golden"
    case = [c for c in json.load(open(golden / "bench_configs.json")) if c["id"] == "cfg1_copy"][0]
    z = np.load(golden / "bench_configs.npz")
    cfg = ModelConfig(src_vocab_size=case["vocab"], tgt_vocab_size=case["vocab"], **PRODUCT_PREDICTION)
    sd = copy_task_state_dict(cfg, case["seed"])
    assert state_dict_checksum(sd) == case["checksum"]
    src = synthetic_sources(32, case["vocab"], seed=case["src_seed"])
    assert np.array_equal(src.numpy(), z["cfg1_copy_src"].astype(np.int64))
    nq = 3
    o = GreedySpeculativeOracle(OracleTransformer(sd, cfg.num_heads), case["max_len"], case["draft_len"], case["n_drafts"], 0, 1, 2, 7, keep_trace=True)
    out = o.generate(src[:nq].clone()).numpy()
    ref = z["cfg1_copy_out"].astype(np.int64)
    assert np.array_equal(out[:, 0], ref[:nq, 0])
    # per-iteration accepted length / pick of the three queries inside the reference's 32-query trace
    ref_nacc = z["cfg1_copy_nacc"].astype(np.int64).reshape(-1, case["n_drafts"])
    ref_pick = z["cfg1_copy_pick"].astype(np.int64)
    ref_acc = ref_nacc[np.arange(len(ref_pick)), ref_pick]
    off = 0
    per_query = {q: [] for q in range(nq)}
    live = list(range(32))
    fin_at = {}
    # reconstruct which query each cell of the reference trace belongs to: live list in order, a query leaves after the iteration
    # in which its row of the output becomes complete (length of its prediction reached)
    lens = (ref[:, 0] != 0).sum(-1)
    prod = {q: 1 for q in range(32)}           # tokens produced so far (BOS counts)
    for n_rows in case["rows_per_iter"]:
        assert n_rows == len(live)
        nxt = []
        for k, q in enumerate(live):
            a, pk = int(ref_acc[off + k]), int(ref_pick[off + k])
            if q < nq:
                per_query[q].append((a, pk))
            prod[q] += a + 1
            if prod[q] < lens[q]:
                nxt.append(q)
        off += n_rows
        live = nxt
    for q in range(nq):
        mine = [(t["n_accepted"][t["rows"].index(q)], t["draft_index"][t["rows"].index(q)]) for t in o.trace if q in t["rows"]]
        assert mine == per_query[q][:len(mine)] and len(mine) == len(per_query[q]), q


@pytest.mark.parametrize("line", [0, 1, 2, 3])
def test_oracle_reproduces_the_reference_at_configs0(line):
    """BASELINE.json configs[0] (greedy speculative, draft_len 10, bs 1, the reference's own test sources) at the full
    product-prediction architecture with trained-like weights: tokens, decoder calls, accepted length and chosen draft of
    every iteration (tests/golden/make_golden_bench.py, cases cfg0_copy_line*)."""
    import json
    from pathlib import Path
    import numpy as np
    import torch
    from oracle.greedy_speculative import GreedySpeculativeOracle
    from oracle.transformer import OracleTransformer
    from translation_transformer_b200.weights import ModelConfig, PRODUCT_PREDICTION, copy_task_state_dict, state_dict_checksum
    golden = Path(__file__).resolve().parent / "golden"
    name = f"cfg0_copy_line{line}"
    case = [c for c in json.load(open(golden / "bench_configs.json")) if c["id"] == name][0]
    z = np.load(golden / "bench_configs.npz")
    cfg = ModelConfig(src_vocab_size=case["vocab"], tgt_vocab_size=case["vocab"], **PRODUCT_PREDICTION)
    sd = copy_task_state_dict(cfg, case["seed"])
    assert state_dict_checksum(sd) == case["checksum"]
    src = torch.from_numpy(z[name + "_src"].astype(np.int64))
    o = GreedySpeculativeOracle(OracleTransformer(sd, cfg.num_heads), case["max_len"], case["draft_len"], case["n_drafts"], 0, 1, 2, 7, keep_trace=True)
    out = o.generate(src.clone()).numpy()
    assert np.array_equal(out, z[name + "_out"].astype(np.int64))
    assert o.model_calls_num == case["model_calls"]
    ref_nacc = z[name + "_nacc"].astype(np.int64).reshape(-1, case["n_drafts"])
    ref_pick = z[name + "_pick"].astype(np.int64)
    assert [t["draft_index"][0] for t in o.trace] == ref_pick.tolist()
    assert [t["n_accepted"][0] for t in o.trace] == ref_nacc[np.arange(len(ref_pick)), ref_pick].tolist()
