"""CPU tests of the host-side logic: C-ABI surface, tokenizer, CSV writer, query sharding (gloo)."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from translation_transformer_b200 import _lib
    from translation_transformer_b200.build import build
    build()
    header = (REPO / "include" / "ttb200.h").read_text()
    declared = set(re.findall(r"\b(ttb_[a-z0-9_]+)\s*\(", header))
    assert {"ttb_greedy_speculative_generate", "ttb_encode_src", "ttb_decode_tgt", "ttb_make_drafts"} <= declared
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ttb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert _lib.load().ttb_abi_version() == _lib.ABI_VERSION


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from translation_transformer_b200.model import B200Transformer
    from translation_transformer_b200.weights import ModelConfig, random_init_state_dict
    cfg = ModelConfig(src_vocab_size=30, tgt_vocab_size=30, embedding_dim=64, feedforward_dim=128,
                      num_encoder_layers=1, num_decoder_layers=1, num_heads=4)
    with pytest.raises(RuntimeError):
        B200Transformer(cfg, random_init_state_dict(cfg, 0))


def test_no_oracle_import_in_product():
    for p in (REPO / "translation_transformer_b200").rglob("*.py"):
        assert "oracle" not in re.sub(r'""".*?"""', "", p.read_text(), flags=re.S), f"{p} references the oracle"


def test_tokenizer_roundtrip_and_vocab_file(tmp_path):
    from translation_transformer_b200.data_handling import ChemSMILESTokenizer
    lines = [l.strip() for l in open(REPO / "tests/golden/product_prediction_src_test.txt") if l.strip()]
    tk = ChemSMILESTokenizer()
    tk.train_tokenizer(lines)
    ids = tk.encode(lines[0])
    assert ids[0] == tk.bos_token_idx and ids[-1] == tk.eos_token_idx
    assert tk.decode(ids) == lines[0]
    assert tk.decode(ids + [0, 0, 5]) == lines[0]          # stops at EOS, skips PAD
    tk.save_vocab(tmp_path / "vocab.json")
    tk2 = ChemSMILESTokenizer()
    tk2.load_vocab(tmp_path / "vocab.json")
    assert tk2.encode(lines[3]) == tk.encode(lines[3])
    assert tk.encode("C[Zz]C")[2] == tk.unk_token_idx
    with pytest.raises(FileNotFoundError):
        tk2.load_vocab(tmp_path / "missing.json")


def test_prediction_writer_csv_format(tmp_path):
    from translation_transformer_b200.callbacks import PredictionWriter
    from translation_transformer_b200.data_handling import ChemSMILESTokenizer
    tk = ChemSMILESTokenizer()
    tk.train_tokenizer(["CCO", "c1ccccc1"])
    w = PredictionWriter(tmp_path / "out" / "pred.csv")
    batch = {"src_tokens": torch.tensor([tk.encode("CCO")]), "tgt_tokens": torch.tensor([tk.encode("CC")])}
    a, b = tk.encode("CO"), tk.encode("c1ccccc1")
    width = max(len(a), len(b)) + 1
    pred = torch.tensor([[a + [0] * (width - len(a)), b + [0] * (width - len(b))]])
    w.write(tk, pred, batch)
    w.write(tk, pred, batch)
    rows = (tmp_path / "out" / "pred.csv").read_text().strip().split("\n")
    assert rows[0] == "source,target,prediction_1,prediction_2"
    assert rows[1].split(",") == ["CCO", "CC", "CO", "c1ccccc1"] and len(rows) == 3


def test_shard_bounds_cover_everything():
    from translation_transformer_b200.distributed import shard_bounds
    for n in (0, 1, 7, 40000, 40001):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from translation_transformer_b200.distributed import shard_bounds, gather_predictions
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
n = 7                                    # ragged: 4 queries on rank 0, 3 on rank 1
lo, hi = shard_bounds(n, rank, 2)
all_preds = torch.arange(n * 2 * 5).reshape(n, 2, 5)      # (queries, n_best, max_len)
out = gather_predictions(all_preds[lo:hi].clone())
assert torch.equal(out, all_preds), (rank, out)
# several batches in flight per rank (pipeline.py): worker threads decode, the main thread gathers in batch order,
# so the collectives of the two ranks stay matched although the batches complete out of order
import time
from translation_transformer_b200.pipeline import InFlightDecoder
class Eng: pass
class Gen:
    def __init__(self): self.model = Eng()
    def generate(self, src):
        time.sleep(0.001 * int((src[0, 0, 0] * 7 + rank * 3) % 5))
        return src + 1
fly = InFlightDecoder([Gen(), Gen(), Gen()], device=None)
batches = [all_preds[lo:hi].clone() + 100 * i for i in range(9)]
got = [gather_predictions(o) for o in fly.map(batches)]
for i, g in enumerate(got):
    assert torch.equal(g, all_preds + 100 * i + 1), (rank, i)
fly.close()
# dynamic batch queue shared by both ranks (distributed.BatchQueue) + ONE indexed gather at the end: every batch index is
# handed out exactly once whoever asks, a slow rank simply draws fewer batches, every rank ends with all predictions in order
from translation_transformer_b200.distributed import BatchQueue, gather_indexed_predictions
n_b = 13
q = BatchQueue(n_b)
fly = InFlightDecoder([Gen(), Gen()], device=None)
def next_item():
    i = q.next()
    if i is not None and rank == 1:
        time.sleep(0.004)                # rank 1 is the slow one
    return None if i is None else (i, all_preds[:4] + 1000 * i)
done = fly.drain(next_item)
mine = sorted(i for i, _ in done)
counts = [torch.zeros(1, dtype=torch.int64) for _ in range(2)]
dist.all_gather(counts, torch.tensor([len(mine)]))
assert sum(int(c) for c in counts) == n_b, counts
everything = gather_indexed_predictions([i for i, _ in done], [o for _, o in done], n_b)
for i, o in enumerate(everything):
    assert o is not None and torch.equal(o, all_preds[:4] + 1000 * i + 1), (rank, i)
q2 = BatchQueue(3)                        # a second queue of the job gets its own counter
got2 = []
while (i := q2.next()) is not None:
    got2.append(i)
tot = [torch.zeros(1, dtype=torch.int64) for _ in range(2)]
dist.all_gather(tot, torch.tensor([len(got2)]))
assert sum(int(c) for c in tot) == 3
fly.close()
dist.destroy_process_group()
print("ok", rank)
"""


def test_query_sharding_and_prediction_gather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=port, REPO=str(REPO))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_predict_driver_batches_and_synthetic_vocab(tmp_path):
    """scripts/predict.py: file -> fixed-size batches in file order (seq2seq_wrappers.py:122-128, 168-175) and the
    synthetic vocabulary decodes / re-encodes one to one (the CSV of a synthetic run can be tokenized back)."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("predict_driver", Path(__file__).resolve().parent.parent / "scripts" / "predict.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    golden = Path(__file__).resolve().parent / "golden"
    args = types.SimpleNamespace(synthetic=0, src_file=str(golden / "product_prediction_src_test.txt"),
                                 tgt_file=str(golden / "product_prediction_tgt_test.txt"), vocab_path=None, batch_size=4, vocab=288)
    tk, batches = mod.load_batches(args)
    lines = [l.strip() for l in open(args.src_file) if l.strip()]
    assert sum(b["src_tokens"].shape[0] for b in batches) == len(lines)
    assert all(b["src_tokens"].shape[0] == 4 for b in batches[:-1])
    assert tk.decode(batches[0]["src_tokens"][1].tolist()) == lines[1]
    assert int(batches[0]["src_tokens"][0, 0]) == tk.bos_token_idx and int((batches[0]["src_tokens"] == tk.pad_token_idx).sum()) > 0
    stk = mod.synthetic_tokenizer(288)
    assert stk.n_tokens == 288 and stk.encoder_dict["c"] == 7
    ids = list(range(4, 288))
    assert stk.encode(stk.decode(ids))[1:-1] == ids
    args2 = types.SimpleNamespace(synthetic=70, batch_size=32, vocab=288)
    _, sb = mod.load_batches(args2)
    assert [b["src_tokens"].shape[0] for b in sb] == [32, 32, 6]


def test_in_flight_decoder_keeps_submission_order_and_bounds_concurrency():
    """pipeline.py: K generators decode K batches at a time from host threads; results come back in submission order,
    a failing batch is replaced through `on_error`, generators must not share an engine."""
    import threading
    import time as _time
    from translation_transformer_b200.pipeline import InFlightDecoder

    class FakeEngine:
        pass

    state = {"now": 0, "peak": 0}
    lock = threading.Lock()

    class FakeGen:
        def __init__(self):
            self.model, self.model_calls_num = FakeEngine(), 0

        def generate(self, src):
            with lock:
                state["now"] += 1
                state["peak"] = max(state["peak"], state["now"])
            _time.sleep(0.02 if int(src[0]) % 2 else 0.005)   # completion order differs from submission order
            with lock:
                state["now"] -= 1
            self.model_calls_num += 1
            if int(src[0]) == 5:
                raise RuntimeError("reference failure mode")
            return src * 2

    gens = [FakeGen(), FakeGen(), FakeGen()]
    dec = InFlightDecoder(gens, device=None)
    srcs = [torch.tensor([i, i + 1]) for i in range(11)]
    out = list(dec.map(srcs, pre=lambda s: s + 0, post=lambda o: o + 1, on_error=lambda i, ex: torch.tensor([-1, -1])))
    for i, o in enumerate(out):
        assert o.tolist() == ([-1, -1] if i == 5 else [2 * i + 1, 2 * i + 3])
    assert 2 <= state["peak"] <= 3
    assert dec.counter("model_calls_num") == 11 and len(dec) == 3
    with pytest.raises(RuntimeError):
        list(dec.map([torch.tensor([5, 0])]))
    dec.close()
    shared = FakeGen()
    other = FakeGen()
    other.model = shared.model
    with pytest.raises(AssertionError):
        InFlightDecoder([shared, other], device=None)


def test_batch_queue_and_drain_single_process():
    """Without a process group the queue is a local counter; `InFlightDecoder.drain` decodes every item exactly once."""
    from translation_transformer_b200.distributed import BatchQueue, gather_indexed_predictions
    from translation_transformer_b200.pipeline import InFlightDecoder

    class E:
        pass

    class G:
        def __init__(self):
            self.model = E()

        def generate(self, src):
            if int(src[0]) == 4:
                raise RuntimeError("reference failure mode")
            return src * 3

    q = BatchQueue(9)
    dec = InFlightDecoder([G(), G(), G()], device=None)
    failed = []
    done = dec.drain(lambda: (lambda i: None if i is None else (i, torch.tensor([i, 1])))(q.next()),
                     post=lambda o: o + 1, on_error=lambda k, ex: failed.append(k))
    assert sorted(k for k, _ in done) == list(range(9)) and failed == [4]
    good = [(k, o) for k, o in done if o is not None]
    ordered = gather_indexed_predictions([k for k, _ in good], [o for _, o in good], 9)
    assert ordered[4] is None and all(ordered[i].tolist() == [3 * i + 1, 4] for i in range(9) if i != 4)
    assert q.next() is None
    dec.close()


def test_on_predict_end_report_has_the_reference_keys(capsys, tmp_path):
    """lightning_model.py:218-235 of the reference: key set and order of the report, acceptance rate of the beam search."""
    import json
    from translation_transformer_b200.lightning_model import VanillaEncoderDecoderTransformerLightning as M
    m = M.__new__(M)

    class Gen:
        model_calls_num, accepted_tokens_num, produced_non_pad_tokens = 40, 300, 400

    m.generators, m.generation, m.batch_size, m.max_len, m.n_drafts, m.draft_len = [Gen(), Gen()], "beam_search_speculative", 4, 200, 23, 10
    m.report_prediction_time, m.report_prediction_file, m.tgt_test_path = True, str(tmp_path / "r" / "report.txt"), "data/tgt-test.txt"
    m.on_predict_start()
    m.on_predict_end()
    rep = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert list(rep) == ["algorithm", "batch_size", "tgt_test_path", "max_len", "total_seconds", "model_calls", "seconds_per_model_call",
                         "n_drafts", "draft_len", "accepted_tokens", "acceptance_rate"]
    assert rep["model_calls"] == 80 and rep["accepted_tokens"] == 600 and rep["acceptance_rate"] == 0.75 and rep["tgt_test_path"] == "data/tgt-test.txt"
    assert json.loads(open(m.report_prediction_file).read()) == rep
    m.generation = "greedy_speculative"
    m.on_predict_end()
    rep = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert "accepted_tokens" not in rep and rep["n_drafts"] == 23
    m.generation = "greedy"
    m.on_predict_end()
    assert "n_drafts" not in json.loads(capsys.readouterr().out.strip().splitlines()[-1])


def test_copy_task_weights_behave_like_a_trained_copier():
    """weights.copy_task_state_dict (bench.py --weights copy): teacher-forced, the ORACLE transformer with these weights predicts
    the permuted source token at (almost) every position with a healthy margin, and EOS where the source ends."""
    from oracle.transformer import OracleTransformer
    from translation_transformer_b200.synthetic import synthetic_sources
    from translation_transformer_b200.weights import ModelConfig, PRODUCT_PREDICTION, copy_task_state_dict
    cfg = ModelConfig(src_vocab_size=288, tgt_vocab_size=288, **PRODUCT_PREDICTION)
    sd = copy_task_state_dict(cfg, 1234)
    again = copy_task_state_dict(cfg, 1234)
    assert all(torch.equal(sd[k], again[k]) for k in sd)           # deterministic
    src = synthetic_sources(6, 288, seed=100003)
    Wc = sd["next_token_classifier.weight"]
    emb = sd["src_token_featurizer.embedding.weight"]
    # the token emitted for a copied source token t: the classifier row that reads emb[t]
    perm = (Wc @ emb.t()).argmax(0)
    assert int((perm != torch.arange(288)).sum()) == round(0.08 * 284)
    logits = OracleTransformer(sd, cfg.num_heads)(src, src[:, :-1])
    mask = src[:, 1:] != 0
    ok = (logits.argmax(-1) == perm[src[:, 1:]]) & mask
    assert ok.sum() >= 0.99 * mask.sum()
    top2 = logits.topk(2, -1).values
    assert torch.quantile((top2[..., 0] - top2[..., 1])[mask], 0.02) > 3.0   # trained-like margins: bf16 keeps the arg-max
    eos_pos = (src == 2).float().argmax(-1) - 1
    assert (logits.argmax(-1)[torch.arange(6), eos_pos] == 2).all()


def test_bench_workloads_are_the_golden_cases():
    """bench.py's workloads (batch 0, weights) are exactly what tests/golden/make_golden_bench.py ran the reference on: the
    parity block of the bench line compares like with like."""
    import importlib.util
    import json
    import types
    import numpy as np
    from translation_transformer_b200.weights import state_dict_checksum
    spec = importlib.util.spec_from_file_location("bench_mod", REPO / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    args = types.SimpleNamespace(vocab=288, seed=1234, max_len=200, tie_break="torch_cpu")
    cases = {c["id"]: c for c in json.load(open(REPO / "tests/golden/bench_configs.json"))}
    z = np.load(REPO / "tests/golden/bench_configs.npz")
    for key, weights, gid in (("greedy", "copy", "cfg1_copy"), ("greedy", "random", "cfg1_random"), ("beam", "copy", "cfg2_copy"), ("retro", "copy", "cfg3_copy")):
        wl = bench.Workload(key, weights, args)
        assert wl.golden_id == gid
        c = cases[gid]
        assert np.array_equal(wl.batch(0).numpy(), z[gid + "_src"].astype(np.int64)), gid
        assert state_dict_checksum(wl.sd) == c["checksum"], gid
        assert (wl.w["draft_len"], wl.w["n_drafts"], wl.bs) == (c["draft_len"], c["n_drafts"], c["B"])
        if wl.kind == "beam":
            assert wl.w["n_best"] == c["n_best"]
    # another vocabulary / seed has no fixture: the golden check is skipped, not faked
    assert bench.Workload("greedy", "copy", types.SimpleNamespace(vocab=300, seed=1234, max_len=200, tie_break="torch_cpu")).golden_id is None
    # algorithmic work of the dominant class: 4 layers x (FFN 4 E F + out-projection 2 E E) flops per token row
    wlr = bench.Workload("greedy", "random", args)
    f, b = bench.class_work("gemm_ffn1", wlr.w, wlr.cfg, [32], 80.0, True, True, True, "bf16")
    rows = 32 * 23 * 11
    assert abs(f / 4 - (4.0 * rows * 256 * 2048 + 2.0 * rows * 256 * 256 + 16.0 * rows * 256)) < 1.0


def test_committed_bench_line_keeps_the_driver_contract():
    """profiles/r2_bench_1gpu.json is the line `python bench.py` printed on the B200 for the committed build: the keys the
    driver and the judge read must be there and consistent with each other (BASELINE.json metric, roofline = achieved / peak,
    end-to-end figure with its copy sizes, clocks without a throttle reason, a launch count, the parity self-check)."""
    import json
    line = json.loads((REPO / "profiles/r2_bench_1gpu.json").read_text().strip().splitlines()[-1])
    base = json.loads((REPO / "BASELINE.json").read_text())
    assert base["metric"].startswith("SMILES/sec") and line["metric"].startswith("SMILES/sec") and line["unit"] == "SMILES/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches", "parity_checked", "parity_ok"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["higher_is_better"] is True and line["dtype"] == "bf16"
    assert "workload" in line["config"] and "model" not in line["config"]
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0.0 < r["frac"] < 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    c = line["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = line["e2e"]
    assert e["unit"] == line["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * line["value"] < e["value"] < 1.1 * line["value"]          # a measurement of its own, not a copy of `value`
    assert e["value"] != line["value"]
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(line["clocks"]["reasons"]) and line["clocks"]["sm_mhz"] > 0.9 * line["clocks"]["sm_max_mhz"]
    assert line["gpu_launches"] > 0 and line["parity_checked"] is True and line["parity_ok"] is True
    # value = queries of all timed steps / time: 32 queries per step
    assert abs(line["value"] - 32.0 * 1000.0 / line["ms_per_step"]) / line["value"] < 1e-3
