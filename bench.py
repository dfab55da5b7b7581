#!/usr/bin/env python
"""Benchmark of the hot path: speculative greedy decoding, product-prediction Molecular Transformer.

    python bench.py --gpus N --steps K --warmup W            (our arm: libttb200 on B200)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm: CPU port of the reference)

One "step" = one batch of `--batch-size` synthetic USPTO-MIT-shape queries decoded to completion
through `TranslationInferenceGreedySpeculative.generate` (BASELINE.json configs[1]).  The K timed steps are
decoded with `--in-flight` (default 3) batches at a time per GPU, each on its own engine and stream
(pipeline.py; same predictions, batches are independent); the strictly sequential figure is reported
beside it as `one_batch_in_flight`.  Multi-GPU runs
shard the queries (each rank decodes its own batches, weak scaling) and all-gather the predictions.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

from translation_transformer_b200.synthetic import synthetic_sources  # noqa: E402
from translation_transformer_b200.weights import ModelConfig, PRODUCT_PREDICTION, random_init_state_dict  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures under profiles/
# (kernel class -> bytes); classes without a capture report null
NCU_TRAFFIC = {"gemm_ffn1": 14.76e6, "gemm_self_out": 12.76e6, "self_attn": 13.50e6, "cross_attn": 7.03e6,
               "gemm_qkv": 4.57e6, "gemm_classifier": 4.32e6}   # profiles/r1k_top_kernels_ncu_full.txt (cold-cache replays: mostly the weight fetch)

PAD, BOS, EOS, REPLACE = 0, 1, 2, 7   # REPLACE plays the role of the "c" token (lightning_model.py:117)
METRIC = "SMILES/sec (greedy speculative, product prediction)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=18)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--draft-len", type=int, default=10)
    ap.add_argument("--n-drafts", type=int, default=23)
    ap.add_argument("--max-len", type=int, default=200)
    ap.add_argument("--vocab", type=int, default=288)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--eos-bias", type=float, default=0.0, help="added to the classifier bias of EOS (0 = plain random init)")
    ap.add_argument("--pad-bias", type=float, default=0.0)
    ap.add_argument("--cpu-queries", type=int, default=2, help="queries per CPU-baseline sample (about 6 s of host time each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-insitu", action="store_true", help="skip the extra CUPTI-profiled step (kernels_in_situ)")
    ap.add_argument("--tie-break", default="torch_cpu", choices=["torch_cpu", "lowest_index"])
    ap.add_argument("--in-flight", type=int, default=3,
                    help="bs=32 batches decoded concurrently per GPU, one engine + stream each (translation_transformer_b200/pipeline.py); "
                         "1 = strictly one batch after the other, also always measured and reported as `one_batch_in_flight`")
    ap.add_argument("--clock-period-ms", type=int, default=200, help="nvidia-smi sampling period; 0 disables the sampler")
    return ap.parse_args()


def build_weights(args):
    cfg = ModelConfig(src_vocab_size=args.vocab, tgt_vocab_size=args.vocab, **PRODUCT_PREDICTION)
    sd = {k: v.clone() for k, v in random_init_state_dict(cfg, args.seed).items()}
    sd["tgt_token_featurizer.embedding.weight"] = sd["src_token_featurizer.embedding.weight"]
    sd["next_token_classifier.bias"][EOS] += args.eos_bias
    sd["next_token_classifier.bias"][PAD] += args.pad_bias
    return cfg, sd


def workload_name(args):
    return (f"product-prediction greedy speculative bs={args.batch_size} draft_len={args.draft_len} "
            f"n_drafts={args.n_drafts} max_len={args.max_len}; Molecular Transformer 256/2048/4+4/8 random-init "
            f"(seed {args.seed}, eos_bias {args.eos_bias}, pad_bias {args.pad_bias}), vocab {args.vocab}; "
            f"synthetic USPTO-MIT-shape sources (20..198 tokens, mean 80)")


def batch_for(args, rank, step):
    """Synthetic batch of step `step`.  Weak scaling: per-GPU work is fixed as N grows, so every rank decodes the SAME
    queries of the step (rotated by its rank); with rank-specific random batches the step time would be the slowest
    rank's batch, i.e. the maximum of N samples of the batch-to-batch spread (about +-12 % with these sources), which
    measures the data rather than the system."""
    return torch.roll(synthetic_sources(args.batch_size, args.vocab, seed=100003 + step), shifts=rank, dims=0)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=200):
        self.index, self.proc, self.lines, self.period_ms = index, None, [], period_ms

    def start(self):
        if self.period_ms <= 0:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def class_work(name, args, cfg, hist, src_lens_mean, fused_ln=True, fused_ffn=True, chained_ffn=False):
    """Algorithmic (flops, bytes) of ALL launches of a kernel class over the iterations in `hist`
    (live queries per iteration); per-unit figures are in DESIGN.md §4.  With the fused kernels the
    sub-layer tails (bias + residual + LayerNorm, fp32 and bf16 copies of the residual stream) are part
    of the GEMM classes and the whole feed-forward block is the class `gemm_ffn1`."""
    E, F, V, L = cfg.embedding_dim, cfg.feedforward_dim, cfg.tgt_vocab_size, cfg.num_decoder_layers
    per_q = args.n_drafts * (args.draft_len + 1)
    rows = sum(h * per_q for h in hist)           # token rows summed over iterations
    ab = 2 if args.precision == "bf16" else 4      # activation bytes
    n_it = len(hist)
    ln_flops = 8.0 * rows * E
    stream_bytes = rows * E * (4 + 4 + ab)          # residual in, residual out (fp32), low-precision copy out
    if name == "gemm_ffn1" and fused_ffn:
        f, b = (4.0 * rows * E * F + ln_flops) * L, (rows * E * ab + stream_bytes) * L + 2 * E * F * ab * L * n_it
        if chained_ffn:   # the launch also computes the cross-attention out-projection + LayerNorm (att in, x in/out fp32)
            f += (2.0 * rows * E * E + ln_flops) * L
            b += (rows * E * ab + rows * E * 8) * L + E * E * ab * L * n_it
        return f, b
    if name in ("gemm_self_out", "gemm_cross_out") and fused_ln:
        return (2.0 * rows * E * E + ln_flops) * L, (rows * E * ab + stream_bytes) * L + E * E * ab * L * n_it
    if name == "gemm_ffn2" and fused_ln:
        return (2.0 * rows * E * F + ln_flops) * L, (rows * F * ab + stream_bytes) * L + E * F * ab * L * n_it
    gemm = {"gemm_qkv": (E, 3 * E, ab), "gemm_self_out": (E, E, 4), "gemm_cross_q": (E, E, ab), "gemm_cross_out": (E, E, 4),
            "gemm_ffn1": (E, F, ab), "gemm_ffn2": (F, E, 4)}
    if name in gemm:
        K, N, ob = gemm[name]
        return 2.0 * rows * K * N * L, (rows * K * ab + rows * N * ob) * L + K * N * ab * L * n_it
    if name == "gemm_classifier":
        return 2.0 * rows * E * V, rows * E * ab + rows * V * 4 + E * V * ab * n_it
    if name == "cross_attn":
        lk = src_lens_mean
        return 4.0 * rows * lk * E * L, (rows * E * ab * 2 + sum(hist) * lk * 2 * E * ab) * L
    if name == "self_attn":
        # keys: accepted prefix (grows ~1 token/iteration on average) + causal half of the draft row
        flops = sum(4.0 * h * per_q * (it + 1 + (args.draft_len + 2) / 2.0) * E for it, h in enumerate(hist)) * L
        byts = sum(h * (per_q * 4 * E * ab + (it + 1) * 2 * E * ab) for it, h in enumerate(hist)) * L
        return flops, byts
    if name == "add_layernorm":
        return 8.0 * rows * E * 3 * L, rows * E * (4 + 4 + 4 + ab) * 3 * L
    return 0.0, 0.0


def insitu_kernel_times(fn):
    """One extra step under CUPTI (torch.profiler): per kernel name the number of launches, the average duration and the
    average time between the end of the preceding kernel and its own end, i.e. what the kernel adds to the step inside
    the CUDA-graph / programmatic-dependent-launch pipeline (the event brackets of the instrumented step cannot see
    that: they serialise every launch)."""
    from collections import defaultdict
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                 if e.device_type == torch.autograd.DeviceType.CUDA and not e.name.startswith("Mem")), key=lambda t: t[0])
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    prev_end = None
    for st, en, name in ev:
        a = agg[name.split("(")[0].split("::")[-1].split("<")[0]]
        a[0] += 1
        a[1] += en - st
        a[2] += en - max(st, prev_end) if prev_end is not None and prev_end > st else en - st
        prev_end = en if prev_end is None else max(prev_end, en)
    return {k: {"launches": n, "avg_us": round(d / n, 2), "avg_added_us": round(g / n, 2)} for k, (n, d, g) in agg.items()}


CLASS_KERNEL = {"gemm_ffn1": ("ffn_pair_kernel", "ffn_fused_kernel"), "self_attn": ("attn_mma_kernel",), "cross_attn": ("attn_mma_kernel",),
                "gemm_qkv": ("gemm_pair_k256_kernel", "gemm_bf16_tc_persistent_kernel"), "gemm_self_out": ("gemm_resid_ln_kernel",),
                "gemm_cross_out": ("gemm_resid_ln_kernel",)}


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Reference arm: the CPU port of the reference's algorithm (oracle/), all host threads.
    Each step decodes a bounded sample (`--cpu-queries` queries) of the step's batch."""
    if rank != 0:
        return
    from oracle.greedy_speculative import GreedySpeculativeOracle
    from oracle.transformer import OracleTransformer
    # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which would make this a one-thread baseline
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(os.cpu_count() or 1)
    cfg, sd = build_weights(args)
    model = OracleTransformer(sd, cfg.num_heads)
    nq = args.cpu_queries if args.warmup + args.steps <= 8 else 1    # ~6.5 s of host time per query: keep the run to minutes
    times = []
    for i in range(args.warmup + args.steps):
        src = batch_for(args, 0, i)[:nq]
        gen = GreedySpeculativeOracle(model, args.max_len, args.draft_len, args.n_drafts, PAD, BOS, EOS, REPLACE)
        t0 = time.perf_counter()
        gen.generate(src)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = nq * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "SMILES/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": f"{nq} of {args.batch_size} queries per step"},
            "cpu_baseline": {"value": value, "unit": "SMILES/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{nq} query per step, {len(times)} steps, full decode (max_len {args.max_len})"},
            "e2e": {"value": value, "unit": "SMILES/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from translation_transformer_b200 import _lib
    from translation_transformer_b200.decoding import TranslationInferenceGreedySpeculative
    from translation_transformer_b200.distributed import gather_predictions
    from translation_transformer_b200.model import B200Transformer
    from translation_transformer_b200.pipeline import InFlightDecoder

    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL logs to stdout (its version banner at VERSION and at WARN level)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=dev)
    cfg, sd = build_weights(args)
    n_fly = max(1, args.in_flight)
    engs = [B200Transformer(cfg, sd, precision=args.precision, device=local_rank) for _ in range(n_fly)]
    gens = [TranslationInferenceGreedySpeculative(e, args.max_len, args.draft_len, args.n_drafts, PAD, BOS, EOS, REPLACE,
                                                  tie_break=args.tie_break) for e in engs]
    eng, gen = engs[0], gens[0]          # the instrumented / profiled steps run on the first engine alone
    fly = InFlightDecoder(gens, device=local_rank) if n_fly > 1 else None
    lib = eng.lib
    n_total = args.warmup + args.steps + 1
    host = [batch_for(args, rank, i).pin_memory() for i in range(n_total)]
    devb = [h.to(dev) for h in host]
    out_host = torch.empty(args.batch_size, 1, args.max_len, dtype=torch.int64).pin_memory()
    out_hosts = [torch.empty_like(out_host).pin_memory() for _ in range(args.steps)] if fly else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    errors = []

    def one_step(i, e2e, g=None):
        flush.zero_()
        src = host[i].to(dev, non_blocking=True) if e2e else devb[i]
        try:
            out = (g or gen).generate(src)
        except RuntimeError as ex:   # reference-faithful failure modes (see oracle/greedy_speculative.py)
            errors.append(str(ex)[:80])
            out = torch.zeros(args.batch_size, 1, args.max_len, dtype=torch.int64, device=dev)
        if world > 1:
            gather_predictions(out, counts=[args.batch_size] * world)   # NCCL all-gather of the predictions
        if e2e:
            out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pre_resident(i):
        flush.zero_()                                   # on the worker's stream, like everything of its step
        return devb[i]

    def pre_host(i):
        flush.zero_()
        return host[i].to(dev, non_blocking=True)

    def timed(e2e, first, pipelined=True):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        if fly and pipelined:
            # all K steps are submitted at once; `n_fly` of them are decoded at any time, each on its own engine and stream
            # (flush, input copy, decoding loop, result copy); this thread collects them in step order
            futs = [fly.submit(first + k, pre=pre_host if e2e else pre_resident,
                               post=(lambda o, k=k: (out_hosts[k].copy_(o, non_blocking=True), o)[1]) if e2e else None) for k in range(args.steps)]
            for f in futs:
                try:
                    out = f.result()
                except RuntimeError as ex:   # reference-faithful failure modes (see oracle/greedy_speculative.py)
                    errors.append(str(ex)[:80])
                    out = torch.zeros(args.batch_size, 1, args.max_len, dtype=torch.int64, device=dev)
                if world > 1:
                    gather_predictions(out, counts=[args.batch_size] * world)   # NCCL all-gather of the predictions
        else:
            for k in range(args.steps):
                one_step(first + k, e2e)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (also sizes every workspace) -------------------------------------------------
    for g in gens:
        for i in range(args.warmup):
            one_step(i, False, g)
    if fly:   # and W more through the worker threads, `n_fly` at a time (thread start-up, per-thread CUDA state)
        for f in [fly.submit(i % args.warmup, pre=pre_resident) for i in range(args.warmup * n_fly)]:
            try:
                f.result()
            except RuntimeError:
                pass
    # ---- one instrumented step: CUDA-event time of every kernel class -> dominant kernel ---------
    n_cls = lib.ttb_kernel_class_count()
    names = [lib.ttb_kernel_class_name(i).decode() for i in range(n_cls)]
    lib.ttb_engine_set_profiling(eng._h, (1 << n_cls) - 1)
    one_step(args.warmup + args.steps, False)
    ms_arr, n_arr = (C.c_double * n_cls)(), (C.c_int64 * n_cls)()
    lib.ttb_engine_get_profile(eng._h, n_cls, ms_arr, n_arr)
    shares = {names[i]: {"ms": round(ms_arr[i], 3), "launches": int(n_arr[i])} for i in range(n_cls) if n_arr[i]}
    tot_ms = sum(v["ms"] for v in shares.values()) or 1.0
    for v in shares.values():
        v["share"] = round(v["ms"] / tot_ms, 4)
    dominant = max((n for n in shares if n not in ("encoder", "misc")), key=lambda n: shares[n]["ms"])
    dom_id = names.index(dominant)

    # per-iteration live-query history of the instrumented step (for the algorithmic work of the kernel)
    buf = (C.c_int32 * (args.max_len + 2))()
    n_hist = lib.ttb_engine_get_history(eng._h, buf, args.max_len + 2)
    hists = [list(buf[:n_hist])]
    dom_ms, dom_launches = ms_arr[dom_id], int(n_arr[dom_id])
    lib.ttb_engine_set_profiling(eng._h, 0)

    # ---- timed region 1: inputs resident in HBM, no instrumentation -------------------------------
    def counters():
        return [sum(getattr(g, n) for g in gens) for n in ("model_calls_num", "gpu_launches", "accepted_tokens_num", "produced_tokens_num")]

    one_ms = timed(False, args.warmup, pipelined=False) if fly else None    # strictly one batch after the other
    calls0, launches0, acc0, tok0 = counters()
    sampler = ClockSampler(local_rank, args.clock_period_ms)
    if rank == 0:
        sampler.start()
    timed_ms = timed(False, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    calls1, launches1, acc1, tok1 = counters()
    calls, launches, accepted, produced = calls1 - calls0, launches1 - launches0, acc1 - acc0, tok1 - tok0

    # ---- timed region 2: end to end through the public API with host buffers ---------------------
    e2e_ms = timed(True, args.warmup)

    insitu = None
    if rank == 0 and not args.no_insitu:
        try:
            insitu = insitu_kernel_times(lambda: gen.generate(devb[args.warmup]))   # rank-local: no collective in here
        except Exception as ex:   # CUPTI not available: the event-bracket figures stand alone
            insitu = {"error": str(ex)[:120]}

    lt = torch.tensor([launches, calls, accepted, produced], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
    launches, calls, accepted, produced = [int(x) for x in lt.tolist()]

    if rank == 0:
        queries = world * args.steps * args.batch_size
        value = queries / (timed_ms / 1000.0)
        e2e_value = queries / (e2e_ms / 1000.0)
        peaks = measured_peaks()
        src_lens_mean = float((host[args.warmup + args.steps] != PAD).sum().item()) / args.batch_size
        flops = byts = 0.0
        fused_ln = "add_layernorm" not in shares
        fused_ffn = "gemm_ffn2" not in shares
        chained_ffn = fused_ffn and "gemm_cross_out" not in shares
        for h in hists:
            f, b = class_work(dominant, args, cfg, h, src_lens_mean, fused_ln, fused_ffn, chained_ffn)
            flops += f
            byts += b
        intensity = flops / max(byts, 1.0)
        balance = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
        dom_s = max(dom_ms, 1e-9) / 1000.0
        if intensity >= balance:
            roof = {"bound": "tensor", "achieved": flops / dom_s / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": byts / dom_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
        label = {"gemm_ffn1": ("ffn_pair_kernel (cross out-proj+LayerNorm2 + FFN1+ReLU+FFN2+residual+LayerNorm3, cta_group::2)" if chained_ffn else
                               "ffn_fused_kernel (FFN1+ReLU+FFN2+residual+LayerNorm)") if fused_ffn else "gemm_ffn1",
                 "gemm_self_out": "gemm_resid_ln_kernel (self out-proj)" if fused_ln else "gemm_self_out",
                 "gemm_cross_out": "gemm_resid_ln_kernel (cross out-proj)" if fused_ln else "gemm_cross_out"}.get(dominant, dominant)
        overhead_us = 1000.0 * shares["misc"]["ms"] / shares["misc"]["launches"] if "misc" in shares else None
        roof.update({"frac": roof["achieved"] / roof["peak"], "traffic": NCU_TRAFFIC.get(dominant), "kernel": label, "kernel_class": dominant,
                     "event_bracket_us_of_a_trivial_kernel": overhead_us,
                     "launches": dom_launches, "avg_launch_us": 1000.0 * dom_ms / max(dom_launches, 1),
                     "algorithmic_flops_per_launch": flops / max(dom_launches, 1), "algorithmic_bytes_per_launch": byts / max(dom_launches, 1),
                     "peak_source": peaks["source"], "share_of_step": shares[dominant]["share"],
                     "timing": "CUDA events around every launch of the class on the launching stream, one extra "
                               "instrumented step of the same workload right before the timed region "
                               "(bracketing every launch inside the timed region costs ~2x step time)"})
        if insitu and "error" not in insitu:
            k = next((n for n in CLASS_KERNEL.get(dominant, ()) if n in insitu), None)
            if k:
                us = insitu[k]["avg_added_us"]
                per_launch = (flops if roof["bound"] == "tensor" else byts) / max(dom_launches, 1)
                ach = per_launch / (us * 1e-6) / (1e12 if roof["bound"] == "tensor" else 1e9)
                roof["in_situ"] = {"kernel": k, "launches": insitu[k]["launches"], "avg_us": insitu[k]["avg_us"], "avg_added_us": us,
                                   "achieved": ach, "frac": ach / roof["peak"],
                                   "how": "CUPTI activity records of one extra step run like the timed ones (CUDA graph + programmatic "
                                          "dependent launch): avg_added_us = end of the kernel minus end of the preceding kernel"}
        line = {"metric": METRIC, "value": value, "unit": "SMILES/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": timed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": workload_name(args), "global_batch": world * args.batch_size,
                           "parallelism": f"dp{world} (one batch per rank and step: the step's queries rotated by the rank, so "
                                          f"per-GPU work is identical; predictions all-gathered over NCCL)" if world > 1 else "single GPU",
                           "batches_in_flight": n_fly,
                           "l2": "256 MiB buffer written before every step (L2 flush)"},
                "e2e": {"value": e2e_value, "unit": "SMILES/s", "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": int(host[args.warmup].numel() * 8),
                        "d2h_bytes_per_step": int(out_host.numel() * 8)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernel_shares": shares,
                "decoder_calls": calls, "accepted_tokens_per_call": accepted / max(calls, 1),
                "produced_tokens": produced, "reference_failures": errors[:3]}
        if one_ms is not None:
            line["one_batch_in_flight"] = {"value": queries / (one_ms / 1000.0), "unit": "SMILES/s", "ms_per_step": one_ms / args.steps,
                                           "what": "the same K steps strictly one after the other on one engine (the kernel-level "
                                                   "figures of `roofline` / `kernels_in_situ` are measured in this mode)"}
        if insitu:
            line["kernels_in_situ"] = insitu
        if world == 1:
            # BASELINE.json configs[2] beside the headline (same weights): speculative beam search bs=4, n_best=5
            try:
                from translation_transformer_b200.decoding import TranslationInferenceBeamSearchSpeculative
                bgens = [TranslationInferenceBeamSearchSpeculative(e_, args.max_len, 5, args.draft_len, args.n_drafts, args.vocab, False,
                                                                   PAD, BOS, EOS, REPLACE) for e_ in engs]
                bgen = bgens[0]
                bsrc = [batch_for(args, 0, i)[:4].to(dev) for i in range(3)]
                bgen.generate(bsrc[0])
                torch.cuda.synchronize()
                c0, t0 = bgen.model_calls_num, time.perf_counter()
                for b in bsrc[1:]:
                    bgen.generate(b)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                line["beam_speculative"] = {"workload": "product prediction beam-search speculative bs=4 n_best=5 draft_len=10 n_drafts=23 "
                                                        "(BASELINE.json configs[2]), same weights and sources, KV-cached",
                                            "value": 8 / dt, "unit": "SMILES/s", "ms_per_batch": 1000 * dt / 2,
                                            "decoder_calls_per_batch": (bgen.model_calls_num - c0) / 2, "batches_in_flight": 1}
                if n_fly > 1:    # the same search with `n_fly` batches in flight (one engine each)
                    bfly = InFlightDecoder(bgens, device=local_rank)
                    many = [batch_for(args, 0, i % 3)[:4] for i in range(3 * n_fly)]
                    list(bfly.map(many[:n_fly], pre=lambda t: t.to(dev), on_error=lambda i, ex: None))
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    list(bfly.map(many, pre=lambda t: t.to(dev), on_error=lambda i, ex: None))
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    line["beam_speculative"]["in_flight"] = {"batches_in_flight": n_fly, "value": 4 * len(many) / dt, "unit": "SMILES/s",
                                                             "ms_per_batch": 1000 * dt / len(many)}
                    bfly.close()
            except (RuntimeError, AssertionError) as ex:   # reference-faithful failure modes
                line["beam_speculative"] = {"error": str(ex)[:120]}
        if world == 1 and not args.no_cpu_baseline:
            from oracle.greedy_speculative import GreedySpeculativeOracle
            from oracle.transformer import OracleTransformer
            nq = args.cpu_queries
            o = GreedySpeculativeOracle(OracleTransformer(sd, cfg.num_heads), args.max_len, args.draft_len, args.n_drafts, PAD, BOS, EOS, REPLACE)
            t0 = time.perf_counter()
            try:
                o.generate(host[args.warmup][:nq].clone())
            except RuntimeError:
                pass
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nq / dt, "unit": "SMILES/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"first {nq} queries of the first timed batch, full decode, {o.model_calls_num} decoder calls, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
