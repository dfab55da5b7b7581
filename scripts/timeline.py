"""In-situ kernel timeline of the decoding loop (CUPTI through torch.profiler): per kernel name the average duration and
the average gap to the end of the preceding kernel, plus the raw events of a few iterations in the middle of a batch."""
import json
import sys
from collections import defaultdict
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402


def main():
    beam = "--beam" in sys.argv
    sys.argv = [a for a in sys.argv if a != "--beam"]
    args = bench.parse()
    from translation_transformer_b200.decoding import TranslationInferenceGreedySpeculative
    from translation_transformer_b200.model import B200Transformer
    cfg, sd = bench.build_weights(args)
    eng = B200Transformer(cfg, sd, precision=args.precision, device=0)
    gen = TranslationInferenceGreedySpeculative(eng, args.max_len, args.draft_len, args.n_drafts, bench.PAD, bench.BOS, bench.EOS, bench.REPLACE)
    nq = args.batch_size
    if beam:   # BASELINE.json configs[2]: speculative beam search bs=4, n_best=5
        from translation_transformer_b200.decoding import TranslationInferenceBeamSearchSpeculative
        gen = TranslationInferenceBeamSearchSpeculative(eng, args.max_len, 5, args.draft_len, args.n_drafts, args.vocab, False,
                                                        bench.PAD, bench.BOS, bench.EOS, bench.REPLACE)
        nq = 4
    dev = torch.device("cuda", 0)
    for i in range(3):
        gen.generate(bench.batch_for(args, 0, i)[:nq].to(dev))
    torch.cuda.synchronize()
    src = bench.batch_for(args, 0, 3)[:nq].to(dev)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        gen.generate(src)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
    print("events", len(ev))
    if not ev:
        return
    t0 = ev[0][0]
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    prev_end = None
    for s, e, n in ev:
        short = n.split("(")[0][-60:]
        a = agg[short]
        a[0] += 1
        a[1] += e - s
        if prev_end is not None:
            a[2] += s - prev_end
        prev_end = max(prev_end or e, e)
    total = ev[-1][1] - t0
    print(f"span {total / 1000:.2f} ms")
    print(f"{'kernel':62s} {'n':>6s} {'avg us':>8s} {'gap us':>8s} {'sum ms':>8s}")
    for k, (n, d, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:62s} {n:6d} {d / n:8.2f} {g / n:8.2f} {d / 1000:8.2f}")
    # idle time: intervals in which no kernel (or copy) is running
    idle, cover_end, idle_by_next = 0.0, ev[0][1], defaultdict(float)
    for st_, en_, n_ in ev[1:]:
        if st_ > cover_end:
            idle += st_ - cover_end
            idle_by_next[n_.split("(")[0][-40:]] += st_ - cover_end
        cover_end = max(cover_end, en_)
    print(f"idle {idle / 1000:.2f} ms of {total / 1000:.2f} ms; before: " + ", ".join(f"{k} {v / 1000:.2f}" for k, v in sorted(idle_by_next.items(), key=lambda kv: -kv[1])[:6]))
    mid = len(ev) // 2
    out = Path("gpurun_out") / "timeline_events.txt"
    with open(out, "w") as f:
        for s, e, n in ev[mid:mid + 80]:
            f.write(f"{s - t0:10.1f} {e - t0:10.1f} {e - s:7.2f} {n.split('(')[0][-50:]}\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
