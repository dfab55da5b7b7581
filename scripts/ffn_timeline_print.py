"""Print gpurun_out/ffn_timeline.txt (written by a -DTTB_FFN_TIMELINE build, see scripts/build_variant.sh): clock64 stamps of CTA 0
of the feed-forward launch per role (2 = epilogue thread 64, 1 = MMA issuer, 0 = TMA producer), relative to the kernel start."""
import sys
rows = [tuple(map(int, l.split())) for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ffn_timeline.txt")]
t0 = min(r[2] for r in rows)
only = [int(a) for a in sys.argv[2:]] or [2, 1, 0]
for role in only:
    print("role", role)
    prev = None
    for r, i, c in sorted([x for x in rows if x[0] == role], key=lambda x: x[2]):
        print(f"  {i:5d} {c - t0:7d}" + (f"  +{c - prev}" if prev is not None else ""))
        prev = c
