#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell tensor-core / tensor-memory / TMA mnemonics
(B200_PROFILING.md "What proves a Blackwell-native kernel") for every kernel of libttb200.so.

    python scripts/sass_summary.py [lib.so] > profiles/sass_summary.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,
HMMA = legacy mma.sync path, LDGSTS = cp.async, FFMA2 / FADD2 = packed fp32 (two lanes per instruction)."""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else REPO / "translation_transformer_b200" / "libttb200.so"
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU.EX2", "FFMA2", "FADD2"]

sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", sass)
if names:
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out)) if len(out) == len(names) else {}

rows = []
for block in sass.split("Function : ")[1:]:
    name, _, body = block.partition("\n")
    c = Counter()
    n_inst = 0
    for line in body.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n_inst += 1
        op = m.group(1)
        for k in MNEMONICS:
            if k == "UTCHMMA.2CTA":
                c[k] += op.startswith("UTCHMMA") and ".2CTA" in op
            elif k == "UTCHMMA":
                c[k] += op.startswith("UTCHMMA") or (op.startswith("UTC") and "MMA" in op)
            else:
                c[k] += op.startswith(k)
    pretty = demangle.get(name.strip(), name.strip())
    if pretty.endswith(")"):                       # drop the parameter list (the last top-level parenthesis group)
        depth = 0
        for i in range(len(pretty) - 1, -1, -1):
            depth += pretty[i] == ")"
            depth -= pretty[i] == "("
            if depth == 0:
                pretty = pretty[:i]
                break
    pretty = pretty.replace("void ", "").replace("ttb::", "").replace("(int)", "").replace("__nv_bfloat16", "bf16")
    rows.append((pretty, n_inst, c))

arch = re.search(r"arch = (sm_\w+)", sass)
print(f"# {lib.name}: {len(rows)} kernels, {arch.group(1) if arch else '?'}; columns = instruction counts in the SASS of each kernel")
print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{k:>12s}" for k in MNEMONICS))
for pretty, n_inst, c in sorted(rows, key=lambda r: -r[2]["UTCHMMA"] * 100000 - r[1]):
    print(f"{pretty[:78]:78s} {n_inst:7d} " + " ".join(f"{c[k]:12d}" for k in MNEMONICS))
tot = Counter()
for _, _, c in rows:
    tot.update(c)
print(f"{'TOTAL':78s} {sum(r[1] for r in rows):7d} " + " ".join(f"{tot[k]:12d}" for k in MNEMONICS))
