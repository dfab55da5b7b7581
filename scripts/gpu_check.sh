#!/bin/bash
# usage: scripts/gpu_check.sh TAG  -> GPU parity tests + bench; writes gpurun_out/bench_TAG.json
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo rc $?
tail -c 600 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), "calls", d["decoder_calls"], "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
for k,v in d["kernel_shares"].items(): print(f"  {k:18s} {v['ms']:8.3f} ms {v['launches']:5d}  {1000*v['ms']/v['launches']:7.2f} us")
PY
