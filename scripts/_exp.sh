#!/bin/bash
# one-off A/B harness: kernel check, GPU tests, in-kernel timeline, two bench lines
timeout 200 python scripts/ffn_pair_check.py 2>&1 | tail -2
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
TTB_LIB=libttb200_tl.so timeout 100 python bench.py --no-cpu-baseline --no-insitu --no-extra-workloads --steps 3 --warmup 3 --in-flight 1 2>&1 | tail -c 60
for w in random copy; do timeout 150 python bench.py --no-cpu-baseline --no-extra-workloads --steps 18 --warmup 4 --weights $w 2>/dev/null > gpurun_out/e_$w.json; done
