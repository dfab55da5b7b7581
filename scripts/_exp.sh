#!/bin/bash
# one-off A/B: programmatic dependent launch on/off with several batches in flight
B="python bench.py --no-cpu-baseline --no-insitu --no-extra-workloads --steps 18 --warmup 4"
for env in "" "TTB_NO_PDL=1"; do
  for w in random copy; do
    echo "== $env $w"; env $env timeout 150 $B --weights $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('one_batch_in_flight'))"
  done
done
