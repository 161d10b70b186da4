#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for g in 1 16; do timeout 300 python tools/quick_bench.py $g 2 2>&1 | grep -E "compress|equal"; done
FOURMC_PROFILE=1 timeout 300 python tools/quick_bench.py 16 1 2>&1 | grep "profile" | tail -13
for sl in 64 256 512; do
  echo "== FOURMC_SLICE_BLOCKS=$sl"
  FOURMC_SLICE_BLOCKS=$sl python bench.py --total-gib 16 --batch-gib 16 --steps 2 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('value', round(j['value'],1), 'e2e', j['e2e'])
    else: print(l.rstrip())
" | tail -4
done
