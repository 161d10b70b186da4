#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pp in 2 3 4; do for sl in 128 256; do
  echo "== FOURMC_PIPE=$pp FOURMC_SLICE_BLOCKS=$sl"
  FOURMC_PIPE=$pp FOURMC_SLICE_BLOCKS=$sl python bench.py --total-gib 16 --batch-gib 16 --steps 2 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); e=j['e2e']; print('e2e', round(e['value'],2), 'c', round(e['compress_GBps'],1), 'd', round(e['decompress_GBps'],1))
    else: print(l.rstrip())
" | tail -3
done; done
