#!/bin/bash
# e2e sweep: slice size x pipeline depth of the host-buffer calls (4mc), then the 4mz slice size
for sl in 32 64 128; do for pp in 4; do
  echo "== 4mc FOURMC_PIPE=$pp FOURMC_SLICE_BLOCKS=$sl"
  FOURMC_PIPE=$pp FOURMC_SLICE_BLOCKS=$sl python bench.py --total-gib 16 --batch-gib 16 --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); e=j['e2e']; print('e2e', round(e['value'],2), 'c', round(e['compress_GBps'],1), 'd', round(e['decompress_GBps'],1))
"
done; done
for zs in 192 384 768; do
  echo "== 4mz FOURMC_ZSLICE_BLOCKS=$zs"
  FOURMC_ZSLICE_BLOCKS=$zs python bench.py --codec 4mz --total-gib 16 --batch-gib 16 --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); e=j['e2e']; print('e2e', round(e['value'],2), 'c', round(e['compress_GBps'],1), 'd', round(e['decompress_GBps'],1))
"
done
