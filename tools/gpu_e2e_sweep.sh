#!/bin/bash
run() { python bench.py --total-gib 16 --batch-gib 16 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); e=j['e2e']; print('e2e', round(e['value'],2), 'c', round(e['compress_GBps'],1), 'd', round(e['decompress_GBps'],1))
"; }
for pp in 4 6 8; do for sl in 64 128; do echo "== PIPE=$pp SLICE(decompress)=$sl"; FOURMC_PIPE=$pp FOURMC_SLICE_BLOCKS=$sl FOURMC_CSLICE_BLOCKS=32 run; done; done
