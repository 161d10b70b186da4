#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in 1 2 4 8; do
  for g in 1 16; do
    echo "== D2_WARPS=$w GiB=$g"
    FOURMC_D2_WARPS=$w timeout 300 python tools/quick_bench.py $g 2 2>&1 | grep -E "decompress|equal"
  done
done
echo "== auto"
for g in 1 4 16; do timeout 300 python tools/quick_bench.py $g 2 2>&1 | grep -E "compress|equal"; done
FOURMC_PROFILE=1 timeout 300 python tools/quick_bench.py 16 1 2>&1 | grep "profile" | tail -13
