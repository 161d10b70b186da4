#!/bin/bash
for w in 1 2 4 8; do
  for g in 4 16; do
    echo "== D2_WARPS=$w GiB=$g"
    FOURMC_PROFILE=1 FOURMC_D2_WARPS=$w timeout 300 python tools/quick_bench.py $g 1 2>&1 | grep -E "lz4_copy_kernel" | tail -1
  done
done
