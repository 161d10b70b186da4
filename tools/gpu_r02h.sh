#!/bin/bash
# r02h (1 GPU): two warps per block in D2 at the batch sizes of 2 / 4 / 8 ranks (2048 / 1024 / 512 blocks)
mkdir -p gpurun_out
{
for g in 8 4 2; do
  for w in 0 2 4; do
    FOURMC_D2_WARPS=$w timeout 300 python tools/quick_decode.py $g 1 2>&1 | grep "decompress\|copy_kernel"
  done
done
} | tee gpurun_out/r02h_timing.txt
