#!/bin/bash
# r02q (1 GPU): D2 variants at 16 GiB (W = 1): literal words 3 / 5 unrolled / loop, tokens ahead, match loads early, 28 / 32 CTAs per SM
mkdir -p gpurun_out
{
for v in 0 16 8 24 9 10 26 4 20 1 2; do
FOURMC_D2_VAR=$v timeout 600 python tools/quick_decode.py 16 1
done
} 2>&1 | grep "copy_kernel\|round trip" | tee gpurun_out/r02q_timing.txt
