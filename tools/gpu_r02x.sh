#!/bin/bash
# r02x (1 GPU): CTAs per SM of the one-warp-per-block D2 (64 registers now), limited by unused dynamic shared memory
mkdir -p gpurun_out
{
for pad in 0 1800 3000 4800 8000; do
FOURMC_D2_PAD=$pad timeout 600 python tools/quick_decode.py 16 1
done
} 2>&1 | grep "copy_kernel\|round trip" | tee gpurun_out/r02x_timing.txt
