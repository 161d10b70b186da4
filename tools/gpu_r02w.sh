#!/bin/bash
# r02w (1 GPU): shared blocks back on the byte-granular D2, one warp per block on the word-stage D2; E1 emit tail without local memory
mkdir -p gpurun_out
{
for g in 16 8 6 4 2 1 0.25 0.015625; do timeout 600 python tools/quick_decode.py $g 2; done
timeout 300 python tools/latency_per_block.py
} 2>&1 | grep "copy_kernel\|decompress:\|region_kernel\|compress:\|per 4 MiB" | tee gpurun_out/r02w_timing.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02w_pytest.txt
cat gpurun_out/r02w_pytest.txt
