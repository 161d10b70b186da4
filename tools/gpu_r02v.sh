#!/bin/bash
# r02v (1 GPU): D2 warps per block against the batch size (the choice in dec_blocks), cleaned-up build; GPU suite
mkdir -p gpurun_out
{
for g in 8 4 2 1; do
for w in 1 2 4 8; do
FOURMC_D2_WARPS=$w timeout 600 python tools/quick_decode.py $g 2
done; done
} 2>&1 | grep "copy_kernel\|decompress:" | paste - - | awk '{print $1, $5, $6, "GiB  leg", $10, "ms  copy", $(NF-3), "ms"}' | tee gpurun_out/r02v_timing.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02v_pytest.txt
cat gpurun_out/r02v_pytest.txt
