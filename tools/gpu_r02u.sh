#!/bin/bash
# r02u (1 GPU): where the checksum pass is launched (before the parse kernel / after it / after the copy kernel)
mkdir -p gpurun_out
{
for o in 0 1 2; do
FOURMC_VERIFY_ORDER=$o timeout 600 python tools/quick_decode.py 16 2
FOURMC_VERIFY_ORDER=$o timeout 600 python tools/quick_decode.py 1 2
done
} 2>&1 | grep "decompress:\|copy_kernel\|parse\|verify" | tee gpurun_out/r02u_timing.txt
