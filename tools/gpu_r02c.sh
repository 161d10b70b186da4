#!/bin/bash
# r02c: fused LZ4 decode (lane-per-block parse + checksum + copy warps in one kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02c_pytest.txt
cat gpurun_out/r02c_pytest.txt
{
timeout 600 python tools/quick_decode.py 16 3
FOURMC_DEC_MODE=split timeout 600 python tools/quick_decode.py 16 2
timeout 600 python tools/quick_decode.py 8 2
timeout 600 python tools/quick_decode.py 2 2
timeout 600 python tools/quick_decode.py 0.25 2
FOURMC_DEC_MODE=split timeout 600 python tools/quick_decode.py 0.25 2
timeout 600 python tools/quick_decode.py 4 2 4mc 2
timeout 600 python tools/quick_decode.py 4 2 4mc 1
} 2>&1 | grep -v "^$" | grep -v "region_kernel\|block_write\|block_size\|index_kernel\|scan_lens" | tee gpurun_out/r02c_timing.txt
