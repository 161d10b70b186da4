#!/bin/bash
# r02n (1 GPU): D2 rebuilt around 1 KiB work items and word-granular span assembly: quick round trips first, then the suite
mkdir -p gpurun_out
{
for g in 0.25 1 16; do timeout 600 python tools/quick_decode.py $g 2; done
FOURMC_D2_WARPS=8 timeout 600 python tools/quick_decode.py 0.015625 2
FOURMC_D2_WARPS=4 timeout 600 python tools/quick_decode.py 0.25 2
timeout 600 python tools/quick_decode.py 4 2 4mc 2
timeout 600 python tools/quick_decode.py 4 2 4mc 1
} 2>&1 | grep -v "^$" | grep -v "block_write\|block_size\|index_kernel\|scan_lens\|compress:\|stored_kernel\|finalize\|compact" | tee gpurun_out/r02n_timing.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02n_pytest.txt
cat gpurun_out/r02n_pytest.txt
