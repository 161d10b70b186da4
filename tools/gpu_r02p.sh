#!/bin/bash
# r02p (1 GPU): D2 literal word loop, match loads ahead of the literals, tokens one batch ahead; pre-wait sweep for shared blocks
mkdir -p gpurun_out
{
timeout 600 python tools/quick_decode.py 16 2
FOURMC_D2_MINB=32 timeout 600 python tools/quick_decode.py 16 2
for pw in 0 8 32; do
FOURMC_D2_PREWAIT=$pw timeout 600 python tools/quick_decode.py 0.015625 2
FOURMC_D2_PREWAIT=$pw timeout 600 python tools/quick_decode.py 0.25 2
done
FOURMC_D2_WARPS=1 timeout 600 python tools/quick_decode.py 0.25 2
FOURMC_D2_WARPS=1 timeout 600 python tools/quick_decode.py 0.015625 2
} 2>&1 | grep -v "^$" | grep -v "block_write\|block_size\|index_kernel\|scan_lens\|compress:\|stored_kernel\|finalize\|compact\|region_kernel\|xxh_verify\|parse" | tee gpurun_out/r02p_timing.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02p_pytest.txt
cat gpurun_out/r02p_pytest.txt
