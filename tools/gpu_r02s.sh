#!/bin/bash
# r02s (1 GPU): D2 as a two-stage pipeline (match sources by cp.async one batch ahead, tokens two ahead)
mkdir -p gpurun_out
{
for g in 16 1 0.25 0.015625; do timeout 600 python tools/quick_decode.py $g 2; done
timeout 600 python tools/quick_decode.py 4 2 4mc 2
timeout 300 python tools/latency_per_block.py
} 2>&1 | grep -v "^$" | grep -v "block_write\|block_size\|index_kernel\|scan_lens\|compress:\|stored_kernel\|finalize\|compact\|region_kernel" | tee gpurun_out/r02s_timing.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02s_pytest.txt
cat gpurun_out/r02s_pytest.txt
