#!/bin/bash
# r02z3 (1 GPU): chain kernel without match_any on the common path; lanes take their next position every 4 candidates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "levels_2_to_4" 2>&1 | tail -5
{
timeout 600 python tools/quick_decode.py 1 1 4mc 0 3
timeout 600 python tools/quick_decode.py 1 1 4mc 0 2
timeout 600 python tools/quick_decode.py 1 1 4mc 0 4
timeout 600 python tools/quick_decode.py 1 1 4mc 2 3
} 2>&1 | grep -v "^$" | grep "compress\|region_chain\|lz4_chain\|ratio" | tee gpurun_out/r02z3_timing.txt
