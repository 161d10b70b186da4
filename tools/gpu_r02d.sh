#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02d_pytest.txt
cat gpurun_out/r02d_pytest.txt
{
timeout 600 python tools/quick_decode.py 16 3
timeout 600 python tools/quick_decode.py 8 2
timeout 600 python tools/quick_decode.py 2 2
timeout 600 python tools/quick_decode.py 0.25 2
timeout 600 python tools/quick_decode.py 4 2 4mc 2
timeout 600 python tools/quick_decode.py 4 2 4mc 1
} 2>&1 | grep -v "^$" | grep -v "region_kernel\|block_write\|block_size\|index_kernel\|scan_lens" | tee gpurun_out/r02d_timing.txt
tools/ncu_cap.sh r02d_fused lz4_decode_fused 1 python tools/quick_decode.py 4 1
grep -E "duration|inst_executed.sum|issue_active|warps_active|lsu_wavefronts.avg.pct|stalled" gpurun_out/r02d_fused_metrics.txt | grep -v pcsamp
head -40 gpurun_out/r02d_fused_hotlines.txt
