#!/bin/bash
mkdir -p gpurun_out
{
FOURMC_DF_DBG=1 timeout 600 python tools/quick_decode.py 0.25 2
FOURMC_DF_DBG=1 timeout 600 python tools/quick_decode.py 4 2
FOURMC_DF_DBG=1 timeout 600 python tools/quick_decode.py 16 2
FOURMC_DF_DBG=512000 timeout 600 python tools/quick_decode.py 4 2
FOURMC_DF_DBG=5120000 timeout 600 python tools/quick_decode.py 4 2
} 2>&1 | grep -v "^$" | grep -v "region_kernel\|block_write\|block_size\|index_kernel\|scan_lens\|compress" | tee gpurun_out/r02e_timing.txt
FOURMC_DF_DBG=1 tools/ncu_cap.sh r02e_parse lz4_decode_fused 1 python tools/quick_decode.py 1 1
grep -E "duration|inst_executed.sum|issue_active|warps_active|lsu_wavefronts.avg.pct|stalled" gpurun_out/r02e_parse_metrics.txt | grep -v pcsamp
head -50 gpurun_out/r02e_parse_hotlines.txt
