#!/bin/bash
# r02o (1 GPU): D2 with 128-byte items when warps share a block + undelivered sources demoted to slow matches; ncu capture of the W = 1 kernel
mkdir -p gpurun_out
{
for g in 0.015625 0.25 1 2; do timeout 600 python tools/quick_decode.py $g 2; done
FOURMC_D2_WARPS=4 timeout 600 python tools/quick_decode.py 0.25 2
FOURMC_D2_WARPS=2 timeout 600 python tools/quick_decode.py 1 2
timeout 300 python tools/latency_per_block.py
} 2>&1 | grep -v "^$" | grep -v "block_write\|block_size\|index_kernel\|scan_lens\|compress:\|stored_kernel\|finalize\|compact\|region_kernel" | tee gpurun_out/r02o_timing.txt
tools/ncu_cap.sh r02o_lz4_copy_kernel lz4_copy_kernel 1 python tools/quick_decode.py 16 1
grep -E "duration|inst_executed.sum |issue_active|warps_active|lsu_wavefronts.avg.pct|bank_conflicts_pipe_lsu_mem_shared.sum|dram__bytes|wavefronts_mem_shared.sum |op_ld.sum |op_st.sum |thread_inst_executed_per_inst" gpurun_out/r02o_lz4_copy_kernel_metrics.txt | grep -v pcsamp
head -60 gpurun_out/r02o_lz4_copy_kernel_hotlines.txt
