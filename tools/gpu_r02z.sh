#!/bin/bash
# r02z (1 GPU): where the time of the new level-3 compressor goes: per-kernel event times, one ncu capture of the chain parse
mkdir -p gpurun_out
{
timeout 600 python tools/quick_decode.py 1 1 4mc 0 3
timeout 600 python tools/quick_decode.py 1 1 4mc 0 2
} 2>&1 | grep -v "^$" | tee gpurun_out/r02z_timing.txt
tools/ncu_cap.sh r02z_lz4_region_chain_kernel lz4_region_kernel 1 python tools/quick_decode.py 0.5 1 4mc 0 3
head -60 gpurun_out/r02z_lz4_region_chain_kernel_hotlines.txt
grep -E "duration|issue_active.avg|warps_active|hit_rate|lts__throughput|l1tex__throughput|inst_executed.sum |thread_inst_executed_per_inst|stalled.*(long|short|wait|barrier|mio|lg_throttle|math|branch|not_selected|dispatch|no_inst|imc|sleep|drain|membar|tex)" gpurun_out/r02z_lz4_region_chain_kernel_metrics.txt | cut -c1-160
