#!/bin/bash
# r02r (1 GPU): ncu capture of D2 variant 4 (literal word loop) at 16 GiB
mkdir -p gpurun_out
FOURMC_D2_VAR=4 tools/ncu_cap.sh r02r_lz4_copy_kernel lz4_copy_kernel 1 python tools/quick_decode.py 16 1
grep -E "duration|inst_executed.sum |issue_active|warps_active|lsu_wavefronts.avg.pct|bank_conflicts_pipe_lsu_mem_shared.sum|dram__bytes|wavefronts_mem_shared.sum |thread_inst_executed_per_inst|long_scoreboard_per|short_scoreboard_per|wait_per|not_selected_per|branch_resolving_per|no_instruction_per|math_pipe_throttle_per|lts__t_sector_hit|l1tex__t_sector_hit" gpurun_out/r02r_lz4_copy_kernel_metrics.txt | grep -v pcsamp
head -70 gpurun_out/r02r_lz4_copy_kernel_hotlines.txt
