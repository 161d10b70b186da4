#!/bin/bash
FOURMC_D2_GATHER=1 tools/ncu_cap.sh r02b_copy_gather lz4_copy_kernel 1 python tools/quick_decode.py 4 1
FOURMC_D2_GATHER=0 tools/ncu_cap.sh r02b_copy_span lz4_copy_kernel 1 python tools/quick_decode.py 4 1
head -70 gpurun_out/r02b_copy_gather_metrics.txt
