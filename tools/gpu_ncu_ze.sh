#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:zstd_entropy -c 1 -f -o gpurun_out/ncu_ze2 python tools/quick_4mz_enc.py 1 1 0 > gpurun_out/ncu_ze2.log 2>&1
tail -1 gpurun_out/ncu_ze2.log
