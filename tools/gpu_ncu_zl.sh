#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/quick_4mz_enc.py 16 2 1 2>&1 | tail -1
timeout 600 python tools/quick_4mz_enc.py 4 2 1 2>&1 | tail -1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:zstd_frames_lane -c 1 -f -o gpurun_out/ncu_zl python tools/quick_4mz_enc.py 4 1 1 > gpurun_out/ncu_zl.log 2>&1
tail -2 gpurun_out/ncu_zl.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:zstd_entropy -c 1 -f -o gpurun_out/ncu_ze python tools/quick_4mz_enc.py 1 1 0 > gpurun_out/ncu_ze.log 2>&1
tail -2 gpurun_out/ncu_ze.log
