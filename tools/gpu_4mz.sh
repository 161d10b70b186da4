#!/bin/bash
# 4mz writer bring-up: sanitizer on a small workload, parity tests, device-resident timings with the per-kernel profile
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitize.txt 2>&1
tail -4 gpurun_out/sanitize.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
FOURMC_PROFILE=1 timeout 300 python tools/quick_4mz_enc.py 1 2 0 2>&1 | grep -E "profile|4mz" | tail -12
timeout 300 python tools/quick_4mz_enc.py 1 2 1 2>&1 | tail -3
timeout 300 python tools/quick_4mz_enc.py 8 3 0 2>&1 | tail -2
