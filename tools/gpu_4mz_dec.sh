#!/bin/bash
# 4mz reader bring-up: sanitizer, parity tests, device-resident decode timings (reference-made and GPU-made streams)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitize.txt 2>&1
tail -4 gpurun_out/sanitize.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/quick_4mz.py 256 2 2>&1 | tail -5
FOURMC_PROFILE=1 timeout 600 python tools/quick_4mz_enc.py 4 2 1 2>&1 | grep -E "profile.*zstd_frames|4mz" | tail -8
timeout 600 python tools/quick_4mz_enc.py 16 2 1 2>&1 | tail -3
