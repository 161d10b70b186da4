#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitize.txt 2>&1
tail -3 gpurun_out/sanitize.txt
timeout 1500 python -m pytest tests -m gpu -x -q -k "zstd or 4mz or jni or cli" 2>&1 | tail -8
timeout 600 python tools/quick_4mz.py 256 2 2>&1 | tail -3
timeout 600 python tools/quick_4mz_enc.py 4 2 1 2>&1 | tail -1
timeout 600 python tools/quick_4mz_enc.py 16 2 1 2>&1 | tail -1
for p in 8 16 24; do echo "per_sm=$p"; FOURMC_ZL_PER_SM=$p timeout 600 python tools/quick_4mz_enc.py 16 2 1 2>&1 | tail -1; done
