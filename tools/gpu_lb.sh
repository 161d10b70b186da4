#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_chain.py 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -3
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/quick_levels.py 4 2>&1 | grep level
bash tools/gpu_e2e2.sh
