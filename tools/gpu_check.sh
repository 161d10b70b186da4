#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -3
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/quick_4mz_enc.py 4 3 1 2>&1 | tail -2
