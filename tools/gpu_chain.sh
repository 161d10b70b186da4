#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_chain.py 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q -k "levels" 2>&1 | tail -3
timeout 600 python tools/quick_levels.py 4 2>&1 | grep level
FOURMC_PROFILE=1 FOURMC_CHAIN_DEPTH=4 timeout 300 python tools/quick_bench.py 4 1 2>&1 | grep -E "profile.*(region|write)" | tail -3
