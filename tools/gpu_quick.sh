#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for g in 1 4 16; do timeout 300 python tools/quick_bench.py $g 2 2>&1 | grep -E "compress|equal"; done
FOURMC_PROFILE=1 timeout 300 python tools/quick_bench.py 1 1 2>&1 | grep "profile" | tail -7
FOURMC_PROFILE=1 timeout 300 python tools/quick_bench.py 16 1 2>&1 | grep "profile" | tail -7
