#!/bin/bash
# r02a: gather-path D2 -- GPU tests, then A/B timing against the span-assembly path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02a_pytest.txt
cat gpurun_out/r02a_pytest.txt
{
FOURMC_D2_GATHER=0 timeout 600 python tools/quick_decode.py 16 3
FOURMC_D2_GATHER=1 timeout 600 python tools/quick_decode.py 16 3
FOURMC_D2_GATHER=1 FOURMC_D2_WARPS=2 timeout 600 python tools/quick_decode.py 16 3
FOURMC_D2_GATHER=1 timeout 600 python tools/quick_decode.py 1 3
FOURMC_D2_GATHER=0 timeout 600 python tools/quick_decode.py 1 3
FOURMC_D2_GATHER=1 timeout 600 python tools/quick_decode.py 4 2 4mc 2
FOURMC_D2_GATHER=0 timeout 600 python tools/quick_decode.py 4 2 4mc 2
FOURMC_D2_GATHER=1 timeout 600 python tools/quick_decode.py 4 2 4mc 1
} 2>&1 | grep -v "^$" | tee gpurun_out/r02a_timing.txt
