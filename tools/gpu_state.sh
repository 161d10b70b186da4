#!/bin/bash
# state check: parity tests, quick timings (LZ4 + 4mz), default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
FOURMC_PROFILE=1 timeout 300 python tools/quick_bench.py 4 2 > gpurun_out/prof_4g.txt 2>&1
timeout 600 python tools/quick_4mz.py 256 2 > gpurun_out/quick_4mz.txt 2>&1
timeout 900 python bench.py > gpurun_out/state_bench.json 2> gpurun_out/state_bench.err
cat gpurun_out/pytest_gpu.txt; grep profile gpurun_out/prof_4g.txt | tail -12; grep -v profile gpurun_out/prof_4g.txt | tail; cat gpurun_out/quick_4mz.txt gpurun_out/state_bench.json; tail -5 gpurun_out/state_bench.err
