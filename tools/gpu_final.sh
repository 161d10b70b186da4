#!/bin/bash
bash tools/gpu_profile.sh r01d > gpurun_out/r01d_profile.log 2>&1
tail -c 600 gpurun_out/r01d_bench.json; echo; tail -c 900 gpurun_out/r01d_bench_4mz.json; echo
timeout 900 python tools/quick_4mz.py 1024 2 2>&1 | tail -3
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
