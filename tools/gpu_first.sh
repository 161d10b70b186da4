#!/bin/bash
# first GPU contact: environment, parity tests, quick timings, ncu launch list
mkdir -p gpurun_out
{
  echo "== host"; nproc; lscpu | grep -E "Model name|Socket|Thread|Core|MHz" ; free -g | head -2
  echo "== gpu"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
} > gpurun_out/env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
timeout 600 python tools/quick_bench.py 1 3 > gpurun_out/quick_1g.txt 2>&1
timeout 600 python tools/quick_bench.py 4 3 > gpurun_out/quick_4g.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_1g.csv python tools/quick_bench.py 1 1 > gpurun_out/ncu_run.txt 2>&1
cat gpurun_out/env.txt gpurun_out/pytest_gpu.txt gpurun_out/quick_1g.txt gpurun_out/quick_4g.txt
