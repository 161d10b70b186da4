#!/bin/bash
# iteration run: parity tests, per-kernel profile, quick timings, optional ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
FOURMC_PROFILE=1 timeout 600 python tools/quick_bench.py 1 2 > gpurun_out/prof_1g.txt 2>&1
timeout 600 python tools/quick_bench.py 1 3 > gpurun_out/quick_1g.txt 2>&1
timeout 600 python tools/quick_bench.py 4 3 > gpurun_out/quick_4g.txt 2>&1
timeout 600 python tools/quick_bench.py 16 3 > gpurun_out/quick_16g.txt 2>&1
kill $SMI
if [ -n "$NCU_KERNELS" ]; then
  for k in $NCU_KERNELS; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/ncu_$k python tools/quick_bench.py 1 1 > gpurun_out/ncu_$k.log 2>&1
  done
fi
cat gpurun_out/pytest_gpu.txt; grep -v "^\[fourmc profile\] --" gpurun_out/prof_1g.txt | tail -22; cat gpurun_out/quick_1g.txt gpurun_out/quick_4g.txt gpurun_out/quick_16g.txt
sort gpurun_out/clocks.csv | uniq -c | sort -rn | head -8
