#!/bin/bash
# final round-2 artefacts (1 GPU): bench lines of the four configs, both arms; ncu launch list of the bench command;
# captures of the hot kernels reduced to text (tools/ncu_cap.sh); per-level ratios and rates; per-block latency; GPU suite.
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
tools/ncu_cap.sh ${TAG}_lz4_region_kernel lz4_region_kernel 1 python tools/quick_decode.py 16 1
tools/ncu_cap.sh ${TAG}_lz4_copy_kernel lz4_copy_kernel 1 python tools/quick_decode.py 16 1
tools/ncu_cap.sh ${TAG}_lz4_parse_kernel 'lz4_parse_kernel' 1 python tools/quick_decode.py 16 1
for c in 3 2 4; do
  extra=""; [ $c = 4 ] && extra="--total-gib 16 --steps 3 --warmup 1"
  timeout 1500 python bench.py --config $c $extra > gpurun_out/${TAG}_bench_c$c.json 2>> gpurun_out/${TAG}_bench.err
  timeout 900 python bench.py --config $c --impl reference > gpurun_out/${TAG}_bench_c${c}_reference.json 2>> gpurun_out/${TAG}_bench.err
  cut -c1-260 gpurun_out/${TAG}_bench_c$c.json
done
timeout 600 python tools/quick_levels.py 2 2>&1 | grep -v "^$\|Exception\|Traceback\|File\|TypeError" > gpurun_out/${TAG}_levels.txt
tools/ncu_cap.sh ${TAG}_lz4_region_chain_kernel lz4_region_kernel 1 python tools/quick_decode.py 0.5 1 4mc 0 3
tools/ncu_cap.sh ${TAG}_lz4_chain_kernel lz4_chain_kernel 1 python tools/quick_decode.py 0.5 1 4mc 0 3
timeout 300 python tools/latency_per_block.py > gpurun_out/${TAG}_latency.txt 2>&1
tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_latency.txt gpurun_out/${TAG}_levels.txt
