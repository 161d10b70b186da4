#!/bin/bash
# round-2 artefacts (1 GPU): bench lines of the four configs, both arms; ncu launch list of the bench command; captures
# of the hot kernels reduced to text (tools/ncu_cap.sh); per-block latency; file -> file timing.
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
for c in 2 3 4; do
  extra=""; [ $c = 4 ] && extra="--total-gib 16 --steps 3 --warmup 1"
  timeout 1500 python bench.py --config $c $extra > gpurun_out/${TAG}_bench_c$c.json 2>> gpurun_out/${TAG}_bench.err
  timeout 900 python bench.py --config $c --impl reference > gpurun_out/${TAG}_bench_c${c}_reference.json 2>> gpurun_out/${TAG}_bench.err
done
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
tools/ncu_cap.sh ${TAG}_lz4_copy_kernel lz4_copy_kernel 1 python tools/quick_decode.py 16 1
tools/ncu_cap.sh ${TAG}_lz4_region_kernel lz4_region_kernel 1 python tools/quick_decode.py 16 1
tools/ncu_cap.sh ${TAG}_lz4_parse_kernel 'lz4_parse_kernel' 1 python tools/quick_decode.py 16 1
tools/ncu_cap.sh ${TAG}_lz4_parse_wide_kernel lz4_parse_wide 1 python tools/quick_decode.py 0.5 1
tools/ncu_cap.sh ${TAG}_zstd_frames_lane_kernel zstd_frames_lane 1 python tools/quick_decode.py 4 1 4mz
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python tools/latency_per_block.py > gpurun_out/${TAG}_latency.txt 2>&1
FOURMC_CLI_TIMING=1 timeout 900 python tools/cli_file_timing.py 4 > gpurun_out/${TAG}_cli_t2.txt 2>&1
for f in gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench_reference.json gpurun_out/${TAG}_bench_c2.json gpurun_out/${TAG}_bench_c3.json gpurun_out/${TAG}_bench_c4.json; do cut -c1-260 $f; done
tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_latency.txt; tail -3 gpurun_out/${TAG}_cli_t2.txt | cut -c1-600
grep -E "duration|dram__bytes|lsu_wavefronts.avg.pct|issue_active.avg|bank_conflicts_pipe_lsu_mem_shared.sum" gpurun_out/${TAG}_lz4_copy_kernel_metrics.txt
