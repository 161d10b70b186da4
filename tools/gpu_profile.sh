#!/bin/bash
# round artefacts: bench lines (ours + reference), ncu launch list of the bench command, and a
# full capture of the kernels that dominate the step.  Summaries are copied to profiles/ by
# tools/summarize_profiles.py on the build box.
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# launch list of the same command (fewer steps: ncu serialises and replays)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
for k in lz4_region_kernel lz4_copy_kernel lz4_parse_kernel; do
  timeout 1500 ncu --set full --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_$k \
      python bench.py --steps 1 --warmup 1 --total-gib 16 --no-e2e --no-cpu > gpurun_out/${TAG}_$k.log 2>&1
done
# the 4mz path (configs[2] on the log-text input): bench lines, launch list, captures of its two hot kernels
python bench.py --codec 4mz > gpurun_out/${TAG}_bench_4mz.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --codec 4mz --impl reference > gpurun_out/${TAG}_bench_4mz_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_4mz.csv \
    python bench.py --codec 4mz --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_launches_4mz.log 2>&1
for k in zstd_entropy_kernel zstd_frames_lane_kernel; do
  SKIP=1; [ $k = zstd_entropy_kernel ] && SKIP=8        # the warm-up step's launches (8 groups of 512 blocks per 16 GiB)
  timeout 1500 ncu --set full --clock-control none -k regex:$k -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$k \
      python bench.py --codec 4mz --steps 1 --warmup 1 --total-gib 16 --no-e2e --no-cpu > gpurun_out/${TAG}_$k.log 2>&1
done
cat gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench_reference.json gpurun_out/${TAG}_bench_4mz.json gpurun_out/${TAG}_bench_4mz_reference.json; tail -3 gpurun_out/${TAG}_bench.err
