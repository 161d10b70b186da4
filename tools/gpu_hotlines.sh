#!/bin/bash
# source-level evidence: per hot kernel one `ncu --set full --import-source on` capture (4 GiB batch), reduced on
# the box to the top source lines by stall samples (tools/ncu_lines.py) -- only the text travels back
TAG=${1:-r01f}
mkdir -p gpurun_out /tmp/ncu
cap() {  # name regex command...
  local name=$1 rx=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s ${SKIP:-0} -c 1 -f -o /tmp/ncu/$name "$@" > /tmp/ncu/$name.log 2>&1
  { echo "# $name: top source lines by warp-stall samples (ncu --set full --import-source on; command: $*)"; python tools/ncu_lines.py /tmp/ncu/$name.ncu-rep 30; } > gpurun_out/${TAG}_${name}_hotlines.txt 2>&1
  rm -f /tmp/ncu/$name.ncu-rep
}
cap lz4_region_kernel 'lz4_region_kernel' python tools/quick_bench.py 4 1
cap lz4_parse_kernel 'lz4_parse_kernel' python tools/quick_bench.py 4 1
cap lz4_copy_kernel 'lz4_copy_kernel' python tools/quick_bench.py 4 1
FOURMC_CHAIN_DEPTH=16 cap lz4_region_chain_kernel 'lz4_region_kernel' python tools/quick_bench.py 1 1
cap zstd_entropy_kernel 'zstd_entropy' python tools/quick_4mz_enc.py 2 1 0
cap zstd_frames_lane_kernel 'zstd_frames_lane' python tools/quick_4mz_enc.py 4 1 1
ls -la gpurun_out/${TAG}_*hotlines.txt
