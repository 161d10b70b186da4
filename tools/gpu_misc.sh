#!/bin/bash
mkdir -p gpurun_out
echo "== tile 1024"; timeout 600 python tools/quick_4mz_enc.py 4 3 0 2>&1 | tail -1
echo "== tile 512"; FOURMC_LIB=$PWD/gpurun_tmp_t512.so timeout 600 python tools/quick_4mz_enc.py 4 3 0 2>&1 | tail -1
FOURMC_CHAIN_DEPTH=4 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lz4_region_kernel -c 1 -f -o gpurun_out/ncu_chain python tools/quick_bench.py 1 1 > gpurun_out/ncu_chain.log 2>&1
tail -2 gpurun_out/ncu_chain.log
