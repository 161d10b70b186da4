#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "zstd or 4mz" 2>&1 | tail -3
echo "== tile 2048"; timeout 600 python tools/quick_4mz_enc.py 4 3 1 2>&1 | tail -2
echo "== tile 1024"; FOURMC_LIB=$PWD/gpurun_tmp_t1024.so timeout 600 python tools/quick_4mz_enc.py 4 3 1 2>&1 | tail -2
echo "== 16 GiB auto"; timeout 600 python tools/quick_4mz_enc.py 16 3 1 2>&1 | tail -2
for p in 16 20 28; do echo "per_sm=$p"; FOURMC_ZL_PER_SM=$p timeout 600 python tools/quick_4mz_enc.py 16 2 1 2>&1 | tail -1; done
timeout 600 python tools/quick_4mz.py 256 2 2>&1 | tail -2
