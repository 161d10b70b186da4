#!/bin/bash
# r02f: split decode with one-ahead prefetch in D2, short-block / empty-stream handling, streaming file API
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02f_pytest.txt
cat gpurun_out/r02f_pytest.txt
{
timeout 600 python tools/quick_decode.py 16 3
timeout 600 python tools/quick_decode.py 2 2
timeout 600 python tools/quick_decode.py 0.25 2
} 2>&1 | grep -v "^$" | grep -v "region_kernel\|block_write\|block_size\|index_kernel\|scan_lens" | tee gpurun_out/r02f_timing.txt
FOURMC_CLI_TIMING=1 timeout 900 python tools/cli_file_timing.py 4 2>&1 | tee gpurun_out/r02f_cli_t2.txt
FOURMC_CLI_TIMING=1 timeout 900 python tools/cli_file_timing.py 4 -z 2>&1 | tee gpurun_out/r02f_cli_t2_4mz.txt
