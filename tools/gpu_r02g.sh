#!/bin/bash
# r02g: tests; bench lines of the four configs (1 GPU); file -> file timing with parallel positional I/O
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02g_pytest.txt
cat gpurun_out/r02g_pytest.txt
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err; cat gpurun_out/r02g_bench.json
timeout 900 python bench.py --config 2 --steps 4 --warmup 3 > gpurun_out/r02g_bench_c2.json 2> gpurun_out/r02g_bench_c2.err; tail -3 gpurun_out/r02g_bench_c2.err; cat gpurun_out/r02g_bench_c2.json
timeout 900 python bench.py --config 3 --steps 4 --warmup 2 > gpurun_out/r02g_bench_c3.json 2> gpurun_out/r02g_bench_c3.err; tail -3 gpurun_out/r02g_bench_c3.err; cat gpurun_out/r02g_bench_c3.json
timeout 1200 python bench.py --config 4 --steps 3 --warmup 1 --total-gib 8 > gpurun_out/r02g_bench_c4.json 2> gpurun_out/r02g_bench_c4.err; tail -3 gpurun_out/r02g_bench_c4.err; cat gpurun_out/r02g_bench_c4.json
FOURMC_CLI_TIMING=1 timeout 900 python tools/cli_file_timing.py 4 2>&1 | tee gpurun_out/r02g_cli_t2.txt
