#!/bin/bash
# full parity suite + smoke + short benches of both codecs
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --codec 4mz --steps 2 --warmup 3 > gpurun_out/chk_bench_4mz.json 2> gpurun_out/chk_bench_4mz.err; tail -c 1500 gpurun_out/chk_bench_4mz.json; tail -3 gpurun_out/chk_bench_4mz.err
