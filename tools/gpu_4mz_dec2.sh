#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "zstd or 4mz" 2>&1 | tail -5
timeout 600 python tools/quick_4mz.py 256 2 2>&1 | tail -3
timeout 600 python tools/quick_4mz_enc.py 4 2 1 2>&1 | tail -2
timeout 600 python tools/quick_4mz_enc.py 16 2 1 2>&1 | tail -2
timeout 900 python bench.py --codec 4mz --total-gib 16 --batch-gib 8 --e2e-gib 4 --cpu-gib 1 --steps 2 --warmup 3 2>&1 | tail -3
