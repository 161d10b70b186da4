#!/bin/bash
# r02y (1 GPU): levels 2..4 rebuilt (chain links in HBM over a sliding 64 KiB history, search at every position with
# chain swap, cost-optimal parse per slice): chain tests first, ratios and timings per level, then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "levels_2_to_4 or reproducible" 2>&1 | tail -15 | tee gpurun_out/r02y_chain_tests.txt
timeout 900 python tools/quick_levels.py 2 2>&1 | grep -v "^$" | tee gpurun_out/r02y_levels.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02y_pytest.txt
