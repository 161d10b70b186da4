#!/bin/bash
# r02k (1 GPU): compute-sanitizer over the rebuilt chain kernels: memcheck (4.5 MiB), racecheck / synccheck / initcheck (tiny)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_chain.py > gpurun_out/r02k_memcheck.txt 2>&1; tail -4 gpurun_out/r02k_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_chain.py tiny > gpurun_out/r02k_racecheck.txt 2>&1; tail -6 gpurun_out/r02k_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_chain.py tiny > gpurun_out/r02k_synccheck.txt 2>&1; tail -3 gpurun_out/r02k_synccheck.txt
timeout 600 compute-sanitizer --tool initcheck python tools/sanitize_chain.py tiny > gpurun_out/r02k_initcheck.txt 2>&1; tail -3 gpurun_out/r02k_initcheck.txt
