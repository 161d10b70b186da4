#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_tiny.py > gpurun_out/racecheck.txt 2>&1; tail -6 gpurun_out/racecheck.txt
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_tiny.py > gpurun_out/synccheck.txt 2>&1; tail -3 gpurun_out/synccheck.txt
timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_tiny.py > gpurun_out/initcheck.txt 2>&1; tail -3 gpurun_out/initcheck.txt
