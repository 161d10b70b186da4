#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "live_reference or cli_levels" 2>&1 | tail -3
FOURMC_PROFILE=1 timeout 600 python tools/latency_per_block.py 2>&1 | grep -E "profile" | awk '{a[$3]+=$4; n[$3]++} END {for (k in a) printf "%-28s avg %.3f ms over %d\n", k, a[k]/n[k], n[k]}' | sort
timeout 600 python tools/latency_per_block.py 2>&1 | tail -1
