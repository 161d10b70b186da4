#!/bin/bash
# r02i: multi-warp parse (lz4_parse_wide_kernel) + wider D2 for few blocks: tests, per-block latency, small-batch decode timing
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02i_pytest.txt
cat gpurun_out/r02i_pytest.txt
{
timeout 300 python tools/latency_per_block.py
FOURMC_D1_WIDE=0 FOURMC_D2_WARPS=8 timeout 300 python tools/latency_per_block.py
for g in 0.015625 0.25 1 2 4; do
  timeout 600 python tools/quick_decode.py $g 2
  FOURMC_D1_WIDE=0 timeout 600 python tools/quick_decode.py $g 2
done
FOURMC_D1_WIDE=1 timeout 600 python tools/quick_decode.py 8 2
timeout 600 python tools/quick_decode.py 8 2
} 2>&1 | grep -v "^$" | grep -v "region_kernel\|block_write\|block_size\|index_kernel\|scan_lens\|compress:\|stored_kernel\|finalize\|compact" | tee gpurun_out/r02i_timing.txt
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('bench value', j['value'], 'e2e', j['e2e']['value'], j['e2e']['compress_GBps'], j['e2e']['decompress_GBps'])" | tee -a gpurun_out/r02i_timing.txt
