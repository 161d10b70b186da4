#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -2 gpurun_out/quick_bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/quick_bench.json'))
print('value', j['value'], 'ms', j['ms_per_step'], 'legs', {k:round(v['ms_per_step'],1) for k,v in j['roofline']['legs'].items()}, 'e2e', j['e2e']['value'], j['e2e']['compress_GBps'], j['e2e']['decompress_GBps'])
print({k:round(v['total_ms']/4,2) for k,v in j['detail']['kernel_ms_rank0'].items()})
PY
timeout 300 python tools/latency_per_block.py
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_splits_gpu.py -m gpu -x -q 2>&1 | tail -3
