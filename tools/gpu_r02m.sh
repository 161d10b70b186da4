#!/bin/bash
# r02m (1 GPU): E1 one-word filter, D1 packed exit table, pageable uploads through a pinned bounce buffer
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02m_pytest.txt
cat gpurun_out/r02m_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02m_b.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['detail']['kernel_ms_rank0']; s=j['steps']
print('config1 value', round(j['value'],1), 'ms', round(j['ms_per_step'],1), {n:round(v['total_ms']/s,2) for n,v in k.items() if v['total_ms']/s>0.5})"
for t in 4 0 8; do
FOURMC_COPY_THREADS=$t timeout 1200 python bench.py --config 4 --steps 3 --warmup 1 --total-gib 16 --no-cpu 2>gpurun_out/r02m_c4.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('copy threads $t:', round(j['value'],2), 'GB/s', round(j['detail']['splits_per_s']), 'splits/s')"; tail -2 gpurun_out/r02m_c4.err | cut -c1-300
done
FOURMC_COPY_THREADS=4 timeout 1200 python bench.py --config 4 --steps 3 --warmup 1 --total-gib 16 --no-cpu --split-threads 4 2>gpurun_out/r02m_c4.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('copy threads 4, 4 readers:', round(j['value'],2), 'GB/s', round(j['detail']['splits_per_s']), 'splits/s')"
{ timeout 600 python tools/quick_decode.py 1 2; timeout 600 python tools/quick_decode.py 0.25 2; } 2>&1 | grep "decompress:\|parse\|copy_kernel"
