#!/bin/bash
# r02l (1 GPU): full GPU test suite; E3 block size sweep (FOURMC_WRITE_THREADS); config 4 line with both CPU figures
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02l_pytest.txt
cat gpurun_out/r02l_pytest.txt
for t in 256 128 64; do
FOURMC_WRITE_THREADS=$t timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02l_wt.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['detail']['kernel_ms_rank0']; s=j['steps']
print('write_threads $t value', round(j['value'],1), {n:round(v['total_ms']/s,2) for n,v in k.items() if v['total_ms']/s>0.5})"
done
timeout 1500 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/r02l_bench_c4.json 2> gpurun_out/r02l_bench_c4.err; tail -c 600 gpurun_out/r02l_bench_c4.err; cut -c1-700 gpurun_out/r02l_bench_c4.json
