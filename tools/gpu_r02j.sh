#!/bin/bash
# r02j: after the D2 register fix: bench line, per-block latency, split reads with the cached index, e2e slice sizes
mkdir -p gpurun_out
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -2 gpurun_out/r02j_bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02j_bench.json'))
print('value', j['value'], 'ms', j['ms_per_step'], 'legs', {k:round(v['ms_per_step'],1) for k,v in j['roofline']['legs'].items()}, 'e2e', j['e2e']['value'], j['e2e']['compress_GBps'], j['e2e']['decompress_GBps'])
print({k:round(v['total_ms']/4,2) for k,v in j['detail']['kernel_ms_rank0'].items()})
PY
timeout 300 python tools/latency_per_block.py | tee gpurun_out/r02j_latency.txt
for sb in 64 256; do FOURMC_SLICE_BLOCKS=$sb timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('slice $sb: e2e', j['e2e']['value'], j['e2e']['compress_GBps'], j['e2e']['decompress_GBps'])"; done
timeout 1200 python bench.py --config 4 --steps 3 --warmup 1 --total-gib 8 --no-cpu > gpurun_out/r02j_bench_c4.json 2> gpurun_out/r02j_bench_c4.err; tail -3 gpurun_out/r02j_bench_c4.err; cut -c1-400 gpurun_out/r02j_bench_c4.json; python -c "
import json; j=json.load(open('gpurun_out/r02j_bench_c4.json')); print(j['value'], j['detail']['splits_per_s'])"
timeout 1200 python bench.py --config 4 --steps 3 --warmup 1 --total-gib 8 --no-cpu --split-threads 16 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('16 threads', j['value'], j['detail']['splits_per_s'])"
