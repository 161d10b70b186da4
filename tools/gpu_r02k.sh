#!/bin/bash
# r02k (1 GPU): tests with the templated wide parse, EOL rules, batched splits; config 4 with batched splits; 2 GiB decode with wide<7>
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02k_pytest.txt
cat gpurun_out/r02k_pytest.txt
for a in "--splits-per-call 64 --split-threads 2" "--splits-per-call 64 --split-threads 4" "--splits-per-call 16 --split-threads 4" "--splits-per-call 1 --split-threads 8"; do
timeout 1200 python bench.py --config 4 --steps 3 --warmup 1 --total-gib 16 --no-cpu $a 2>gpurun_out/r02k_c4.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$a', round(j['value'],2), round(j['detail']['splits_per_s']))"; tail -2 gpurun_out/r02k_c4.err | cut -c1-300
done
{ timeout 600 python tools/quick_decode.py 2 2; timeout 600 python tools/quick_decode.py 1 2; FOURMC_D1_WIDE=0 timeout 600 python tools/quick_decode.py 2 2; } 2>&1 | grep "decompress:\|parse\|copy_kernel"
