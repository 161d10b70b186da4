#!/bin/bash
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --total-gib 32 --batch-gib 16 --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('value', j['value'], 'ms/step', j['ms_per_step'], j['detail']['step_ms'])
    else: print('STDOUT NOISE:', l.rstrip())
"
