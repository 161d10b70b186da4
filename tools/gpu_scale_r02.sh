#!/bin/bash
# strong scaling of configs[1] on one 8-GPU box: N = 8, 4, 2, 1 back to back (the driver does the same at round end)
TAG=${1:-r02}
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
      bench.py --gpus $n --steps 4 --warmup 3 > gpurun_out/${TAG}_scale_n$n.json 2> gpurun_out/${TAG}_scale_n$n.err
  tail -2 gpurun_out/${TAG}_scale_n$n.err | cut -c1-300
done
timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_scale_n1.json 2> gpurun_out/${TAG}_scale_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
      bench.py --gpus 8 --config 2 --steps 3 --warmup 2 --no-e2e > gpurun_out/${TAG}_scale_c2_n8.json 2> gpurun_out/${TAG}_scale_c2_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 8 --config 3 --steps 3 --warmup 2 --no-e2e > gpurun_out/${TAG}_scale_c3_n8.json 2> gpurun_out/${TAG}_scale_c3_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_scale_*.json')):
    try:
        j=json.load(open(f)); r=j['roofline']
        print(f, 'value', round(j['value'],1), 'ms', round(j['ms_per_step'],1), 'legs', {k:round(v['ms_per_step'],1) for k,v in r['legs'].items()}, 'e2e', j['e2e'] and round(j['e2e']['value'],1), 'weak', j.get('weak') and round(j['weak']['value'],1))
        print('   ', {k:round(v['total_ms']/j['steps'],2) for k,v in j['detail']['kernel_ms_rank0'].items() if v['total_ms']/j['steps'] > 0.5})
    except Exception as e: print(f, 'ERR', e)
PY
