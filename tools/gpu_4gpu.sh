#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/r01e_bench_4gpu.json 2> gpurun_out/r01e_bench_4gpu.err
python - <<'PY'
import json
lines=[l for l in open("gpurun_out/r01e_bench_4gpu.json") if l.strip()]
print("stdout lines:", len(lines))
j=json.loads(lines[-1]); print("value %.1f n_gpus %d e2e %.1f" % (j["value"], j["n_gpus"], j["e2e"]["value"]), j["detail"]["step_ms"], j["e2e"])
PY
tail -3 gpurun_out/r01e_bench_4gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 --cpu-gib 1 2>/dev/null | cut -c1-200
