#!/bin/bash
# r02h (2 GPUs): strong-scaling bench lines at N=1 and N=2, both codecs at N=2
mkdir -p gpurun_out
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; tail -2 gpurun_out/r02h_bench_n1.err; cut -c1-900 gpurun_out/r02h_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; tail -5 gpurun_out/r02h_bench_n2.err; cat gpurun_out/r02h_bench_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 2 --steps 3 --warmup 2 --no-e2e > gpurun_out/r02h_bench_n2_c2.json 2> gpurun_out/r02h_bench_n2_c2.err; tail -5 gpurun_out/r02h_bench_n2_c2.err; cut -c1-1200 gpurun_out/r02h_bench_n2_c2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 4 --steps 2 --warmup 1 --total-gib 8 > gpurun_out/r02h_bench_n2_c4.json 2> gpurun_out/r02h_bench_n2_c4.err; tail -5 gpurun_out/r02h_bench_n2_c4.err; cut -c1-1500 gpurun_out/r02h_bench_n2_c4.json
