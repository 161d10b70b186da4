#!/bin/bash
# two ranks on one box: both codecs through bench.py (weak scaling, NCCL all-gather of block lengths)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --total-gib 32 --batch-gib 16 > gpurun_out/r01c_bench_2gpu.json 2> gpurun_out/r01c_bench_2gpu.err
tail -c 700 gpurun_out/r01c_bench_2gpu.json; tail -2 gpurun_out/r01c_bench_2gpu.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --codec 4mz --steps 2 --warmup 3 --total-gib 32 --batch-gib 16 > gpurun_out/r01c_bench_4mz_2gpu.json 2> gpurun_out/r01c_bench_4mz_2gpu.err
tail -c 700 gpurun_out/r01c_bench_4mz_2gpu.json; tail -2 gpurun_out/r01c_bench_4mz_2gpu.err
