#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.json | cut -c1-2500; tail -5 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-600
