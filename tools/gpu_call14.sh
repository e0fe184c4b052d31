#!/bin/bash
# 2-GPU check of the sharded bench (NCCL counter reduction, max-over-ranks timing) and of the reference arm under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2_r01_f.json 2> gpurun_out/bench_n2_r01_f.err
tail -1 gpurun_out/bench_n2_r01_f.json | cut -c1-1500; tail -5 gpurun_out/bench_n2_r01_f.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2_r01_f.json 2> gpurun_out/bench_ref_n2_r01_f.err
tail -1 gpurun_out/bench_ref_n2_r01_f.json | cut -c1-300
