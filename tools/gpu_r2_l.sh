#!/bin/bash
# Round 2, call L (two GPUs): the C++ bank host over two devices (NCCL counter reduce), both bench arms at N=2.
set -x -o pipefail
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bank_cli" 2>&1 | tail -3 || exit 1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $RUN bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2_r02_l3.json 2> gpurun_out/bench_ref_n2_r02_l3.err
tail -1 gpurun_out/bench_ref_n2_r02_l3.json | cut -c1-400
timeout 600 $RUN bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_r02_l3.json 2> gpurun_out/bench_n2_r02_l3.err
tail -c 1800 gpurun_out/bench_n2_r02_l3.json; tail -3 gpurun_out/bench_n2_r02_l3.err
