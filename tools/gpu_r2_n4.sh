#!/bin/bash
# Round 2: bench at N=4 (both arms) under torchrun.
set -x -o pipefail
mkdir -p gpurun_out
nvidia-smi -L | head -4
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $RUN bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_ref_n4_r02.json 2> gpurun_out/bench_ref_n4_r02.err
tail -1 gpurun_out/bench_ref_n4_r02.json | cut -c1-300
timeout 600 $RUN bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4_r02.json 2> gpurun_out/bench_n4_r02.err
tail -c 600 gpurun_out/bench_n4_r02.json; tail -3 gpurun_out/bench_n4_r02.err
