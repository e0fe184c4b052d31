#!/bin/bash
# warp-per-stream kernel: parity, lane sweep at three bank sizes, one full ncu capture
set -x
mkdir -p gpurun_out
TAG=${1:-r01b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_bank" 2>&1 | tail -8
for L in 32 4; do
  timeout 300 python tools/probe.py --streams 1024 --frames 8 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-330
done
for L in 32 4 2; do
  timeout 300 python tools/probe.py --streams 4096 --frames 4 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-330
done
for L in 4 2 1; do
  timeout 300 python tools/probe.py --streams 16384 --frames 3 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-330
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_warp_kernel -c 1 -f -o gpurun_out/prof_warp_$TAG \
    python tools/probe.py --streams 1024 --frames 4 --reps 1 --lanes 32 > gpurun_out/ncu_warp_$TAG.log 2>&1
tail -2 gpurun_out/ncu_warp_$TAG.log
ls -la gpurun_out
