#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01f}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_bank and 64" 2>&1 | tail -15
for S in 1024 4096 16384; do
  F=3; [ $S -le 4096 ] && F=6
  timeout 300 python tools/probe.py --streams $S --frames $F --reps 2 --lanes 64 2>&1 | tail -1 | cut -c1-230
done
timeout 300 python tools/probe.py --streams 32768 --frames 2 --reps 1 --lanes 64 2>&1 | tail -1 | cut -c1-230
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_batch_kernel -c 1 -f -o gpurun_out/prof_batch_$TAG \
    python tools/probe.py --streams 16384 --frames 2 --reps 1 --lanes 64 > gpurun_out/ncu_batch_$TAG.log 2>&1
tail -1 gpurun_out/ncu_batch_$TAG.log | cut -c1-200
