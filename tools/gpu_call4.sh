#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_bank and 64" 2>&1 | tail -3
for S in 1536 2048 4096 16384 18944; do
  F=3; [ $S -le 4096 ] && F=6
  timeout 300 python tools/probe.py --streams $S --frames $F --reps 2 --lanes 64 2>&1 | tail -1 | cut -c1-200
done
timeout 300 python tools/probe.py --streams 1536 --frames 6 --reps 2 --lanes 32 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/probe.py --streams 18944 --frames 3 --reps 1 --lanes 2 2>&1 | tail -1 | cut -c1-200
