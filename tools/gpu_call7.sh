#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for S in 4096 18944; do
  F=3; [ $S -le 4096 ] && F=6
  timeout 300 python tools/probe.py --streams $S --frames $F --reps 2 --lanes 64 2>&1 | tail -1 | cut -c1-200
done
timeout 300 python tools/probe.py --streams 1024 --frames 8 --reps 2 2>&1 | tail -1 | cut -c1-200
