#!/bin/bash
# Round 2, call V: half-warp-per-stream kernel (two streams per warp): parity, then time against the warp kernel.
set -x -o pipefail
mkdir -p gpurun_out
timeout 120 python tools/probe.py --streams 64 --frames 3 --reps 1 --lanes 16 2>&1 | tail -1 | cut -c1-300 || exit 1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_bank and 16" 2>&1 | tail -5 || exit 1
for L in 16 32; do
  timeout 90 python tools/probe.py --streams 1024 --frames 25 --reps 3 --lanes $L 2>&1 | tail -1 | cut -c1-260
done
for S in 700 1184; do for L in 16 32; do
  timeout 90 python tools/probe.py --streams $S --frames 12 --reps 2 --lanes $L 2>&1 | tail -1 | cut -c1-260
done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16 or config1_full or config3_true" 2>&1 | tail -5
