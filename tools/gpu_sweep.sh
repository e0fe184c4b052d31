#!/bin/bash
set -x
mkdir -p gpurun_out
./tools/microbench > gpurun_out/microbench2.json; cat gpurun_out/microbench2.json
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^E  .*array\|where" | tail -15
for L in 1 2 4; do
  timeout 300 python tools/probe.py --streams 1024 --frames 8 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-330
done
for L in 1 2 4; do
  timeout 300 python tools/probe.py --streams 16384 --frames 3 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-330
done
for L in 1 2; do
  timeout 300 python tools/probe.py --streams 65536 --frames 1 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-330
done
