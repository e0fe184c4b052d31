#!/bin/bash
# Round 2: automatic kernel choice across the crossover (warp kernel 188 / 124 registers, bank kernel with / without ELB).
set -x -o pipefail
mkdir -p gpurun_out
P="timeout 90 python tools/probe.py --frames 12 --reps 2"
for S in 1024 1184 1500 1776 1800 2368 4736; do
  $P --streams $S 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('S', d['S'], 'auto demod', round(d['ms']['demod'], 3))" || exit 1
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
