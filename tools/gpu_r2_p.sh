#!/bin/bash
# Round 2, call P: biased-sample conversion (QX=3: byte permutes instead of I2F, bias folded out of the block sums).
set -x -o pipefail
mkdir -p gpurun_out
for Q in 3 1; do
  OPVD_BANK_QX=$Q timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-400 || exit 1
done
for Q in 3 1 3 1; do
  OPVD_BANK_QX=$Q timeout 60 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
done
