#!/bin/bash
# Round 2, call I: resident CTAs per SM of the bank kernel (ring rows / register cap variants 4..7).
set -x -o pipefail
mkdir -p gpurun_out
for C in 4 5 6 7; do
  OPVD_BANK_CTAS=$C timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-200 || exit 1
done
P="timeout 60 python tools/probe.py --streams 18944 --frames 6 --reps 2"
for C in 4 5 6 7; do
  OPVD_BANK_CTAS=$C $P --lanes 96 2>&1 | tail -1 | cut -c1-200
done
for C in 5 6 7; do
  OPVD_BANK_CTAS=$C timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_bank or ring or granularity or shapes_at_scale" 2>&1 | tail -3
done
