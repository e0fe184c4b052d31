#!/bin/bash
# Round 2, call K: what slows the small-ring / low-register variants (ring vs register cap, at equal occupancy), and
# aggregate throughput of banks sized for 5, 6, 7 resident CTAs per SM.
set -x -o pipefail
mkdir -p gpurun_out
for C in 61 62; do
  OPVD_BANK_CTAS=$C timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-200 || exit 1
done
for C in 4 61 62; do
  OPVD_BANK_CTAS=$C timeout 60 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
done
OPVD_BANK_CTAS=5 timeout 60 python tools/probe.py --streams 23680 --frames 5 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
OPVD_BANK_CTAS=4 timeout 60 python tools/probe.py --streams 23680 --frames 5 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
OPVD_BANK_CTAS=6 timeout 60 python tools/probe.py --streams 28416 --frames 4 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
OPVD_BANK_CTAS=4 timeout 60 python tools/probe.py --streams 28416 --frames 4 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
OPVD_BANK_CTAS=7 timeout 60 python tools/probe.py --streams 33152 --frames 3 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
OPVD_BANK_CTAS=4 timeout 60 python tools/probe.py --streams 37888 --frames 3 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
