#!/bin/bash
# Round 2, call J: occupancy facts of the bank-kernel variants (ncu LaunchStats + Occupancy sections).
set -x -o pipefail
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for C in 4 5 6 7; do
  OPVD_BANK_CTAS=$C timeout 120 ncu --section LaunchStats --section Occupancy --clock-control none -k regex:demod_bank -c 1 \
      python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 96 2>&1 | grep -E "Registers Per|Shared Memory|Block Limit|Occupancy|Active Warps|Threads|Waves" | sed "s/^/ctas=$C /"
done
