#!/bin/bash
# Round 2: the warp kernel cut for 16 resident CTAs per SM (124 registers) against the 182-register build and the bank kernel.
set -x -o pipefail
mkdir -p gpurun_out
P="timeout 90 python tools/probe.py --frames 12 --reps 2"
OPVD_WARP_MINB=8  $P --streams 1024 --lanes 32 2>&1 | tail -1 | cut -c1-200 || exit 1
OPVD_WARP_MINB=16 $P --streams 1024 --lanes 32 2>&1 | tail -1 | cut -c1-200
for S in 1536 2048 2368; do
  OPVD_WARP_MINB=16 $P --streams $S --lanes 32 2>&1 | tail -1 | cut -c1-200
  $P --streams $S --lanes 96 2>&1 | tail -1 | cut -c1-200
done
for S in 3072 4096; do
  OPVD_WARP_MINB=16 $P --streams $S --lanes 32 2>&1 | tail -1 | cut -c1-200
  $P --streams $S --lanes 96 2>&1 | tail -1 | cut -c1-200
done
