#!/bin/bash
# Round 2, call M: does the bank kernel's time scale with the capture length (6, 10, 14 frames at 18,944 streams), and what
# does tiling the resident bank in time cost / buy (tracker + Viterbi of tile t under the demodulator of tile t+1).
set -x -o pipefail
mkdir -p gpurun_out
for F in 6 10 14; do
  timeout 90 python tools/probe.py --streams 18944 --frames $F --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-230 || exit 1
done
for T in 2 4 7; do
  timeout 90 python tools/probe.py --streams 18944 --frames 14 --reps 2 --lanes 96 --tiles $T 2>&1 | tail -1 | cut -c1-230
done
