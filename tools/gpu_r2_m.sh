#!/bin/bash
# Round 2, call M (rerun after the carveout fix): time tiling of the resident bank (tracker + Viterbi of tile t under the
# demodulator of tile t+1).
set -x -o pipefail
mkdir -p gpurun_out
for T in 1 2 4 7; do
  timeout 90 python tools/probe.py --streams 18944 --frames 14 --reps 2 --lanes 96 --tiles $T 2>&1 | tail -1 | cut -c1-230 || exit 1
done
OPVD_TRACE=1 timeout 90 python tools/probe.py --streams 18944 --frames 14 --reps 1 --lanes 96 --tiles 4 2>&1 | grep "opvd trace" | tail -4
