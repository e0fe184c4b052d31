#!/bin/bash
# Round 2: ncu --set full capture of the final bank kernel (default build) with source correlation.
set -x -o pipefail
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 200 ncu --set full --clock-control none --import-source on -k regex:demod_bank -c 1 -f -o gpurun_out/prof_bank_r02_final \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 96 > gpurun_out/ncu_bank_r02_final.log 2>&1
tail -2 gpurun_out/ncu_bank_r02_final.log
