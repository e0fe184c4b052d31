#!/bin/bash
set -x
mkdir -p gpurun_out
for L in 4 1; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_kernel -c 1 -f -o gpurun_out/prof_demod_L$L \
    python tools/probe.py --streams 1024 --frames 2 --reps 1 --lanes $L > gpurun_out/ncu_L$L.log 2>&1
tail -1 gpurun_out/ncu_L$L.log | cut -c1-200
done
