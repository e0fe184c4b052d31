#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_batch_kernel -c 1 -f -o gpurun_out/prof_batch_r01 \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 64 > gpurun_out/ncu_batch.log 2>&1
tail -2 gpurun_out/ncu_batch.log
