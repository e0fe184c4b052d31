#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode or one_bank or resident" 2>&1 | tail -5
timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 64 2>&1 | tail -1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tasks_kernel -c 1 -f -o gpurun_out/prof_decode2_r01_d \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 64 > gpurun_out/ncu_decode2.log 2>&1
