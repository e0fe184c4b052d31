#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01d}
./tools/microbench > gpurun_out/microbench_$TAG.json; cat gpurun_out/microbench_$TAG.json
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_bank and 32" 2>&1 | tail -4
timeout 300 python tools/probe.py --streams 1024 --frames 8 --reps 2 --lanes 32 2>&1 | tail -1 | cut -c1-230
timeout 300 python tools/probe.py --streams 2048 --frames 8 --reps 1 --lanes 32 2>&1 | tail -1 | cut -c1-230
timeout 300 python tools/probe.py --streams 4096 --frames 4 --reps 1 --lanes 32 2>&1 | tail -1 | cut -c1-230
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_warp_kernel -c 1 -f -o gpurun_out/prof_warp_$TAG \
    python tools/probe.py --streams 1024 --frames 4 --reps 1 --lanes 32 > gpurun_out/ncu_warp_$TAG.log 2>&1
tail -1 gpurun_out/ncu_warp_$TAG.log | cut -c1-200
