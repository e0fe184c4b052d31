#!/bin/bash
# Round 2, call U: true-length parity tests through both kernels; refreshed ncu capture of the warp kernel.
set -x -o pipefail
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1_full or config3_true" --durations=4 2>&1 | tail -8
timeout 240 ncu --set full --clock-control none --import-source on -k regex:demod_warp -c 1 -f -o gpurun_out/prof_warp_r02_u \
    python tools/probe.py --streams 1024 --frames 4 --reps 1 --lanes 32 > gpurun_out/ncu_warp_r02_u.log 2>&1
tail -2 gpurun_out/ncu_warp_r02_u.log
