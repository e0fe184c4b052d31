#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_bank" 2>&1 | tail -3
timeout 300 python tools/probe.py --streams 18944 --frames 3 --reps 2 --lanes 64 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/probe.py --streams 1024 --frames 8 --reps 2 2>&1 | tail -1 | cut -c1-200
