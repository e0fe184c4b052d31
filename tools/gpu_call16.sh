#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 64 2>&1 | tail -1 | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_bank and 64 or granularity or compaction or resident" 2>&1 | tail -3
