#!/bin/bash
# Round 2, call S: warp kernel with the next symbol's conversions issued in the AFC shadow.
set -x -o pipefail
mkdir -p gpurun_out
timeout 90 python tools/probe.py --streams 1024 --frames 25 --reps 3 --lanes 32 2>&1 | tail -1 | cut -c1-260 || exit 1
timeout 90 python tools/probe.py --streams 592 --frames 25 --reps 2 --lanes 32 2>&1 | tail -1 | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "32 or granularity or ring or config1 or config0 or cli_dropin" 2>&1 | tail -3
