#!/bin/bash
# Round 2, call N: estimate kernel with integer (IMAD.WIDE) accumulation: time and parity.
set -x -o pipefail
mkdir -p gpurun_out
timeout 90 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-230 || exit 1
timeout 90 python tools/probe.py --streams 1024 --frames 6 --reps 2 2>&1 | tail -1 | cut -c1-230
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_bank or est or cli_dropin or shapes_at_scale or abi" 2>&1 | tail -3
