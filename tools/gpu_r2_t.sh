#!/bin/bash
# Round 2, call T: bank kernel with the head of the block sums fetched before the z hand-off.
set -x -o pipefail
mkdir -p gpurun_out
timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-200 || exit 1
timeout 90 python tools/probe.py --streams 18944 --frames 6 --reps 3 --lanes 96 2>&1 | tail -1 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "96 or shapes_at_scale or config3 or bank_cli" 2>&1 | tail -3
