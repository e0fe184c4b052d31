#!/bin/bash
# Round 2: compute-sanitizer synccheck / racecheck over small banks of both demodulator kernels and the chain.
set -x
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for L in 96 32; do
  timeout 200 compute-sanitizer --tool synccheck --print-limit 8 python tools/probe.py --streams 48 --frames 1 --reps 1 --lanes $L > gpurun_out/synccheck_r02_$L.log 2>&1
  echo "synccheck lanes=$L: $(grep -E 'ERROR SUMMARY' gpurun_out/synccheck_r02_$L.log)"
done
timeout 90 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "96 or shapes_at_scale or bank_cli" 2>&1 | tail -3
