#!/bin/bash
# pipe kernel v2 (register-staged ring, 256-thread CTA): parity + probe vs batch
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_bank and 128" 2>&1 | tail -5
for L in 128 64; do
  timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes $L 2>&1 | tail -1 | cut -c1-330
done
timeout 300 python tools/probe.py --streams 37888 --frames 3 --reps 2 --lanes 128 2>&1 | tail -1 | cut -c1-330
timeout 300 python tools/probe.py --streams 16384 --frames 6 --reps 2 --lanes 128 2>&1 | tail -1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_pipe_kernel -c 1 -f -o gpurun_out/prof_pipe2_r01 \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 128 > gpurun_out/ncu_pipe2.log 2>&1
tail -2 gpurun_out/ncu_pipe2.log
