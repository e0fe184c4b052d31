#!/bin/bash
# Re-entry check: GPU tests for every kernel variant, pipe-vs-batch probe at the channel-bank size, smoke.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for L in 64 128; do
  timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes $L 2>&1 | tail -1 | cut -c1-600
done
for L in 32 128; do
  timeout 300 python tools/probe.py --streams 1024 --frames 8 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-400
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_pipe_kernel -c 1 -f -o gpurun_out/prof_pipe_r01 \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 128 > gpurun_out/ncu_pipe.log 2>&1
tail -2 gpurun_out/ncu_pipe.log
ls -la gpurun_out
