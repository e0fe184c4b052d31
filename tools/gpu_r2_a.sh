#!/bin/bash
# Round 2, call A: first run of the channel-bank kernel (lanes 96): parity, throughput vs the batched kernel, ncu.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "96" 2>&1 | tail -15
for QX in 1 0 2; do
  OPVD_BANK_QX=$QX timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-330
done
timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2 --lanes 64 2>&1 | tail -1 | cut -c1-330
for S in 4096 8192 16384 37888; do
  timeout 300 python tools/probe.py --streams $S --frames 6 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-330
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_bank -c 1 -f -o gpurun_out/prof_bank_r02_a \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 96 > gpurun_out/ncu_bank_r02_a.log 2>&1
tail -2 gpurun_out/ncu_bank_r02_a.log
ls -la gpurun_out | tail -5
