#!/bin/bash
# Round 2, call B: four-warp bank kernel + rewritten runtime: probes, full GPU tests, ncu.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
nproc; free -g | head -2
P="timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2"
for QX in 2 1 0; do
  OPVD_BANK_QX=$QX $P --lanes 128 2>&1 | tail -1 | cut -c1-200
done
OPVD_BANK_ROTATE=0 $P --lanes 128 2>&1 | tail -1 | cut -c1-200
OPVD_BANK_QX=2 $P --lanes 96 2>&1 | tail -1 | cut -c1-200
OPVD_BANK_QX=1 $P --lanes 96 2>&1 | tail -1 | cut -c1-200
for S in 4096 8192 37888; do
  timeout 300 python tools/probe.py --streams $S --frames 6 --reps 2 --lanes 128 2>&1 | tail -1 | cut -c1-200
done
timeout 300 python tools/probe.py --streams 1024 --frames 25 --reps 2 --lanes 32 2>&1 | tail -1 | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_bank4 -c 1 -f -o gpurun_out/prof_bank4_r02_b \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 128 > gpurun_out/ncu_bank4_r02_b.log 2>&1
tail -2 gpurun_out/ncu_bank4_r02_b.log
