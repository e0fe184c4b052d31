#!/bin/bash
# Round 2, call C: half-window bank kernel (lanes 128): probes, parity tests, ncu.
set -x
mkdir -p gpurun_out
P="timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2"
for QX in 2 1 0; do
  OPVD_BANK_QX=$QX $P --lanes 128 2>&1 | tail -1 | cut -c1-200
done
OPVD_BANK_QX=1 $P --lanes 96 2>&1 | tail -1 | cut -c1-200
for S in 4096 8192 16384 37888; do
  timeout 300 python tools/probe.py --streams $S --frames 6 --reps 2 --lanes 128 2>&1 | tail -1 | cut -c1-200
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_bank16 -c 1 -f -o gpurun_out/prof_bank16_r02_c \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 128 > gpurun_out/ncu_bank16_r02_c.log 2>&1
tail -2 gpurun_out/ncu_bank16_r02_c.log
