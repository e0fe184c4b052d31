#!/bin/bash
# Round 2, call D: three-warp bank kernel with the barrier-woken staging warp; full tests; bench.
set -x
mkdir -p gpurun_out
P="timeout 300 python tools/probe.py --streams 18944 --frames 6 --reps 2"
for QX in 1 2 0; do
  OPVD_BANK_QX=$QX $P --lanes 96 2>&1 | tail -1 | cut -c1-200
done
for S in 4096 8192 16384 37888; do
  timeout 300 python tools/probe.py --streams $S --frames 6 --reps 2 --lanes 96 2>&1 | tail -1 | cut -c1-200
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err
tail -c 3000 gpurun_out/bench_r02_d.json; tail -5 gpurun_out/bench_r02_d.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_bank -c 1 -f -o gpurun_out/prof_bank_r02_d \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 96 > gpurun_out/ncu_bank_r02_d.log 2>&1
tail -2 gpurun_out/ncu_bank_r02_d.log
