#!/bin/bash
# Round 2, call O: ncu --set full captures of the two product demodulator kernels of the final tree (source-level).
set -x -o pipefail
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-200 || exit 1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:demod_bank -c 1 -f -o gpurun_out/prof_bank_r02_o \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 96 > gpurun_out/ncu_bank_r02_o.log 2>&1
tail -2 gpurun_out/ncu_bank_r02_o.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:demod_warp -c 1 -f -o gpurun_out/prof_warp_r02_o \
    python tools/probe.py --streams 1024 --frames 4 --reps 1 --lanes 32 > gpurun_out/ncu_warp_r02_o.log 2>&1
tail -2 gpurun_out/ncu_warp_r02_o.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:decode_tasks -c 1 -f -o gpurun_out/prof_decode_r02_o \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 96 > gpurun_out/ncu_decode_r02_o.log 2>&1
tail -2 gpurun_out/ncu_decode_r02_o.log
