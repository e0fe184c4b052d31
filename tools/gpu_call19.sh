#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bank_cli or (one_bank and 32) or cli_dropin or granularity or compaction or init_offset" 2>&1 | tail -4
for S in 1024 1776; do
timeout 300 python tools/probe.py --streams $S --frames 25 --reps 2 --lanes 32 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['S'], d['frames'], 'demod_ms', round(d['ms']['demod'],2), 'Gsps', round(d['S']*(d['frames']*86720+8000)/d['ms']['demod']/1e6,1))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_warp_kernel -c 1 -f -o gpurun_out/prof_warp_r01_e \
    python tools/probe.py --streams 1024 --frames 4 --reps 1 > gpurun_out/ncu_warp_r01_e.log 2>&1
