#!/bin/bash
# kernel-variant sweep over bank sizes (tunes demod_auto_lanes)
mkdir -p gpurun_out
for S in 1776 2368 4096 8192 12288; do
  for L in 32 64; do
    F=$(( 120000 / S + 2 ))
    timeout 300 python tools/probe.py --streams $S --frames $F --reps 2 --lanes $L 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['S'], $L, d['frames'], 'demod_ms', round(d['ms']['demod'],2), 'Gsps', round(d['S']*(d['frames']*86720+8000)/d['ms']['demod']/1e6,1))"
  done
done
