#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(one_bank and 32) or cli_dropin or granularity or compaction or config0 or config3 or 16_byte or resident" 2>&1 | tail -2
for S in 1024 1776; do
timeout 300 python tools/probe.py --streams $S --frames 25 --reps 2 --lanes 32 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['S'], d['frames'], 'demod_ms', round(d['ms']['demod'],2), 'Gsps', round(d['S']*(d['frames']*86720+8000)/d['ms']['demod']/1e6,1))"
done
