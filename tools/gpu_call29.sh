#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(one_bank and 64)" 2>&1 | tail -2
for F in 6 14; do
timeout 300 python tools/probe.py --streams 18944 --frames $F --reps 2 --lanes 64 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['S'], d['frames'], 'rep', d['rep'], 'demod_ms', round(d['ms']['demod'],2), 'Gsps', round(d['S']*(d['frames']*86720+8000)/d['ms']['demod']/1e6,1))"
done
