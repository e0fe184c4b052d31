#!/bin/bash
for S in 8192 4096; do for F in 6 28; do
timeout 300 python tools/probe.py --streams $S --frames $F --reps 2 --lanes 64 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['S'], d['frames'], 'rep', d['rep'], 'demod_ms', round(d['ms']['demod'],2), 'Gsps', round(d['S']*(d['frames']*86720+8000)/d['ms']['demod']/1e6,1))"
done; done
