#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(one_bank and 64) or config_shapes or config3 or 16_byte" 2>&1 | tail -3
for S in 18944 8192; do
timeout 300 python tools/probe.py --streams $S --frames 6 --reps 2 --lanes 64 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['S'], d['frames'], 'demod_ms', round(d['ms']['demod'],2), 'Gsps', round(d['S']*(d['frames']*86720+8000)/d['ms']['demod']/1e6,1))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_batch_kernel -c 1 -f -o gpurun_out/prof_batch_r01_f \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 64 > gpurun_out/ncu_batch_f.log 2>&1
