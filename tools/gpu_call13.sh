#!/bin/bash
# Round-1 "d" evidence: GPU tests, smoke, bench (both arms), ncu launch list, ncu full captures of the three demod kernels.
set -x
mkdir -p gpurun_out
TAG=r01_f
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
nproc; free -g | head -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -1 gpurun_out/bench_ref_$TAG.json | cut -c1-600
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -1 gpurun_out/bench_$TAG.json | cut -c1-3000; tail -3 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
grep -c . gpurun_out/launches_$TAG.csv


timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tasks_kernel -c 1 -f -o gpurun_out/prof_decode_$TAG \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 64 > gpurun_out/ncu_decode_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_batch_kernel -c 1 -f -o gpurun_out/prof_batch_$TAG \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 64 > gpurun_out/ncu_batch_$TAG.log 2>&1
ls -la gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_warp_kernel -c 1 -f -o gpurun_out/prof_warp_$TAG \
    python tools/probe.py --streams 1024 --frames 4 --reps 1 > gpurun_out/ncu_warp_$TAG.log 2>&1
for S in 2368 3072 4096; do
  for L in 32 64; do
    F=$(( 120000 / S + 2 ))
    timeout 300 python tools/probe.py --streams $S --frames $F --reps 2 --lanes $L 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sweep', d['S'], $L, d['frames'], 'demod_ms', round(d['ms']['demod'],2), 'Gsps', round(d['S']*(d['frames']*86720+8000)/d['ms']['demod']/1e6,1))"
  done
done
