#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list, ncu full capture of the demod kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
nproc; free -g | head -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for L in 1 2 4; do
  timeout 300 python tools/probe.py --streams 1024 --frames 8 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-400
done
for L in 1 2; do
  timeout 300 python tools/probe.py --streams 16384 --frames 3 --reps 1 --lanes $L 2>&1 | tail -1 | cut -c1-400
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -1 gpurun_out/bench_ref_$TAG.json
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -1 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
grep -c . gpurun_out/launches_$TAG.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_.*kernel -c 1 -f -o gpurun_out/prof_demod_$TAG \
    python tools/probe.py --streams 1024 --frames 4 --reps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
