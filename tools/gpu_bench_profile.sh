#!/bin/bash
# One gpurun call: bench (both arms), ncu launch list of the bench command, one full ncu capture of the
# demod kernel on a short bank.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
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
