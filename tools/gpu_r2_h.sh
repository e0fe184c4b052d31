#!/bin/bash
# Round 2, call H: the evidence run.  Sanity probe, full GPU tests, smoke, both bench arms, ncu launch list of the bench
# command, memcheck over both demodulator kernels.  Tight timeouts; stop at the first hang.
set -x -o pipefail
mkdir -p gpurun_out
TAG=${1:-r02_h}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
nproc
timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-200 || exit 1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -1 gpurun_out/bench_ref_$TAG.json | cut -c1-600
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 2500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
grep -c . gpurun_out/launches_$TAG.csv
export PATH=/usr/local/cuda/bin:$PATH
for L in 96 32; do
  timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python tools/probe.py --streams 80 --frames 2 --reps 1 --lanes $L > gpurun_out/memcheck_${TAG}_$L.log 2>&1
  echo "memcheck lanes=$L: $(grep -E 'ERROR SUMMARY' gpurun_out/memcheck_${TAG}_$L.log)"
done
