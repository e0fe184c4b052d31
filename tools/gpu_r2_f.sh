#!/bin/bash
# Round 2, call F: ELB variant of the bank kernel (early/late block sums on the AFC warp), spin hand-off in the
# two-warp-per-stream kernel, crossover probes, tests, bench.  Tight timeouts; stop at the first hang.
set -x -o pipefail
mkdir -p gpurun_out
timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-200 || exit 1
timeout 60 python tools/probe.py --streams 1024 --frames 6 --reps 1 --lanes 64 2>&1 | tail -1 | cut -c1-200 || exit 1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_bank or ring or granularity or clock_offset" 2>&1 | tail -5 || exit 1
P="timeout 60 python tools/probe.py --streams 18944 --frames 6 --reps 2"
for ELB in 1 0; do for QX in 1 2 0; do
  OPVD_BANK_ELB=$ELB OPVD_BANK_QX=$QX $P --lanes 96 2>&1 | tail -1 | cut -c1-200
done; done
for L in 32 64; do
  timeout 60 python tools/probe.py --streams 1024 --frames 25 --reps 2 --lanes $L 2>&1 | tail -1 | cut -c1-200
done
for S in 1280 1536 1792; do for L in 32 96; do
  timeout 60 python tools/probe.py --streams $S --frames 12 --reps 2 --lanes $L 2>&1 | tail -1 | cut -c1-200
done; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_f.json 2> gpurun_out/bench_r02_f.err
tail -c 1500 gpurun_out/bench_r02_f.json; tail -5 gpurun_out/bench_r02_f.err
timeout 240 ncu --set full --clock-control none --import-source on -k regex:demod_bank -c 1 -f -o gpurun_out/prof_bank_r02_f \
    python tools/probe.py --streams 18944 --frames 2 --reps 1 --lanes 96 > gpurun_out/ncu_bank_r02_f.log 2>&1
tail -2 gpurun_out/ncu_bank_r02_f.log
