#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01j}
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -1 gpurun_out/bench_$TAG.json | cut -c1-3000; tail -5 gpurun_out/bench_$TAG.err
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
