#!/bin/bash
# Round 2, call Y: timeline of the e2e leg (OPVD_TRACE): where the copy engine idles.
set -x -o pipefail
mkdir -p gpurun_out
OPVD_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-bank --no-cpu-baseline > gpurun_out/y.json 2> gpurun_out/y.err
grep "opvd trace" gpurun_out/y.err | tail -32
