#!/bin/bash
# Round 2, call X: e2e leg against the number of time tiles and the ring depth, and the copy-only rate of the same pushes.
set -x -o pipefail
mkdir -p gpurun_out
for CFG in "8 2" "12 2" "16 2" "24 2"; do set -- $CFG
  timeout 200 python bench.py --steps 3 --warmup 3 --no-bank --no-cpu-baseline --e2e-tiles $1 --e2e-ring-tiles $2 2>gpurun_out/x.err | tail -1 > gpurun_out/x_$1_$2.json
  python -c "
import json
e = json.loads(open('gpurun_out/x_$1_$2.json').read())['e2e']
print('tiles', $1, 'ring', $2, 'tiled_push', e['h2d_tiled_push_gbs'], 'e2e', e['value'], 'h2d', e['h2d_gbs'], 'ceiling', e['h2d_ceiling_gbs'], 'frac', e['frac_of_h2d_ceiling'])" || tail -5 gpurun_out/x.err
done
