#!/bin/bash
# Round 2: early/late block sums on the AFC warp (ELB) against bank size: where it pays.
set -x -o pipefail
mkdir -p gpurun_out
P="timeout 90 python tools/probe.py --frames 6 --reps 2 --lanes 96"
OPVD_BANK_ELB=1 $P --streams 4096 2>&1 | tail -1 | cut -c1-200 || exit 1
for S in 1536 4736 9472 14208 18944; do for E in 0 1; do
  OPVD_BANK_ELB=$E $P --streams $S 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('S', d['S'], 'elb', $E, 'demod', round(d['ms']['demod'], 3))"
done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "96 or shapes_at_scale or bank_cli or config1_full or config3" 2>&1 | tail -3
