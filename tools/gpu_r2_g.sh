#!/bin/bash
# Round 2, call G: on-time blocks in two passes (OB2), conversion splits, bank CLI ingest tests.
set -x -o pipefail
mkdir -p gpurun_out
timeout 60 python tools/probe.py --streams 4096 --frames 3 --reps 1 --lanes 96 2>&1 | tail -1 | cut -c1-200 || exit 1
P="timeout 60 python tools/probe.py --streams 18944 --frames 6 --reps 2"
for OB in 1 0; do for QX in 1 2 0; do
  OPVD_BANK_OB2=$OB OPVD_BANK_QX=$QX $P --lanes 96 2>&1 | tail -1 | cut -c1-200
done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bank_cli or modem or one_bank" 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
