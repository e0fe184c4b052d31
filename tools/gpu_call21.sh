#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "granularity or compaction or cli_dropin or bank_cli or init_offset" 2>&1 | tail -3
for T in 1 2 4 8; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-bank --no-cpu-baseline --e2e-tiles $T 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tiles', $T, 'e2e', d['e2e']['value'], 'value', d['value'])"
done
