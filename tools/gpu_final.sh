#!/bin/bash
# last regression of the round on the committed tree: smoke, all GPU tests, one bench line
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'bank', d['channel_bank']['value'], 'bank demod frac', d['channel_bank']['roofline']['frac'], 'spot', d['parity_spotcheck'])"
