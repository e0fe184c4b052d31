"""Summarise an `ncu --page source --csv --print-source sass` export: executed warp-instructions and
stall samples per opcode (development aid)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
ex = collections.Counter(); st = collections.Counter()
tot = 0; tots = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ci["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("I2F","F2F","MUFU","SHFL","LDS","STS","LDG","STG","DSETP","BRA")) else op.split(".")[0]
    n = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    ex[op] += n; st[op] += s; tot += n; tots += s
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"total executed {tot}  per-unit {tot/div:.1f}   samples {tots}")
for op, n in ex.most_common(40):
    print(f"{op:14s} {n/div:10.1f} {100*n/tot:6.2f}%   stall-samples {100*st[op]/max(tots,1):6.2f}%")
