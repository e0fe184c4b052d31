"""Summarise an `ncu --page source --csv --print-source sass` export of a warp-specialised kernel: stall samples per
barrier-delimited code region with their stall reasons, and the hottest instructions (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
S = [int(r[ci["# Samples"]] or 0) for r in data]
E = [int(r[ci["Instructions Executed"]] or 0) for r in data]
tot = sum(S)
print(len(data), "instructions,", tot, "samples")
cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
def agg(a, b):
    out = {c: sum(int(r[ci[c]] or 0) for r in data[a:b]) for c in cols}
    t = sum(out.values()) or 1
    return {k[6:]: round(100 * v / t, 1) for k, v in sorted(out.items(), key=lambda kv: -kv[1]) if v > 0.03 * t}
bars = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[ci["Source"]]]
prev = 0
for b in bars + [len(data)]:
    if sum(S[prev:b]) > 0.002 * tot:
        print(prev, b, sum(S[prev:b]), f'{100*sum(S[prev:b])/tot:.1f}%', 'exec', max(E[prev:b]), agg(prev, b))
    prev = b
for i in sorted(sorted(range(len(data)), key=lambda i: -S[i])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]):
    r = data[i]
    st = {c[6:]: int(r[ci[c]] or 0) for c in cols if int(r[ci[c]] or 0) > 0.0007 * tot}
    print(i, r[ci["Source"]].strip()[:60].ljust(60), S[i], f'{100*S[i]/tot:.1f}%', E[i], st)
