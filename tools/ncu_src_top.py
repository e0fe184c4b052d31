"""Top stall sites of an `ncu --page source --csv --print-source sass` export, plus per-region totals.
usage: ncu_src_top.py export.csv [N] [lo-hi ...]   (regions are instruction index ranges, 1-based)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def val(r, k):
    try: return int(r[ci[k]] or 0)
    except ValueError: return 0
tot = sum(val(r, "# Samples") for r in body)
print("total samples", tot, "instructions", len(body))
for rg in sys.argv[3:]:
    lo, hi = map(int, rg.split("-"))
    sub = body[lo - 1:hi]
    s = sum(val(r, "# Samples") for r in sub)
    ex = sum(val(r, "Instructions Executed") for r in sub)
    by = {k: sum(val(r, k) for r in sub) for k in stalls}
    top = sorted(by.items(), key=lambda kv: -kv[1])[:7]
    print(f"region {rg}: samples {s} ({100*s/tot:.1f}%) executed {ex}  " + " ".join(f"{k[6:]}={v}" for k, v in top))
idx = sorted(range(len(body)), key=lambda i: -val(body[i], "# Samples"))[:N]
for i in sorted(idx):
    r = body[i]
    by = sorted(((val(r, k), k[6:]) for k in stalls), reverse=True)[:3]
    print(f"{i+1:5d} {val(r,'# Samples'):6d} {r[ci['Source']][:70]:70s} " + " ".join(f"{k}={v}" for v, k in by if v))
