"""Cycle budget of a kernel's per-symbol loop from an `ncu --page source --csv --print-source sass` export: the
instructions executed about once per symbol (warp level), in buckets of 25, with the warp-state samples of each bucket
scaled to cycles per symbol (samples are uniform in time, so share of samples = share of the warp's time).
usage: ncu_loop_budget.py export.csv N_SYMBOL_EXECUTIONS CYCLES_PER_SYMBOL"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
nsym, cyc = float(sys.argv[2]), float(sys.argv[3])
hdr = rows[1]; ci = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
def val(r, k):
    try: return int(r[ci[k]] or 0)
    except ValueError: return 0
def opc(r):
    t = r[ci["Source"]].split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
tot = sum(val(r, "# Samples") for r in body)
hot = [i for i, r in enumerate(body) if 0.8 * nsym < val(r, "Instructions Executed") < 1.3 * nsym]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
share = sum(val(body[i], "# Samples") for i in hot) / tot
print(f"{len(hot)} instructions executed once per symbol (index {hot[0] + 1}..{hot[-1] + 1}), {100 * share:.1f} % of all warp samples "
      f"= {cyc * share:.0f} of {cyc:.0f} cycles per symbol")
allby = collections.Counter()
for b in range(0, len(hot), 25):
    idx = hot[b:b + 25]
    s = sum(val(body[i], "# Samples") for i in idx)
    by = collections.Counter()
    for i in idx:
        for k in stalls:
            by[k[6:]] += val(body[i], k)
    allby.update(by)
    ops = collections.Counter(opc(body[i]) for i in idx)
    print(f"{idx[0] + 1:5d}-{idx[-1] + 1:5d} {cyc * s / tot:5.0f} cycles  " + " ".join(f"{k}={cyc * v / tot:.0f}" for k, v in by.most_common(4))
          + "   | " + " ".join(f"{k}:{v}" for k, v in ops.most_common(6)))
print("whole loop: " + " ".join(f"{k}={cyc * v / tot:.0f}" for k, v in allby.most_common(8)))
