"""Static opcode histogram of a SASS line range: `cuobjdump -sass x.o | python tools/sass_static.py FIRST LAST`
(line numbers among the instruction lines).  Development aid: instruction budget of a loop body before GPU time."""
import sys, re, collections
lines = [l for l in sys.stdin if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l)]
a, b = int(sys.argv[1]), int(sys.argv[2])
c = collections.Counter()
for l in lines[a - 1:b]:
    t = l.split("*/", 1)[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.rstrip(";")
    key = op.split(".")[0]
    if key in ("I2F", "F2F", "MUFU", "LDS", "STS", "LDG", "STG", "BAR"): key = ".".join(op.split(".")[:2])
    c[key] += 1
tot = sum(c.values())
fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
print(f"instructions {tot}  fp64-pipe {fp64}")
print("  ".join(f"{k}:{v}" for k, v in c.most_common()))
