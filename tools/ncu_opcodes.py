"""Opcode histogram (executed warp instructions) from an ncu source-page CSV."""
import csv
import sys
from collections import Counter

path, per = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = rows[1]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
data = [(r[isrc].strip(), int(r[iex])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot = sum(d[1] for d in data)
print("total warp instructions", tot, "distinct SASS", len(data), "per unit", tot / per)
c = Counter()
for s, e in data:
    toks = s.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    c[".".join(op.split(".")[:2]) if op.startswith(("LDL", "STL", "LDS", "STS", "RED", "MUFU")) else op.split(".")[0]] += e
for op, n in c.most_common(30):
    print("%-14s %6.2f%%  %8.2f per unit" % (op, 100 * n / tot, n / per))
