"""Attribute executed warp instructions to CUDA source lines.

usage: ncu_lines.py <ncu source-page csv (sass)> <object .o> <mangled kernel substring> [per-unit divisor]
Joins the ncu SASS listing (in order) with `nvdisasm -g` line annotations of the same
function (the CLI's cuda view carries no metrics)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

src_csv, obj, kname = sys.argv[1:4]
per = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True,
               stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True,
                     text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iex = hdr.index("Instructions Executed")
ex = [int(r[iex]) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
print("sass in ncu:", len(ex), "sass in nvdisasm:", len(lines))
n = min(len(ex), len(lines))
agg = defaultdict(int)
for i in range(n):
    agg[lines[i][0]] += ex[i]
tot = sum(ex)
srcs = {}
def getline(f, ln):
    for root in ("flam3_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srcs:
                srcs[p] = open(p).read().splitlines()
            if 0 < ln <= len(srcs[p]):
                return srcs[p][ln - 1].strip()[:90]
    return ""
for (f, ln), e in sorted(agg.items(), key=lambda kv: -kv[1])[:70]:
    print("%6.2f%% %8.2f  %s:%d  %s" % (100.0 * e / tot, e / per, f, ln, getline(f, ln)))
