"""Dump the SASS (with execution counts) that nvdisasm attributes to a source line.
usage: ncu_line_sass.py <ncu sass csv> <object .o> <kernel substring> <file:line>"""
import csv, os, re, subprocess, sys, tempfile
src_csv, obj, kname, where = sys.argv[1:5]
wf, wl = where.split(":"); wl = int(wl)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---") or (l.startswith(".text.") and l.rstrip().endswith(":")): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: lines.append((cur, m.group(1), m.group(2)))
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; iex = hdr.index("Instructions Executed")
ex = [int(r[iex]) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
for (loc, addr, txt), e in zip(lines, ex):
    if loc == (wf, wl) and e > 0:
        print("%s %12d  %s" % (addr, e, txt))
