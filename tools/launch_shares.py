"""Sum gpu__time_duration per kernel name from an ncu --csv launch list (long format)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None
tot = collections.OrderedDict()
cnt = collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0][:70]
    v = float(d["Metric Value"].replace(",", ""))
    unit = d.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] = tot.get(name, 0.0) + v * scale
    cnt[name] += 1
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-72s %5d launches %10.3f ms  %5.1f %%" % (k, cnt[k], v, 100 * v / s))
print("total %.3f ms" % s)
