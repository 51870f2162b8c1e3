"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list in order (us, grid), plus totals by kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = d["Kernel Name"]
    if flt and flt not in k:
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(d["Metric Unit"], 1)
    if "--seq" in sys.argv:
        print(f"{d['ID']:>5s} {k[:60]:60s} {v:8.1f} us  grid {d.get('Grid Size')}")
    a = agg.setdefault(k[:60], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={a[0]:4d} tot={a[1]:9.1f} us avg={a[1] / a[0]:8.1f}  {100 * a[1] / tot:5.1f}%")
print(f"total {tot:.1f} us")
