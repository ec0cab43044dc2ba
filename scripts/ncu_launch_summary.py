"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python scripts/ncu_launch_summary.py launches.csv"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict(); tot = 0.0; n = 0
for r in rows[1:]:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    k = r[ix["Kernel Name"]][:64]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us; tot += us; n += 1
print(f"{n} launches, {tot:.1f} us of kernel time (cold-cache, serialised: compare SHARES, not absolutes)")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:66s} n={c:4d} total={t:9.1f} us  mean={t/c:8.1f} us  share={100*t/tot:5.1f}%")
