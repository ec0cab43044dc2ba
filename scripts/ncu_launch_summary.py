"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict(); n = 0
for r in rows[1:]:
    if len(r) < len(hdr): continue
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    agg.setdefault(r[ix["Kernel Name"]].split("(")[0][:64], []).append(v); n += 1
tot = sum(sum(v) for v in agg.values())
print(f"{n} launches, {tot:.1f} us of kernel time (cold-cache, serialised: compare SHARES, not absolutes)")
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print(f"{k:66s} n={len(v):4d} total={sum(v):9.1f} us  mean={sum(v)/len(v):8.1f} us  share={100*sum(v)/tot:5.1f}%")
