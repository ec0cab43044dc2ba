"""Summarise `ncu --page source --csv` output: stall reasons, hottest SASS instructions, opcode mix.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name K > src.csv ; python scripts/ncu_source_summary.py src.csv [top]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[rows.index(hdr) + 1:]:          # first table only (ncu repeats the table per view)
    if "Kernel Name" in r[:1]:
        break
    if len(r) == len(hdr) and r[ix["# Samples"]].isdigit():
        data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data)
ninst = sum(int(r[ix["Instructions Executed"]]) for r in data)
nthr = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
print(f"samples {tot}  SASS lines {len(data)}  warp-inst {ninst}  thread-inst {nthr}  avg active {nthr/max(ninst,1):.2f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print("stalls:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
    s = int(r[ix["# Samples"]])
    why = {h[6:]: r[ix[h]] for h in stalls if int(r[ix[h]] or 0) > s * 0.25}
    print(f"{s:6d} {100*s/tot:4.1f}%  {r[ix['Source']].strip()[:70]:70s} exec {r[ix['Instructions Executed']]:>9s} thr {r[ix['Avg. Threads Executed']]:>5s} {why}")
c = Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
print("opcode mix (warp-inst):", ", ".join(f"{k} {100*v/ninst:.1f}%" for k, v in c.most_common(16)))
