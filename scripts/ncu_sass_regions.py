"""Cut a kernel's SASS (ncu --page source --csv) into runs of equal execution count: warp instructions, active lanes and stall samples per run.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv ; python scripts/ncu_sass_regions.py src.csv [units]
`units` = how many times the kernel's outer unit (tile, visit batch, ...) ran, to print executions per unit."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[rows.index(hdr) + 1:]:
    if "Kernel Name" in r[:1]:
        break
    if len(r) == len(hdr) and r[ix["# Samples"]].isdigit():
        data.append(r)
runs = []
for r in data:
    e = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]]); s = int(r[ix["# Samples"]])
    key = e / units
    if runs and abs(runs[-1][0] - key) <= max(0.15 * runs[-1][0], 0.3 / max(units, 1) if units == 1 else 0.3):
        runs[-1][1] += 1; runs[-1][2] += e; runs[-1][3] += t; runs[-1][4] += s
    else:
        runs.append([key, 1, e, t, s, r[ix["Source"]].strip()[:48]])
tot_e = sum(r[2] for r in runs); tot_s = sum(r[4] for r in runs); tot_t = sum(r[3] for r in runs)
print(f"warp instructions {tot_e}, average active lanes {tot_t / max(tot_e, 1):.2f}")
print("exec/unit  #sass  warp-inst%  lanes  samples%  first instruction")
for k, n, e, t, s, src in runs:
    if e / tot_e > 0.004:
        print(f"{k:9.2f} {n:6d} {100 * e / tot_e:9.1f}% {t / max(e, 1):6.1f} {100 * s / max(tot_s, 1):8.1f}%  {src}")
