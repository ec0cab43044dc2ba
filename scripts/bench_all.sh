#!/bin/bash
# every workload of bench.py once on one GPU (+ the reference arm of the default one); lines land in gpurun_out/r2_bench_*.json
mkdir -p gpurun_out
for w in c3 c1 c4 c5; do
  timeout 300 python bench.py --workload $w --steps ${STEPS:-10} --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err || tail -5 gpurun_out/r2_bench_$w.err
done
# C2's single-thread reference sample takes 80 s of CPU (its sweep is quadratic on 4,096 overlapping tori): skipped here, see profiles/r2_bench_c2.json of the round's first session
timeout 300 python bench.py --workload c2 --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err || tail -5 gpurun_out/r2_bench_c2.err
timeout 300 python bench.py --workload c3 --trees reference --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_reftrees.json 2> gpurun_out/r2_bench_c3_reftrees.err || tail -5 gpurun_out/r2_bench_c3_reftrees.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_c3_refarm.json 2> gpurun_out/r2_bench_c3_refarm.err || tail -5 gpurun_out/r2_bench_c3_refarm.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_c*.json")):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f.split("/")[-1], round(d["ms_per_step"], 4), "ms", f'{d["value"]:.4g}', d["unit"], "| e2e", round(d["e2e"].get("ms_per_step", 0), 4), "ms",
                  "| roofline", d.get("roofline", {}).get("kernel"), round(d.get("roofline", {}).get("frac", 0) or 0, 4), "| cpu", f'{d.get("cpu_baseline", {}).get("value", 0):.4g}',
                  "| same_work", f'{d.get("same_work", {}).get("value", 0):.4g}' if isinstance(d.get("same_work", {}).get("value", 0), (int, float)) else d.get("same_work"), "| single", f'{d.get("single_thread", {}).get("value", 0):.4g}')
            if "frame" in d: print("   ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["frame"].items()})
PY
