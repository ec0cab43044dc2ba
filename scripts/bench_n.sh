#!/bin/bash
# bench.py on N GPUs of this box, launched as the driver launches it; the line lands in gpurun_out/r2_bench_c3_n$N.json
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err || tail -5 gpurun_out/r2_bench_c3_n$N.err
python - <<PY
import json
for ln in open("gpurun_out/r2_bench_c3_n$N.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("N=$N", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["ms_per_step"], 4), {k: round(v, 4) for k, v in d["frame"].items() if k.startswith("ms_")}, d.get("clocks"), d["config"]["parallelism"][-60:])
        r = d.get("response")
        if r: print("   response", round(r["ms_response"], 4), "frame with response", round(r["ms_frame_with_response"], 4))
PY
