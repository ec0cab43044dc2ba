#!/bin/bash
# A/B of the traversal's knobs on one GPU: per-rank stage times of a C3 frame at world sizes 1, 2, 4, 8 (scripts/shard_probe.py)
# usage: scripts/trav_ab.sh "IMRCD_TRAV_X=1 IMRCD_TRAV_Y=2" "..." ...
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 60 python scripts/shard_probe.py 100000 10 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln)
        print(d['world'], 'trav', d['ms_traverse'], 'narrow', d['ms_narrow'], 'reduce', d['ms_reduce'], 'broad', d['ms_broad'], 'total', d['ms_total'], 'iters', d['warp_iterations'], 'lanes', d['lanes_per_iteration'], 'polls', d['idle_polls'], 'qitems', d['queue_items'], 'sat', d['sat'])
    elif 'rror' in ln: print(ln.rstrip())
"
done
