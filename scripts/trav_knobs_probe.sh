for v in 0 5 7; do for k in 32 8 2; do for b in 1024 128; do echo "== variant $v keep $k backoff $b"; IMRCD_TRAV_VARIANT=$v IMRCD_TRAV_KEEP=$k IMRCD_TRAV_BACKOFF=$b python scripts/shard_probe.py 100000 10 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln)
        if d['world'] in (1,8): print(d['world'], 'trav', d['ms_traverse'], 'total', d['ms_total'], 'lanes', d['lanes_per_iteration'], 'polls', d['idle_polls'], 'qitems', d['queue_items'])
"; done; done; done
