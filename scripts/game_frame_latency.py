"""Latency of a small frame: the SnakeGame frames of tests/golden/snake_frames.npz (35 entries) through the library, wall clock per
ExecuteCollisionDetection and the device-side stage times.  usage: python scripts/game_frame_latency.py [repeats]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inmyroom_vulkan_b200.collision import IMRCD_BUILD_REFERENCE, CollisionDetection, Context, OBBtree  # noqa: E402

Z = np.load(os.path.join(ROOT, "tests", "golden", "snake_frames.npz"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
ctx = Context(0)
trees = [OBBtree.from_primitives(ctx, [(Z[f"mesh{k}.points"], Z[f"mesh{k}.normals"], Z[f"mesh{k}.indices"], 4)], build_mode=IMRCD_BUILD_REFERENCE)
         for k in range(int(Z["n_meshes"][0]))]
cd = CollisionDetection(ctx=ctx)
rows = []
for f in Z["frames"].tolist():
    mesh_ids = np.array([trees[m].mesh_id for m in Z[f"f{f}.mesh"]], np.uint32)
    walls = []
    for r in range(reps):
        cd.Reset()
        cd.add_entries(Z[f"f{f}.cur"], mesh_ids, Z[f"f{f}.callback"], Z[f"f{f}.entity"], Z[f"f{f}.prev"])
        t0 = time.perf_counter()
        cd.ExecuteCollisionDetection()
        walls.append(time.perf_counter() - t0)
    st = cd.stats()
    rows.append((f, 1e3 * np.median(walls[5:]), st))
keys = ["ms_total", "ms_broad", "ms_pair_setup", "ms_traverse", "ms_narrow", "ms_reduce", "ms_response"]
print("frame  wall_ms  " + "  ".join(keys) + "  pairs colliding rays")
for f, w, st in rows:
    print(f"{f:5d}  {w:7.3f}  " + "  ".join(f"{st.get(k, float('nan')):{len(k)}.3f}" for k in keys) + f"  {st['n_pairs']} {st['n_colliding']} {st['n_rays_shot']}")
print("median wall %.3f ms, median device total %.3f ms" % (np.median([w for _, w, _ in rows]), np.median([st["ms_total"] for _, _, st in rows])))
