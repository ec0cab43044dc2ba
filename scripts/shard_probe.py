#!/usr/bin/env python
"""One rank's share of a sharded C3 frame on ONE GPU (imrcd_frame_set_shard(0, N)): stage times and traversal diagnostics per N.
What an N-GPU run spends per rank, without the N GPUs:  python scripts/shard_probe.py [bodies] [reps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inmyroom_vulkan_b200 import scenes                                  # noqa: E402
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree   # noqa: E402

bodies = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = Context(0)
scene = scenes.scene_static_vs_bodies(scenes.uv_sphere(66, 65), bodies, seed=2026, body_scale=(0.2, 0.5))
trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
cd = CollisionDetection(ctx=ctx)
for world in (1, 2, 4, 8):
    cd.set_shard(0, world)
    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities); cd.upload()
    cd.run()
    acc = {}
    for _ in range(reps):
        cd.run()
        st = cd.stats()
        for k, v in st.items():
            if k.startswith("ms_"):
                acc[k] = acc.get(k, 0.0) + v / reps
    st = cd.stats()
    out = {"world": world, **{k: round(v, 4) for k, v in acc.items()}, "pairs": st["n_pairs"], "sat": st["n_sat_tests"], "tri_tests": st["n_tri_tests"],
           "hits": st["n_hits"], "entries_local": st["n_entries_local"], "launches": st["total_launches"], "queue_items": st["n_queue_items"],
           "warp_iterations": st["n_warp_iterations"], "lanes_per_iteration": round(st["n_sat_tests"] / max(st["n_warp_iterations"], 1), 2),
           "idle_polls": st["trav_idle_polls"], "busy_cycles_per_warp": st["trav_busy_cycles"]}
    print(json.dumps(out))
