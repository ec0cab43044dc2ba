#!/usr/bin/env python
"""One rank's share (set_shard(0, world)) of the C3 frame, a few runs: the target of ncu captures at a small per-rank size.
usage: python scripts/shard_one.py [world] [bodies] [runs]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inmyroom_vulkan_b200 import scenes                                  # noqa: E402
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree   # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bodies = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ctx = Context(0)
scene = scenes.scene_static_vs_bodies(scenes.uv_sphere(66, 65), bodies, seed=2026, body_scale=(0.2, 0.5))
trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
cd = CollisionDetection(ctx=ctx)
cd.set_shard(0, world)
cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities); cd.upload()
for _ in range(runs):
    cd.run()
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in cd.stats().items()})
