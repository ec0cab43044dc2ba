"""Run a few frames of a bench workload for ncu captures (not a benchmark): python scripts/profile_frame.py c3 100000 3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
bodies = int(sys.argv[2]) if len(sys.argv) > 2 else (100000 if name == "c3" else 4096)
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scene, desc = bench.make_workload(name, bodies)
ctx = Context(0)
trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
cd = CollisionDetection(ctx=ctx)
for f in range(frames):
    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities)
    cd.ExecuteCollisionDetection()
    st = cd.stats()
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
