"""Frames only, for ncu: build the workload's trees, warm up, then run K frames between cudaProfilerStart/Stop.
usage: ncu --profile-from-start off ... python scripts/profile_frame.py [c3|c2] [bodies] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
bodies = int(sys.argv[2]) if len(sys.argv) > 2 else (100000 if name == "c3" else 4096)
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
moved = len(sys.argv) > 4 and sys.argv[4] == "moved"      # every dynamic body moved since the last frame: the response stage runs
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
ctx = Context(0, stream.cuda_stream)
scene, desc = bench.make_workload(name, bodies)
trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
mesh_ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
cd = CollisionDetection(ctx=ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
prev = None
if moved:
    prev = scene.matrices.copy()
    ns = len(scene.meshes) - 1 if name == "c3" else 0
    prev[ns:, 12:15] += (np.random.default_rng(1).normal(size=(scene.n_entries - ns, 3)) * 0.02).astype(np.float32)
cd.Reset(); cd.add_entries(scene.matrices, mesh_ids, scene.should_callback, scene.entities, prev); cd.upload()
for _ in range(3):
    cd.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(frames):
    with torch.cuda.stream(stream):
        flush.zero_()
    cd.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
st = cd.stats()
print(desc)
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
