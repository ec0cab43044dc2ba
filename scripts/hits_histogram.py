import sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree
ctx = Context(0)
for name, bodies in (("c3", 100000), ("c2", 4096)):
    scene, desc = bench.make_workload(name, bodies)
    trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
    cd = CollisionDetection(ctx=ctx)
    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities); cd.ExecuteCollisionDetection()
    ep, hits = cd.results(True)
    nh = ep["n_hits"]
    print(name, "colliding", len(ep), "hits", len(hits), "per-pair hits: mean", nh.mean(), "median", np.median(nh), "p99", np.percentile(nh, 99), "max", nh.max(),
          "pairs>1024:", (nh > 1024).sum(), ">4096:", (nh > 4096).sum())
    # distinct triangles per pair side
    key = hits["pair"].astype(np.uint64) << np.uint64(32)
    ua = np.unique(key | hits["tri_first"]); ub = np.unique(key | hits["tri_second"])
    print("  distinct (pair,triA)", len(ua), "(pair,triB)", len(ub))
