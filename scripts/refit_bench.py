"""C5 (BASELINE.json configs[4]): 256 characters x ~20k triangles re-posed every frame: new triangle positions -> batched OBB-tree refit -> collide.
   python scripts/refit_bench.py [characters] [frames] [--reference-cpu K]   -> one JSON line
The re-pose itself (skinning / morph targets) is the renderer's compute shader in the reference (dynamicMeshShader_glsl.comp:99-145) and out of
scope; here the re-posed positions are produced on the host by a smooth deformation and handed to imrcd_mesh_update_positions.  The reference has no
refit: its CPU time for the same frame is REBUILDING every tree (OBBtree::OBBtree) from the re-posed triangles, timed on K characters."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree, refit_meshes

n_char = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 256
frames = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 5


def repose(pos, phase):
    p = pos.reshape(-1, 3).astype(np.float64)
    ang = 0.6 * np.sin(phase) * p[:, 2] + 0.3 * np.cos(2 * phase) * p[:, 0]
    c, s = np.cos(ang), np.sin(ang)
    q = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1], p[:, 2] + 0.25 * np.sin(phase + 2.0 * p[:, 0])], 1)
    return np.ascontiguousarray(q.astype(np.float32).reshape(-1, 9))


base = scenes.torus(142, 71)                      # 20,164 triangles
ground = scenes.grid_sheet(64, 64, 60.0, 60.0, bump=0.5)
ctx = Context(0)
t0 = time.perf_counter()
trees = [OBBtree(ctx, base.positions, base.normals, base.vertex_ids) for _ in range(n_char)]
g_tree = OBBtree(ctx, ground.positions, ground.normals, ground.vertex_ids)
build_s = time.perf_counter() - t0
sc = scenes.scene_instances(base, n_char, seed=7, neighbours=6.0)
mats = np.concatenate([sc.matrices, np.eye(4, dtype=np.float32).reshape(1, 16)])
mesh_ids = np.array([t.mesh_id for t in trees] + [g_tree.mesh_id], np.uint32)
cb = np.ones(n_char + 1, np.uint8); ents = np.arange(1, n_char + 2, dtype=np.uint32)
cd = CollisionDetection(ctx=ctx)
# a few distinct poses are precomputed on the host (the renderer would produce them on the device); character k uses pose (k + frame) % P
P = 8
poses = [repose(base.positions, 0.8 * k) for k in range(P)]
rows = []
for f in range(frames + 1):
    t0 = time.perf_counter()
    for k, t in enumerate(trees):
        t.update_positions(poses[(k + f) % P])
    t1 = time.perf_counter()
    refit_ms = refit_meshes(ctx)
    t2 = time.perf_counter()
    cd.Reset(); cd.add_entries(mats, mesh_ids, cb, ents); cd.ExecuteCollisionDetection()
    t3 = time.perf_counter()
    st = cd.stats()
    if f:
        rows.append(dict(upload_ms=(t1 - t0) * 1e3, refit_ms_device=refit_ms, refit_ms_wall=(t2 - t1) * 1e3, collide_ms_device=st["ms_total"], collide_ms_wall=(t3 - t2) * 1e3,
                         tri_tests=st["n_tri_tests"], hits=st["n_hits"], colliding=st["n_colliding"], pairs=st["n_pairs"], sat_tests=st["n_sat_tests"],
                         ms_broad=st["ms_broad"], ms_traverse=st["ms_traverse"], ms_narrow=st["ms_narrow"], ms_reduce=st["ms_reduce"]))
med = {k: float(np.median([r[k] for r in rows])) for k in rows[0]}
n_tri = base.n_tri * n_char
out = {"workload": f"C5: {n_char} characters x {base.n_tri} triangles re-posed per frame + a {ground.n_tri}-triangle ground, refit then collide", "triangles_refit": n_tri,
       "frames": frames, **med, "refit_mtri_per_s_device": n_tri / med["refit_ms_device"] / 1e3,
       "refit_algorithmic_bytes_per_tri": 72, "refit_achieved_gbs": n_tri * 72 / (med["refit_ms_device"] * 1e-3) / 1e9, "initial_build_s": build_s}
if "--reference-cpu" in sys.argv:
    k = int(sys.argv[sys.argv.index("--reference-cpu") + 1])
    import bench                                  # the CPU checker is reached only through bench.py's cpu_baseline doorway
    orc, _ = bench.load_cpu_checker()
    t0 = time.perf_counter()
    for c in range(k):
        orc.tree_build(poses[c % P], base.normals, base.vertex_ids)
    dt = time.perf_counter() - t0
    out["cpu_reference_rebuild"] = {"kind": orc.kind, "characters": k, "seconds": dt, "mtri_per_s": base.n_tri * k / dt / 1e6, "cores": 1,
                                    "note": "the reference has no refit: it would rebuild every tree from the re-posed triangles"}
print(json.dumps(out))
