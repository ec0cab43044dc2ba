"""Quick device-side timing probe for the frame pipeline (not the bench contract; see bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = Context(0)
mesh = scenes.torus(100, 50)
t0 = time.time(); tree = OBBtree(ctx, mesh.positions, mesh.normals, mesh.vertex_ids); t1 = time.time()
print("build ms (device)", tree.build_ms(), "wall", (t1 - t0) * 1e3, "info", tree.info())
scene = scenes.scene_instances(mesh, n, seed=1234, neighbours=4.0)
cd = CollisionDetection(ctx=ctx)
ids = np.full(n, tree.mesh_id, np.uint32)
for r in range(reps):
    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities)
    t0 = time.time(); cd.upload(); cd.run(); cd.fetch(); t1 = time.time()
    st = cd.stats()
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()}), "wall_ms", round((t1 - t0) * 1e3, 3))
print("tri tests/s (device total)", st["n_tri_tests"] / (st["ms_total"] * 1e-3) / 1e9, "G/s ; SAT/s (traverse)", st["n_sat_tests"] / (st["ms_traverse"] * 1e-3) / 1e9, "G/s")

# determinism probe: two runs must give identical combo and hit sets
def snapshot():
    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities)
    cd.upload(); cd.run(); cd.fetch()
    c = cd.combos(); _, h = cd.results(True)
    bp = cd.broad_pairs()
    ck = np.sort(np.ascontiguousarray(np.concatenate([bp[c[:, 0]], c[:, 1:]], 1)).view([('', np.uint32)] * 6).ravel())
    hk = np.sort(np.ascontiguousarray(np.stack([bp[h['pair'], 0], bp[h['pair'], 1], h['tri_first'], h['tri_second']], 1)).view([('', np.uint32)] * 4).ravel())
    return ck, hk
if os.environ.get("PROBE_DET", "1") == "1":
    a = snapshot(); b = snapshot()
    print("combos identical:", np.array_equal(a[0], b[0]), len(a[0]), len(b[0]), " hits identical:", np.array_equal(a[1], b[1]), len(a[1]), len(b[1]))
    if not np.array_equal(a[1], b[1]):
        sa = set(map(tuple, a[1].tolist())); sb = set(map(tuple, b[1].tolist()))
        print("only a:", sorted(sa - sb)[:5], "only b:", sorted(sb - sa)[:5])
