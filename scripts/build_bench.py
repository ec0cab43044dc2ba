"""C4 (BASELINE.json configs[3]): OBB-tree build throughput over one synthetic 10M-triangle mesh (build-only path).
   python scripts/build_bench.py [n_million] [reps] [--reference-cpu K]   -> one JSON line
Checks the result with size-independent properties (leaf order is a permutation, leaves tile [0,n), <= 4 triangles per leaf,
children tile their parent, sampled boxes contain their triangles) and times the reference's own build on a bounded sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import Context, OBBtree, IMRCD_BUILD_MORTON, IMRCD_BUILD_REFERENCE

nm = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nx = int(round((nm * 1e6) ** 0.5)); nz = nx // 2          # 2 * nx * nz = nx^2 triangles
mesh = scenes.grid_sheet(nx, nz, 1500.0, 750.0, bump=40.0)
n = mesh.n_tri
ctx = Context(0)
ms = []
for r in range(reps + 1):
    t0 = time.perf_counter()
    tree = OBBtree(ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
    wall = time.perf_counter() - t0
    if r:
        ms.append((tree.build_ms(), wall * 1e3))
dev = float(np.median([m[0] for m in ms])); wall = float(np.median([m[1] for m in ms]))
out = {"workload": f"C4: displaced grid {nx}x{nz}, {n} triangles, GPU Morton build", "triangles": n, "build_ms_device": dev, "build_ms_wall_incl_h2d": wall,
       "mtri_per_s_device": n / dev / 1e3, "algorithmic_bytes_per_tri": 268, "achieved_gbs": n * 268 / (dev * 1e-3) / 1e9}
# size-independent checks on the last tree
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
if n <= 3_000_000 or "--check" in sys.argv:
    from test_gpu_build import check_tree_structure, check_boxes_contain
    flat = tree.export()
    lo, hi = check_tree_structure(flat, mesh)
    check_boxes_contain(flat, lo, hi, sample=300)
    out["checked"] = "structure + sampled containment ok"
    out["tree_vertices"] = int(flat.nv)
if "--reference-cpu" in sys.argv:
    k = int(sys.argv[sys.argv.index("--reference-cpu") + 1])
    import bench                                  # the CPU checker is reached only through bench.py's cpu_baseline doorway
    orc, _ = bench.load_cpu_checker()
    sub = scenes.grid_sheet(int((k / 2) ** 0.5 * 2 ** 0.5), int((k / 2) ** 0.5 / 2 ** 0.5), 1500.0, 750.0, bump=40.0)
    t0 = time.perf_counter(); orc.tree_build(sub.positions, sub.normals, sub.vertex_ids); dt = time.perf_counter() - t0
    out["cpu_reference"] = {"kind": orc.kind, "triangles": sub.n_tri, "seconds": dt, "mtri_per_s": sub.n_tri / dt / 1e6, "cores": 1}
    t0 = time.perf_counter(); t = OBBtree(ctx, sub.positions, sub.normals, sub.vertex_ids, build_mode=IMRCD_BUILD_REFERENCE); dt = time.perf_counter() - t0
    out["gpu_reference_mode"] = {"triangles": sub.n_tri, "build_ms_device": t.build_ms(), "mtri_per_s": sub.n_tri / t.build_ms() / 1e3}
print(json.dumps(out))
