#!/usr/bin/env python
"""Timeline of k_traverse (diagnostic): IMRCD_TRAV_TRACE=1 python scripts/trav_trace.py [world] [bodies]
Per 10-us bin: warps inside an iteration, lanes with an item; then the per-warp summary."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["IMRCD_TRAV_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inmyroom_vulkan_b200 import scenes                                  # noqa: E402
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree   # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bodies = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
ctx = Context(0)
scene = scenes.scene_static_vs_bodies(scenes.uv_sphere(66, 65), bodies, seed=2026, body_scale=(0.2, 0.5))
trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
cd = CollisionDetection(ctx=ctx)
cd.set_shard(0, world)
cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities); cd.upload()
for _ in range(3):
    cd.run()
st = cd.stats()
fn = ctx.lib.imrcd_debug_trav_trace
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
nw = C.c_uint32()
fn(ctx.h, None, 0, C.byref(nw))
buf = np.zeros((nw.value, 256, 4), np.uint32)
fn(ctx.h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(nw))
used = buf[:, :, 1] != 0
t0 = buf[:, :, 0][used].min()
b = (buf[:, :, 0].astype(np.int64) - int(t0)) / 1e3
e = (buf[:, :, 1].astype(np.int64) - int(t0)) / 1e3
print(f"world {world}: ms_traverse {st['ms_traverse']:.3f}, {st['n_sat_tests']} SAT, {st['n_warp_iterations']} iterations, {nw.value} warps, traced {int(used.sum())}")
end = e[used].max()
print(f"span of traced iterations: {end:.1f} us; iteration length: median {np.median((e - b)[used]):.2f} us, p90 {np.percentile((e - b)[used], 90):.2f}, max {(e - b)[used].max():.2f}")
bins = np.arange(0, end + 10, 10.0)
for lo in bins:
    inside = used & (b < lo + 10) & (e > lo)
    print(f"  t={lo:6.0f} us  warps busy {int(inside.any(1).sum()):5d}  iterations {int(inside.sum()):6d}  mean lanes {buf[:, :, 2][inside].mean() if inside.any() else 0:5.1f}  mean deque {buf[:, :, 3][inside].mean() if inside.any() else 0:6.1f}")
iters = used.sum(1)
print("iterations per warp: min %d median %d max %d; warps that never worked: %d" % (iters.min(), np.median(iters), iters.max(), int((iters == 0).sum())))
# iteration length by number of node pairs in the iteration (a warp with few pairs deals their axes to several lanes)
ln = buf[:, :, 2]
dur = (e - b)
for lo, hi in ((1, 2), (3, 4), (5, 8), (9, 16), (17, 24), (25, 32)):
    m = used & (ln >= lo) & (ln <= hi)
    if m.any():
        late = m & (b > 0.6 * end)
        print(f"pairs {lo:2d}-{hi:2d}: {int(m.sum()):7d} iterations, median {np.median(dur[m]):.2f} us, p10 {np.percentile(dur[m], 10):.2f}, p90 {np.percentile(dur[m], 90):.2f}"
              + (f"; in the last 40 % of the kernel: {int(late.sum())} iterations, median {np.median(dur[late]):.2f} us" if late.any() else ""))
