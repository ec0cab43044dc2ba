#!/usr/bin/env python
"""Diagnostic: every box of a Morton-built tree against the triangles below it (FP64 check of the stored FP32 boxes), and the pair of
the full-size C2 frame where the device misses a hit the reference finds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree
from oracle import bind
from helpers import gpu_frame

ctx = Context(0)
scene, _ = bench.make_workload("c2", 4096)
tree = OBBtree(ctx, scene.meshes[0].positions, scene.meshes[0].normals, scene.meshes[0].vertex_ids)
ft = tree.export()
nv = ft.boxes.shape[0]
# subtree triangle ranges by a post-order pass (pre-order array: children have larger indices)
lo = np.zeros(nv, np.int64); hi = np.zeros(nv, np.int64)
for v in range(nv - 1, -1, -1):
    if ft.left[v] < 0:
        lo[v] = ft.tri_off[v]; hi[v] = ft.tri_off[v] + ft.tri_cnt[v]
    else:
        lo[v] = min(lo[ft.left[v]], lo[ft.right[v]]); hi[v] = max(hi[ft.left[v]], hi[ft.right[v]])
P = ft.tri_pos.reshape(-1, 3, 3).astype(np.float64)
worst = 0.0; bad = 0
for v in range(nv):
    b = ft.boxes[v].astype(np.float64); c = b[0:3]; axes = b[3:12].reshape(3, 3)
    pts = P[lo[v]:hi[v]].reshape(-1, 3) - c
    for a in axes:
        l2 = a @ a
        if l2 == 0: continue
        t = np.abs(pts @ a) / l2          # |projection| in units of the half extent: must be <= 1
        m = t.max()
        if m > 1.0: bad += 1
        worst = max(worst, m)
print(f"tree: {nv} vertices, {P.shape[0]} triangles; boxes that fail to contain a vertex along one of their own axes: {bad}; worst ratio {worst:.9f}")
print("contiguous ranges ok:", all(hi[v] - lo[v] == (hi[ft.left[v]] - lo[ft.left[v]]) + (hi[ft.right[v]] - lo[ft.right[v]]) for v in range(nv) if ft.left[v] >= 0))

cd = CollisionDetection(ctx=ctx)
st, bp, ep, hits = gpu_frame(cd, scene, [tree])
ref = bind.RefOracle(); port = bind.PortOracle()
rt = ref.tree_build(scene.meshes[0].positions, scene.meshes[0].normals, scene.meshes[0].vertex_ids)
pt = port.tree_import(ft)
et = [rt] * scene.n_entries
detail, fp = bind.frame_pairs_detail(ref, scene.matrices, et, bp, threads=len(os.sched_getaffinity(0)))
with np.errstate(over="ignore"):
    g_fp = np.zeros(len(bp), np.uint64)
    np.add.at(g_fp, hits["pair"], bind.hit_fingerprint(hits["tri_first"], hits["tri_second"]))
differ = np.nonzero(g_fp != fp)[0]
print("pairs that differ:", len(differ))
def poke(P, Q):
    n = np.cross(P[1] - P[0], P[2] - P[0]); n /= np.linalg.norm(n)
    d = (Q - P[0]) @ n
    return min(d.max(), -d.min())
best = None
pos = scene.meshes[0].positions
for k in differ.tolist():
    i, j = bp[k].tolist()
    r = ref.pair(rt, scene.matrices[i], rt, scene.matrices[j])
    want = set(map(tuple, np.asarray(r.hit_ids).reshape(-1, 2).tolist()))
    got = {(int(h["tri_first"]), int(h["tri_second"])) for h in hits[hits["pair"] == k]}
    rel = ref.pair_matrix(scene.matrices[i], scene.matrices[j]); M = np.asarray(rel, np.float64).reshape(4, 4).T
    for (ta, tb) in want - got:
        A = pos[ta].reshape(3, 3).astype(np.float64); B = pos[tb].reshape(3, 3).astype(np.float64) @ M[:3, :3].T + M[:3, 3]
        d = min(poke(A, B), poke(B, A))
        if best is None or d > best[0]:
            best = (d, k, ta, tb)
d, k, ta, tb = best
i, j = bp[k].tolist()
print("deepest lost hit: depth", d, "pair", k, (i, j), "triangles", ta, tb)
rel = ref.pair_matrix(scene.matrices[i], scene.matrices[j]); M = np.asarray(rel, np.float64).reshape(4, 4).T
flags, seg = ref.tri_tri(pos[ta], pos[tb], rel)
print("reference tri-tri flags", flags, "segment", seg)
leafpos = {int(o): n for n, o in enumerate(ft.tri_orig)}
def chain(t):
    p = leafpos[t]; out = []
    v = 0
    while True:
        out.append(v)
        if ft.left[v] < 0: break
        v = ft.left[v] if lo[ft.left[v]] <= p < hi[ft.left[v]] else ft.right[v]
    return out
ca, cb = chain(ta), chain(tb)
sep = [(va, vb) for va in ca for vb in cb if not port.sat(ft.boxes[va], ft.boxes[vb], rel)]
print("separated ancestor pairs:", sep, "of", len(ca) * len(cb), "leaf pair", ca[-1], cb[-1])
A = pos[ta].reshape(3, 3).astype(np.float64); B = pos[tb].reshape(3, 3).astype(np.float64) @ M[:3, :3].T + M[:3, 3]
for va, vb in sep[:2]:
    ba = ft.boxes[va].astype(np.float64); bb = ft.boxes[vb].astype(np.float64)
    l = [ba[0:3], ba[3:6], ba[6:9], ba[9:12]]
    rb = [M[:3, :3] @ bb[0:3] + M[:3, 3], M[:3, :3] @ bb[3:6], M[:3, :3] @ bb[6:9], M[:3, :3] @ bb[9:12]]
    axes = [np.cross(l[2], l[3]), np.cross(l[1], l[3]), np.cross(l[1], l[2]), np.cross(rb[2], rb[3]), np.cross(rb[1], rb[3]), np.cross(rb[1], rb[2])]
    axes += [np.cross(x, y) for x in l[1:] for y in rb[1:]]
    for q, ax in enumerate(axes):
        n = np.linalg.norm(ax)
        if n == 0: print("  axis", q, "zero"); continue
        ax = ax / n
        pl = l[0] @ ax; rl = sum(abs(v @ ax) for v in l[1:]); pr = rb[0] @ ax; rr = sum(abs(v @ ax) for v in rb[1:])
        gap = abs(pl - pr) - (rl + rr)
        ta_int = (A @ ax).min(), (A @ ax).max(); tb_int = (B @ ax).min(), (B @ ax).max()
        if gap > -1e-4: print(f"  axis {q}: |axis| {n:.3e} gap {gap:+.3e} (negative = overlap)  box A [{pl - rl:+.6f},{pl + rl:+.6f}] tri A [{ta_int[0]:+.6f},{ta_int[1]:+.6f}]  box B [{pr - rr:+.6f},{pr + rr:+.6f}] tri B [{tb_int[0]:+.6f},{tb_int[1]:+.6f}]")
    print("  box A half lengths", [float(np.linalg.norm(v)) for v in l[1:]], "box B", [float(np.linalg.norm(v)) for v in rb[1:]])
    print("  box A axes orthogonality", float(l[1] @ l[2]), float(l[1] @ l[3]), float(l[2] @ l[3]))
sys.exit(0)
