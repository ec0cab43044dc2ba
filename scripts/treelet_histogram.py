#!/usr/bin/env python
"""Sizes of the fit's treelets (maximal subtrees of <= 128 triangles and <= 128 records, csrc/imrcd_fit.cuh) of a GPU-built tree:
python scripts/treelet_histogram.py [grid side]   (a displaced grid like bench.py's C4 / a uv-sphere like C5's characters)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import Context, OBBtree

def treelets(t):
    nv = t.nv
    tri = np.zeros(nv, np.int64); rec = np.ones(nv, np.int64)
    leaf = t.left < 0
    # pre-order flat form: children after parents -> accumulate from the back
    tri[:] = np.where(leaf, t.tri_cnt, 0)
    parent = np.full(nv, -1, np.int64)
    inner = np.nonzero(~leaf)[0]
    parent[t.left[inner]] = inner; parent[t.right[inner]] = inner
    for v in range(nv - 1, 0, -1):
        p = parent[v]
        if p >= 0: tri[p] += tri[v]; rec[p] += rec[v]
    small = (tri <= 128) & (rec <= 128)
    root = small.copy()
    root[1:] &= ~small[parent[1:]] | (parent[1:] < 0)
    return tri[root], rec[root], tri[0]

ctx = Context(0)
for name, mesh in (("a C5 character", scenes.character().mesh), ("displaced grid 708 x 354 (0.5 M triangles, C4's shape)", scenes.grid_sheet(708, 354, 60.0, 30.0, bump=0.5)), ("torus 100 x 50", scenes.torus(100, 50))):
    tree = OBBtree(ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
    t = tree.export()
    ts, rs, total = treelets(t)
    print(name, "triangles", int(total), "treelets", len(ts), "mean size", round(float(ts.mean()), 1))
    h, _ = np.histogram(ts, bins=[1, 2, 5, 9, 17, 33, 65, 97, 129])
    print("   treelets by triangles [1,2) [2,5) [5,9) [9,17) [17,33) [33,65) [65,97) [97,129):", h.tolist())
    print("   triangles in them:", [int(ts[(ts >= a) & (ts < b)].sum()) for a, b in ((1, 2), (2, 5), (5, 9), (9, 17), (17, 33), (33, 65), (65, 97), (97, 129))])
