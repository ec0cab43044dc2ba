"""GPU OBB-tree build (IMRCD_BUILD_MORTON): structural validity, conservativeness, and frame parity.

The Morton tree is a different tree from the reference's top-down one, so parity is asserted on
(a) the port oracle traversing the EXPORTED GPU tree  -> must equal the GPU frame bit for bit, and
(b) the reference traversing ITS OWN tree             -> colliding-entity set and triangle-pair set must agree.
"""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import CollisionDetection, OBBtree
from helpers import compare_frame, gpu_frame, oracle_frame

pytestmark = pytest.mark.gpu


def check_tree_structure(flat, mesh):
    nv, n = flat.nv, flat.n_tri
    assert n == mesh.n_tri
    # leaf order is a permutation of the input
    assert np.array_equal(np.sort(flat.tri_orig), np.arange(n, dtype=np.uint32))
    assert np.array_equal(flat.tri_pos, mesh.positions[flat.tri_orig])
    assert np.array_equal(flat.tri_vid, mesh.vertex_ids[flat.tri_orig])
    leaf = flat.left < 0
    assert (flat.tri_cnt[leaf] >= 1).all() and (flat.tri_cnt[leaf] <= 4).all()          # OBBtree.h:49
    assert flat.tri_cnt[leaf].sum() == n
    # leaves in pre-order tile [0, n)
    offs = flat.tri_off[leaf]; cnts = flat.tri_cnt[leaf]
    assert np.array_equal(offs, np.concatenate([[0], np.cumsum(cnts)[:-1]]))
    # triangle range of every vertex (post-order accumulate) and box containment
    lo = np.where(leaf, flat.tri_off, 0).astype(np.int64); hi = np.where(leaf, flat.tri_off + flat.tri_cnt, 0).astype(np.int64)
    for v in range(nv - 1, -1, -1):
        if not leaf[v]:
            l, r = flat.left[v], flat.right[v]
            assert l > v and r > v
            assert hi[l] == lo[r], "children do not tile the parent's range"
            lo[v] = lo[l]; hi[v] = hi[r]
    assert lo[0] == 0 and hi[0] == n
    return lo, hi


def check_boxes_contain(flat, lo, hi, sample=400, seed=0):
    rng = np.random.default_rng(seed)
    vs = np.unique(np.concatenate([[0], rng.integers(0, flat.nv, size=min(sample, flat.nv))]))
    worst = 0.0
    for v in vs:
        b = flat.boxes[v].astype(np.float64)
        c, sides = b[:3], b[3:].reshape(3, 3)
        pts = flat.tri_pos[lo[v]:hi[v]].reshape(-1, 3).astype(np.float64) - c
        for a in range(3):
            h = np.linalg.norm(sides[a])
            ax = sides[a] / h
            # the other two sides are orthogonal to ax up to rounding: the slab |ax.(p-c)| <= h bounds the box
            over = np.abs(pts @ ax).max() - h
            worst = max(worst, over)
    assert worst <= 0.0, f"a box does not contain its triangles (excess {worst})"


@pytest.mark.parametrize("mesh", [scenes.torus(100, 50), scenes.uv_sphere(66, 65), scenes.box_mesh(1, 2, 3, sub=6),
                                  scenes.grid_sheet(40, 30, 1500.0, 900.0, bump=30.0)], ids=lambda m: m.name)
def test_morton_tree_structure_and_containment(gpu_ctx, mesh):
    t = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
    flat = t.export()
    lo, hi = check_tree_structure(flat, mesh)
    check_boxes_contain(flat, lo, hi)


def test_morton_build_two_million_triangles(gpu_ctx):
    """Structure and containment at a size where the tree has thousands of treelets and ~20 levels above them (BASELINE config 4 is the same
    mesh at 10 M triangles: scripts / bench.py --workload c4 check it the same way)."""
    mesh = scenes.grid_sheet(1450, 725, 1500.0, 750.0, bump=40.0)
    assert mesh.n_tri > 2_000_000
    t = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
    flat = t.export()
    lo, hi = check_tree_structure(flat, mesh)
    check_boxes_contain(flat, lo, hi, sample=600)
    # every box on the path from the root to a few leaves, the largest ones included
    leaf_v = np.nonzero(flat.left < 0)[0]
    parent = np.full(flat.nv, -1, np.int64)
    inner = np.nonzero(flat.left >= 0)[0]
    parent[flat.left[inner]] = inner; parent[flat.right[inner]] = inner
    for v in leaf_v[:: max(1, len(leaf_v) // 7)][:7].tolist():
        chain = []
        while v >= 0:
            chain.append(v); v = int(parent[v])
        sub = type("F", (), {})()
        check_boxes_contain_list(flat, lo, hi, chain)


def check_boxes_contain_list(flat, lo, hi, vs):
    for v in vs:
        b = flat.boxes[v].astype(np.float64)
        c, sides = b[:3], b[3:].reshape(3, 3)
        pts = flat.tri_pos[lo[v]:hi[v]].reshape(-1, 3).astype(np.float64) - c
        for a in range(3):
            h = np.linalg.norm(sides[a])
            assert np.abs(pts @ (sides[a] / h)).max() <= h, v


def test_morton_build_tiny_and_empty(gpu_ctx):
    m = scenes.box_mesh(1, 1, 1, sub=1)
    for k in (1, 2, 3, 4, 5):
        sub = scenes.Mesh(m.positions[:k].copy(), m.normals[:k].copy(), m.vertex_ids[:k].copy(), f"k{k}")
        flat = OBBtree(gpu_ctx, sub.positions, sub.normals, sub.vertex_ids).export()
        lo, hi = check_tree_structure(flat, sub)
        check_boxes_contain(flat, lo, hi)
        if k <= 4:
            assert flat.nv == 1            # zero nodes, leaf root (OBBtree.cpp:346-356)
    # missing normals -> face normals (Triangle.cpp:214-234); missing vertex ids -> iota
    flat = OBBtree(gpu_ctx, m.positions, None, None).export()
    p = flat.tri_pos.reshape(-1, 3, 3)
    fn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]); fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    assert np.allclose(flat.tri_nrm.reshape(-1, 3, 3), fn[:, None, :], atol=1e-6)


def _frame_with_gpu_trees(gpu_ctx, port, oracle, scene, strict_sets=True):
    g_trees = [OBBtree(gpu_ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, scene, g_trees)
    # (a) the port oracle on the exported GPU trees: everything bit-exact
    p_trees = [port.tree_import(t.export()) for t in g_trees]
    pres = oracle_frame(port, scene, p_trees, port=port)
    compare_frame(pres, st, bp, ep, hits, rel_of=lambda k: port.pair_matrix(scene.matrices[k[0]], scene.matrices[k[1]]))
    # (b) the reference on its own trees: tree-independent outputs
    o_trees = [oracle.tree_build(m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    ores = oracle_frame(oracle, scene, o_trees, port=port)
    # Which entity is "first" follows the U-order of the ROOT boxes (SweepAndPrune.cpp:63), and the Morton root box is
    # not the reference's root box, so orientation is tree-dependent: compare unordered pairs.
    def canon(pair, ta, tb):
        a, b = pair
        return (a, b, ta, tb) if a < b else (b, a, tb, ta)
    o_coll = {tuple(sorted(k)) for k, r in ores["per_pair"].items() if r.colliding}
    g_coll = {tuple(sorted((int(p["entry_first"]), int(p["entry_second"])))) for p in ep}
    o_hits = {canon(k, int(a), int(b)) for k, r in ores["per_pair"].items() for a, b in r.hit_ids.tolist()}
    g_hits = {canon(tuple(bp[h["pair"]].tolist()), int(h["tri_first"]), int(h["tri_second"])) for h in hits}
    return dict(st=st, ref=ores["totals"], only_ref=o_hits - g_hits, only_gpu=g_hits - o_hits, coll_equal=(o_coll == g_coll),
                only_ref_coll=o_coll - g_coll, only_gpu_coll=g_coll - o_coll, n_hits=len(o_hits))


def _report(r):
    print("morton vs reference trees:", {k: (v if not isinstance(v, set) else (len(v), sorted(v)[:4])) for k, v in r.items()})


def _assert_tree_independent_parity(r, scene, port):
    """The Morton tree is not the reference's tree.  Two effects make the triangle-pair set tree-dependent at the
    1e-4 level (DESIGN.md "parity levels"): (i) the reference's boxes are padded by only 2*FLT_EPSILON (OBB.cpp:123)
    and can cull a true hit that conservative boxes keep; (ii) the tri-tri predicate's EPSILON acts on unnormalised
    plane distances (Triangle.cpp:899-901) and accepts some pairs whose boxes are disjoint in a tight tree.
    Asserted here: identical colliding-entity set; set difference tiny; every differing pair really satisfies the
    reference predicate (so neither side invents hits)."""
    _report(r)
    assert r["coll_equal"], "colliding-entity set differs from the reference"
    n_diff = len(r["only_ref"]) + len(r["only_gpu"])
    assert n_diff <= max(3, int(3e-4 * r["n_hits"])), f"too many differing triangle pairs: {n_diff} of {r['n_hits']}"
    for (a, b, ta, tb) in list(r["only_ref"]) + list(r["only_gpu"]):
        ma, mb = scene.meshes[scene.mesh_index[a]], scene.meshes[scene.mesh_index[b]]
        ok = False
        for first, second, tf, ts, mf, ms in ((a, b, ta, tb, ma, mb), (b, a, tb, ta, mb, ma)):   # either orientation
            rel = port.pair_matrix(scene.matrices[first], scene.matrices[second])
            f, _ = port.tri_tri(mf.positions[tf:tf + 1], ms.positions[ts:ts + 1], rel)
            ok = ok or (f[0] & 1)
        assert ok, f"differing pair {(a, b, ta, tb)} does not satisfy the reference predicate in either orientation"


def test_frame_with_morton_trees_torus(gpu_ctx, port, oracle):
    scene = scenes.scene_instances(scenes.torus(100, 50), 256, seed=1234)
    r = _frame_with_gpu_trees(gpu_ctx, port, oracle, scene)
    _assert_tree_independent_parity(r, scene, port)


def test_frame_with_morton_trees_static_scene(gpu_ctx, port, oracle):
    static = scenes.atrium_static(detail=1)
    keep = list(range(0, 8)) + list(range(60, 70)) + list(range(120, 130))
    static = ([static[0][i] for i in keep], static[1][keep])
    scene = scenes.scene_static_vs_bodies(scenes.uv_sphere(24, 17), 300, seed=5, body_scale=(0.5, 1.5), static=static)
    r = _frame_with_gpu_trees(gpu_ctx, port, oracle, scene)
    _assert_tree_independent_parity(r, scene, port)
