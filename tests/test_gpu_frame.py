"""GPU parity, frame level: broad + mid + narrow through the C ABI against the oracle, with the oracle's
own trees imported (isolates traversal / predicates from the GPU build)."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import CollisionDetection, OBBtree
from helpers import compare_frame, gpu_frame, oracle_frame, same_entity_pairs

pytestmark = pytest.mark.gpu


def _run(gpu_ctx, oracle, scene):
    o_trees = [oracle.tree_build(m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    g_trees = [OBBtree.from_flat(gpu_ctx, t.flat) for t in o_trees]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, scene, g_trees)
    ores = oracle_frame(oracle, scene, o_trees)
    compare_frame(ores, st, bp, ep, hits, rel_of=lambda k: oracle.pair_matrix(scene.matrices[k[0]], scene.matrices[k[1]]))
    return st, ores


def test_frame_torus_instances_imported_trees(gpu_ctx, oracle):
    scene = scenes.scene_instances(scenes.torus(100, 50), 256, seed=1234)
    st, ores = _run(gpu_ctx, oracle, scene)
    assert st["n_hits"] > 1000 and st["n_colliding"] > 10


def test_frame_nonuniform_scale_static_vs_bodies(gpu_ctx, oracle):
    """Sponza-style non-uniform node scale makes the transformed boxes parallelepipeds (SURVEY trap 2)."""
    static = scenes.atrium_static(detail=1)
    keep = list(range(0, 8)) + list(range(60, 70)) + list(range(120, 130))
    static = ([static[0][i] for i in keep], static[1][keep])
    scene = scenes.scene_static_vs_bodies(scenes.uv_sphere(24, 17), 300, seed=5, body_scale=(0.5, 1.5), static=static)
    st, ores = _run(gpu_ctx, oracle, scene)
    assert st["n_pairs"] > 50 and st["n_hits"] > 0


def test_frame_small_and_degenerate_meshes(gpu_ctx, oracle):
    """<= 4-triangle meshes have zero nodes and a leaf root (OBBtree.cpp:346-356); single entries are a no-op."""
    tiny = scenes.box_mesh(1, 1, 1, sub=1)
    tiny4 = scenes.Mesh(tiny.positions[:4].copy(), tiny.normals[:4].copy(), tiny.vertex_ids[:4].copy(), "tiny4")
    one = scenes.Mesh(tiny.positions[:1].copy(), tiny.normals[:1].copy(), tiny.vertex_ids[:1].copy(), "one")
    rng = np.random.default_rng(3)
    n = 60
    meshes = [tiny4, one, scenes.box_mesh(1, 1, 1, sub=2)]
    mats = scenes.trs_matrices(rng.normal(size=(n, 3)) * 1.5, scenes.random_quaternions(rng, n), np.ones((n, 3)))
    scene = scenes.Scene(meshes, (np.arange(n) % 3).astype(np.uint32), mats, np.ones(n, np.uint8), np.arange(1, n + 1, dtype=np.uint32))
    st, ores = _run(gpu_ctx, oracle, scene)
    assert st["n_pairs"] > 10


def test_frame_fewer_than_two_entries_is_noop(gpu_ctx):
    cd = CollisionDetection(ctx=gpu_ctx)
    cd.Reset()
    cd.ExecuteCollisionDetection()       # CollisionDetection.cpp:40
    assert cd._n == 0


def test_should_callback_filter(gpu_ctx, oracle):
    """A pair needs shouldCallback on either side (SweepAndPrune.cpp:60)."""
    scene = scenes.scene_instances(scenes.torus(40, 20), 200, seed=9)
    scene.should_callback[::2] = 0
    _run(gpu_ctx, oracle, scene)


def test_zero_copy_submission_matches_add_entries(gpu_ctx):
    """imrcd_frame_map_entries / commit_entries (entries written straight into pinned staging) == add_entries."""
    mesh = scenes.torus(40, 20)
    scene = scenes.scene_instances(mesh, 300, seed=17)
    tree = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
    ids = np.full(scene.n_entries, tree.mesh_id, np.uint32)
    cd = CollisionDetection(ctx=gpu_ctx)

    def snapshot():
        cd.ExecuteCollisionDetection()
        ep, hits = cd.results(True)
        bp = cd.broad_pairs()
        key = np.stack([bp[hits["pair"], 0], bp[hits["pair"], 1], hits["tri_first"], hits["tri_second"]], 1)
        order = np.lexsort(key.T[::-1])
        eo = np.lexsort((ep["entry_second"], ep["entry_first"]))
        return cd.stats(), key[order], hits["source"][order], ep[eo]

    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities)
    a = snapshot()
    cd.Reset()
    half = scene.n_entries // 2
    for lo, hi in ((0, half), (half, scene.n_entries)):          # two map/commit rounds append
        v = cd.map_entries(hi - lo)
        v.current[:] = scene.matrices[lo:hi]; v.mesh_ids[:] = ids[lo:hi]; v.should_callback[:] = scene.should_callback[lo:hi]; v.entities[:] = scene.entities[lo:hi]
        cd.commit_entries(hi - lo)
    b = snapshot()
    for k in ("n_pairs", "n_sat_tests", "n_combos", "n_tri_tests", "n_hits", "n_colliding", "n_rays"):
        assert a[0][k] == b[0][k], k
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    same_entity_pairs(a[3], b[3])                                # colliding pairs, ray counts exact; contact points to FP32 rounding


def test_shards_partition_the_frame(gpu_ctx):
    """imrcd_frame_set_shard(r, n): the n slices are disjoint and add up to the unsharded frame (pairs, hits, colliding pairs)."""
    static = scenes.atrium_static(detail=1)
    keep = list(range(0, 8)) + list(range(60, 70))
    static = ([static[0][i] for i in keep], static[1][keep])
    scene = scenes.scene_static_vs_bodies(scenes.uv_sphere(24, 17), 2000, seed=11, body_scale=(0.5, 1.5), static=static)
    trees = [OBBtree(gpu_ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]

    def run(rank, world):
        cd = CollisionDetection(ctx=gpu_ctx)
        cd.set_shard(rank, world)
        st, bp, ep, hits = gpu_frame(cd, scene, trees)
        cd.set_shard(0, 1)
        pairs = set(map(tuple, bp.tolist()))
        assert len(pairs) == len(bp)
        hk = {(int(bp[h["pair"]][0]), int(bp[h["pair"]][1]), int(h["tri_first"]), int(h["tri_second"])) for h in hits}
        coll = {(int(p["entry_first"]), int(p["entry_second"])): (int(p["n_rays_first"]), int(p["n_rays_second"]), tuple(p["avg_first"].tolist()), tuple(p["avg_second"].tolist())) for p in ep}
        return pairs, hk, coll

    full = run(0, 1)
    assert len(full[0]) > 500 and len(full[2]) > 20
    for world in (2, 3, 8):
        parts = [run(r, world) for r in range(world)]
        for k in range(3):
            keys = [set(p[k]) for p in parts]
            assert sum(len(x) for x in keys) == len(set().union(*keys)), "shards overlap"
            assert set().union(*keys) == set(full[k]), "shards do not add up to the frame"
        merged = {}
        for p in parts:
            merged.update(p[2])
        assert merged.keys() == full[2].keys()        # per-pair results do not depend on the shard that computed them
        for key, (ra, rb, pa, pb) in merged.items():
            fa, fb, qa, qb = full[2][key]
            assert (ra, rb) == (fa, fb)
            for x, y in ((pa, qa), (pb, qb)):         # contact points: equal to FP32 rounding (order-free FP64 sums, see helpers.same_entity_pairs)
                x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
                assert (np.isnan(x).any() and np.isnan(y).any()) or np.linalg.norm(x - y) <= 1e-6 * max(np.linalg.norm(y), 1e-30)
        sizes = [len(p[0]) for p in parts]
        assert max(sizes) <= 1.5 * (sum(sizes) / world) + 32


def test_zero_weight_hits_follow_the_combo_rule(gpu_ctx, oracle):
    """A triangle pair that touches in one point is an intersecting, non-coplanar hit whose segment has length 0: its candidate is dropped
    unless the same combo gives the triangle another hit with weight (TriangleCandidateRays::IsNull, CreateUncollideRays.cpp:22-25,117-127).
    Exact small-integer coordinates, identity matrices: (a) the touching pair alone: one hit, no ray, not colliding; (b) with a second
    triangle of the same leaf cutting through: both hits count for first's triangle, only the cutting one for second's."""
    A = np.array([[0, 0, 0, 4, 0, 0, 0, 4, 0], [0, 0, -5, 4, 0, -5, 0, 4, -5]], np.float32)            # second triangle far below: never hit
    touch = np.array([[1, 1, 0, 1, 1, 2, 2, 1, 2]], np.float32)                                       # vertex (1,1,0) lies inside A0
    cut = np.array([[2, 1, -1, 2, 1, 1, 2, 2, 1]], np.float32)                                        # crosses z = 0 inside A0
    ident = np.eye(4, dtype=np.float32).reshape(1, 16)
    for name, B, want_hits in (("touch", touch, 1), ("touch+cut", np.concatenate([touch, cut]), 2)):
        meshes = [scenes.Mesh(m, np.tile(np.array([0, 0, 1], np.float32), (m.shape[0], 3)), np.arange(3 * m.shape[0], dtype=np.uint32).reshape(-1, 3), nm)
                  for m, nm in ((A, "A"), (B, "B"))]
        scene = scenes.Scene(meshes, np.array([0, 1], np.uint32), np.concatenate([ident, ident]), np.ones(2, np.uint8), np.array([1, 2], np.uint32))
        st, ores = _run(gpu_ctx, oracle, scene)
        assert st["n_hits"] == want_hits and st["n_coplanar_hits"] == 0, name
        r = next(iter(ores["per_pair"].values()))
        assert r.n_hits == want_hits and (r.colliding, st["n_colliding"]) == ((want_hits == 2), int(want_hits == 2)), name
        if want_hits == 1:
            assert float(r.hit_seg[0, 6]) == 0.0 and (r.rays_first, r.rays_second) == (0, 0)


def test_run_async_finish_equals_run(gpu_ctx):
    """imrcd_frame_run_async + imrcd_frame_finish is imrcd_frame_run in two halves: same statistics, same records; a fresh context whose
    first frame overflows its initial buffers reports the re-run through finish()."""
    from inmyroom_vulkan_b200.collision import Context
    mesh = scenes.torus(40, 20)
    scene = scenes.scene_instances(mesh, 300, seed=17)
    tree = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
    ids = np.full(scene.n_entries, tree.mesh_id, np.uint32)
    cd = CollisionDetection(ctx=gpu_ctx)
    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities); cd.upload(); cd.run(); cd.fetch()
    a_st = cd.stats(); a_ep, _ = cd.results(want_hits=False)
    cd.Reset(); cd.add_entries(scene.matrices, ids, scene.should_callback, scene.entities); cd.upload(); cd.run_async()
    reran = cd.finish(); cd.fetch()
    b_st = cd.stats(); b_ep, _ = cd.results(want_hits=False)
    assert not reran
    for k in ("n_pairs", "n_sat_tests", "n_combos", "n_tri_tests", "n_hits", "n_colliding", "n_rays"):
        assert a_st[k] == b_st[k], k
    key = lambda e: np.lexsort((e["entry_second"], e["entry_first"]))
    same_entity_pairs(a_ep[key(a_ep)], b_ep[key(b_ep)])
    # a scene with more than 2^20 hits on a fresh context: the first frame overflows the initial hit buffer
    ctx2 = Context(0)
    big = scenes.torus(100, 50)
    sc2 = scenes.scene_instances(big, 3000, seed=3, neighbours=8.0)
    t2 = OBBtree(ctx2, big.positions, big.normals, big.vertex_ids)
    cd2 = CollisionDetection(ctx=ctx2)
    cd2.Reset(); cd2.add_entries(sc2.matrices, np.full(sc2.n_entries, t2.mesh_id, np.uint32), sc2.should_callback, sc2.entities); cd2.upload(); cd2.run_async()
    reran = cd2.finish()
    st = cd2.stats()
    assert st["n_hits"] > (1 << 20) and reran
    cd2.run_async(); assert not cd2.finish() and cd2.stats()["n_hits"] == st["n_hits"]
    ctx2.close()


def test_broad_phase_paths_agree(gpu_ctx, port, monkeypatch):
    """The sort-free broad phase (frames with few flagged entries: every entry against the dense list of the flagged ones) and the
    sort-and-sweep one give the same ordered pair set, unsharded and sharded, and both equal the port of SweepAndPrune.cpp:15-88."""
    from inmyroom_vulkan_b200.collision import Context
    monkeypatch.setenv("IMRCD_FEW_FLAGGED_MAX", "0")
    swept = Context(0)                                          # reads the knob at its first frame
    static = scenes.atrium_static(detail=1)
    keep = list(range(0, 8)) + list(range(60, 70))
    static = ([static[0][i] for i in keep], static[1][keep])
    sc_a = scenes.scene_static_vs_bodies(scenes.uv_sphere(12, 9), 3000, seed=21, body_scale=(0.5, 1.5), static=static)
    sc_b = scenes.scene_instances(scenes.torus(16, 8), 700, seed=22, neighbours=8.0)
    sc_c = scenes.scene_instances(scenes.torus(16, 8), 500, seed=23, neighbours=8.0)
    sc_c.should_callback[::3] = 0                               # a mix: two thirds flagged
    for sc in (sc_a, sc_b, sc_c):
        got = {}
        for name, ctx in (("few", gpu_ctx), ("sweep", swept)):
            trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
            cd = CollisionDetection(ctx=ctx)
            _, bp, _, _ = gpu_frame(cd, sc, trees)
            got[name] = set(map(tuple, bp.tolist()))
            assert len(got[name]) == len(bp)
            for world in (2, 5):
                parts = []
                for r in range(world):
                    cd.set_shard(r, world)
                    parts.append(set(map(tuple, gpu_frame(cd, sc, trees)[1].tolist())))
                cd.set_shard(0, 1)
                assert sum(len(x) for x in parts) == len(set().union(*parts)) and set().union(*parts) == got[name], (name, world)
        assert got["few"] == got["sweep"] and len(got["few"]) > 100
        ptrees = [port.tree_import(OBBtree(gpu_ctx, m.positions, m.normals, m.vertex_ids).export()) for m in sc.meshes]
        want, _ = port.broad(sc.matrices, [ptrees[m] for m in sc.mesh_index], sc.should_callback)
        assert got["few"] == set(map(tuple, want.tolist()))
    swept.close()


def test_frame_with_coplanar_and_zero_area_triangles(gpu_ctx, port):
    """Frame-level coverage of the path DESIGN section 8 documents: coincident faces (intersecting AND coplanar pairs, dropped by
    CreateUncollideRays.cpp:88) and zero-area triangles (their normal is 0, everything is "coplanar" to them, and
    tri_tri_intersect_with_isectline reads its segment uninitialised, Triangle.cpp:956-960).  There the reference's answer depends on
    stack contents; the library and the port define the unread values as zeros.  So the checker of this scene is the port, on the same
    (reference-identical) trees: counters, hit sets and segments bit for bit, n_coplanar_hits > 0."""
    box = scenes.box_mesh(1.0, 1.0, 1.0, sub=2)
    # a mesh with zero-area triangles: every fourth triangle of the box collapsed onto an edge (two equal corners)
    deg = scenes.Mesh(box.positions.copy(), box.normals.copy(), box.vertex_ids.copy(), "degenerate")
    deg.positions[::4, 6:9] = deg.positions[::4, 3:6]
    meshes = [box, deg]
    t = np.array([[0, 0, 0], [0.5, 0.25, 0.0], [2.0, 0, 0], [2.0, 0.5, 0.0], [0.25, 0.25, 2.0], [4.0, 4.0, 4.0], [4.5, 4.0, 4.25], [0.0, 0.0, 2.0]], np.float64)
    n = len(t)
    q = np.tile([0.0, 0.0, 0.0, 1.0], (n, 1))              # axis-aligned: faces of neighbouring boxes lie exactly in common planes
    mats = scenes.trs_matrices(t, q, np.ones((n, 3)))
    scene = scenes.Scene(meshes, np.array([0, 0, 0, 1, 0, 1, 1, 1], np.uint32), mats, np.ones(n, np.uint8), np.arange(1, n + 1, dtype=np.uint32))
    p_trees = [port.tree_build(m.positions, m.normals, m.vertex_ids) for m in meshes]
    g_trees = [OBBtree.from_flat(gpu_ctx, tr.flat) for tr in p_trees]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, scene, g_trees)
    ores = oracle_frame(port, scene, p_trees, port=port)
    assert ores["totals"]["coplanar"] > 20 and ores["totals"]["hits"] > 20
    compare_frame(ores, st, bp, ep, hits, rel_of=lambda k: port.pair_matrix(scene.matrices[k[0]], scene.matrices[k[1]]))
    assert st["n_coplanar_hits"] == ores["totals"]["coplanar"] > 0
