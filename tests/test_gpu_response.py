"""GPU parity of the response stage: ray-vs-OBB-tree queries bit for bit, and the deltaVector handed to
CollisionCallback (ShootUncollideRays.cpp:14-93, CollisionDetection.cpp:80-103) against the oracle, through the C ABI."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import CollisionDetection, OBBtree
from helpers import f32_bits, gpu_frame

pytestmark = pytest.mark.gpu


def _rays(rng, n):
    o = (rng.normal(size=(n, 3)) * np.where(np.arange(n) % 3 == 0, 3.0, 0.2)[:, None]).astype(np.float32)
    o[::50] = 0                                            # the un-centred branch (Ray.cpp:138)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return o, d


@pytest.mark.parametrize("mesh", [scenes.torus(40, 20), scenes.uv_sphere(24, 17), scenes.box_mesh(1, 1, 1, sub=1)], ids=lambda m: m.name)
def test_ray_tree_bit_exact(gpu_ctx, oracle, mesh):
    """Ray::IntersectOBBtree (Ray.cpp:136-236) with the oracle's own tree imported: flags, distance, barycentrics, triangle."""
    rng = np.random.default_rng(7)
    ot = oracle.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids)
    gt = OBBtree.from_flat(gpu_ctx, ot.flat)
    n = 3000
    s = rng.random((n, 3)) * 1.5 + 0.25
    mats = scenes.trs_matrices(rng.normal(size=(n, 3)) * 0.3, scenes.random_quaternions(rng, n), s)
    o, d = _rays(rng, n)
    hit, back, dist, bary, tri = gpu_ctx.test_ray_tree(gt, mats, o, d)
    n_hit = 0
    for k in range(n):
        h, b, ds, br, t = oracle.ray_tree(ot, mats[k], o[k], d[k])
        assert bool(hit[k]) == h, k
        if h:
            n_hit += 1
            assert bool(back[k]) == b and int(tri[k]) == t, (k, tri[k], t)
            assert f32_bits(np.array([dist[k]]))[0] == f32_bits(np.array([ds]))[0] and np.array_equal(f32_bits(bary[k]), f32_bits(br)), (k, dist[k], ds)
    assert n_hit > 200


def _moved_scene(seed, n, mesh, step=0.02):
    sc = scenes.scene_instances(mesh, n, seed=seed, neighbours=6.0)
    rng = np.random.default_rng(seed + 100)
    prev = sc.matrices.copy()
    prev[:, 12:15] += (rng.normal(size=(n, 3)) * step).astype(np.float32)
    prev[::7] = sc.matrices[::7]                           # some entries did not move
    sc.previous = prev
    return sc


def _check_deltas(oracle, sc, o_trees, ep, rtol):
    n_nonzero = 0
    for p in ep:
        i, j = int(p["entry_first"]), int(p["entry_second"])
        col, d1, d2 = oracle.pair_delta(o_trees[sc.mesh_index[i]], sc.matrices[i], sc.previous[i], o_trees[sc.mesh_index[j]], sc.matrices[j], sc.previous[j])
        assert col
        for g, r in ((p["delta_first"], d1), (p["delta_second"], d2)):
            g = np.asarray(g, np.float64); r = np.asarray(r, np.float64)
            if np.isnan(r).any():                          # 0 rays on one side: NaN contact point -> NaN deltas, like the reference
                assert np.isnan(g).any(), ((i, j), g, r)
                continue
            assert np.linalg.norm(g - r) <= rtol * max(np.linalg.norm(r), 1e-30) + 1e-12, ((i, j), g, r)
        n_nonzero += bool(np.nan_to_num(d1).any() or np.nan_to_num(d2).any())
    return n_nonzero


def test_delta_vectors_imported_trees(gpu_ctx, oracle):
    """deltaVector of every colliding pair.  The response is a float sum / max over the pair's rays, which the reference visits in
    std::unordered_map order (CreateUncollideRays.cpp:138): agreement to 1e-4 of the vector's length, not bit for bit (the port oracle,
    fed the reference's rays in the reference's order, is bit-identical: tests/test_oracle_vs_ref.py)."""
    mesh = scenes.torus(40, 20)
    sc = _moved_scene(23, 120, mesh)
    o_trees = [oracle.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    g_trees = [OBBtree.from_flat(gpu_ctx, t.flat) for t in o_trees]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, sc, g_trees)
    assert st["n_colliding"] > 20 and st["n_rays_shot"] > 500 and st["n_responses"] > 100
    assert _check_deltas(oracle, sc, o_trees, ep, 1e-4) > 10
    # a pair of entries that both stood still gets zero vectors (CollisionDetection.cpp:99-103)
    still = [p for p in ep if int(p["entry_first"]) % 7 == 0 and int(p["entry_second"]) % 7 == 0]
    for p in still:
        assert not np.asarray(p["delta_first"]).any() and not np.asarray(p["delta_second"]).any()


def test_delta_vectors_zero_without_previous(gpu_ctx):
    """previous == current for every entry: the response stage is skipped and every deltaVector is (0,0,0)."""
    mesh = scenes.torus(40, 20)
    sc = scenes.scene_instances(mesh, 100, seed=3, neighbours=6.0)
    tree = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, sc, [tree])
    assert st["n_colliding"] > 5 and st["n_rays_shot"] == 0
    assert not ep["delta_first"].any() and not ep["delta_second"].any()


def test_delta_vectors_large_pairs(gpu_ctx, oracle):
    """Pairs with thousands of hits: the size class of the contact reduction whose tables live in global scratch."""
    mesh = scenes.torus(200, 100)
    rng = np.random.default_rng(10)                        # a seed without coplanar triangle pairs: the reference reads uninitialised
    n = 4                                                  # memory on some of those (Triangle.cpp:956-960, see imrcd_math.cuh tt_segment)
    q = scenes.random_quaternions(rng, n)
    mats = scenes.trs_matrices(rng.normal(size=(n, 3)) * 0.15, q, np.ones((n, 3)) * (1.0 + 0.07 * np.arange(n))[:, None])
    prev = mats.copy(); prev[:, 12:15] += (rng.normal(size=(n, 3)) * 0.02).astype(np.float32)
    sc = scenes.Scene([mesh], np.zeros(n, np.uint32), np.ascontiguousarray(mats, np.float32), np.ones(n, np.uint8), np.arange(1, n + 1, dtype=np.uint32))
    sc.previous = np.ascontiguousarray(prev, np.float32)
    o_trees = [oracle.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids)]
    g_trees = [OBBtree.from_flat(gpu_ctx, o_trees[0].flat)]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, sc, g_trees)
    assert st["n_colliding"] == 6 and int(ep["n_hits"].min()) > 1024 and st["n_coplanar_hits"] == 0
    for p in ep:                                           # ray counts of the large class against the oracle
        i, j = int(p["entry_first"]), int(p["entry_second"])
        r = oracle.pair(o_trees[0], sc.matrices[i], o_trees[0], sc.matrices[j])
        assert (int(p["n_rays_first"]), int(p["n_rays_second"]), int(p["n_hits"])) == (r.rays_first, r.rays_second, r.n_hits)
    _check_deltas(oracle, sc, o_trees, ep[:6], 1e-4)
