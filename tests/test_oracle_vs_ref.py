"""CPU: the plain-C restatement against the UNMODIFIED reference compiled in place (oracle/_ref), on fresh seeded
inputs larger than the committed fixtures.  Skipped where neither /root/reference nor a prebuilt _ref exists."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from helpers import contacts_close, f32_bits, oracle_frame


def _rand_mats(rng, n, tscale=1.0):
    s = rng.random((n, 3)) * 1.5 + 0.25
    return scenes.trs_matrices(rng.normal(size=(n, 3)) * tscale, scenes.random_quaternions(rng, n), s)


def test_eig3_matches_lapack_up_to_sign(port):
    """eig3 (JAMA tred2/tql2, eig3.cpp:256-265): eigenvalues ascending, eigenvectors in V's COLUMNS."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        B = rng.normal(size=(3, 3)); A = B @ B.T
        V, d = port.eig3(A)
        w, U = np.linalg.eigh(A)
        assert np.allclose(d, w, rtol=1e-10, atol=1e-12)
        for k in range(3):
            assert np.allclose(A @ V[:, k], d[k] * V[:, k], atol=1e-9)


def test_predicates_random(port, ref):
    rng = np.random.default_rng(2)
    n = 3000
    a = rng.normal(size=(n, 12)).astype(np.float32); b = rng.normal(size=(n, 12)).astype(np.float32); b[:, :3] *= 2.0
    m = _rand_mats(rng, n)
    for i in range(n):
        assert port.sat(a[i], b[i], m[i]) == ref.sat(a[i], b[i], m[i])
    ta = rng.normal(size=(50000, 9)).astype(np.float32); tb = (rng.normal(size=(50000, 9)) * 0.7).astype(np.float32)
    fp, sp = port.tri_tri(ta, tb, m[0]); fr, sr = ref.tri_tri(ta, tb, m[0])
    assert np.array_equal(fp, fr) and np.array_equal(f32_bits(sp), f32_bits(sr))


def test_obb_fit_random(port, ref):
    rng = np.random.default_rng(3)
    for k in range(300):
        cnt = int(rng.integers(1, 200))
        p = (rng.normal(size=(cnt, 3)) * (rng.random(3) * 100 + 0.01) + rng.normal(size=3) * 1000).astype(np.float32)
        assert np.array_equal(f32_bits(port.obb_from_points(p)), f32_bits(ref.obb_from_points(p))), k


@pytest.mark.parametrize("mesh", [scenes.torus(60, 30), scenes.uv_sphere(24, 17), scenes.grid_sheet(30, 20, 1500.0, 900.0, bump=30.0),
                                  scenes.box_mesh(1, 2, 3, sub=5)], ids=lambda m: m.name)
def test_tree_build_identical(port, ref, mesh):
    a = port.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids).flat
    b = ref.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids).flat
    for f in ("boxes", "tri_pos", "tri_nrm"):
        assert np.array_equal(f32_bits(getattr(a, f)), f32_bits(getattr(b, f))), f
    for f in ("left", "right", "tri_off", "tri_cnt", "tri_vid", "tri_orig"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_frame_identical(port, ref):
    sc = scenes.scene_instances(scenes.torus(40, 20), 120, seed=21, neighbours=6.0)
    sc.should_callback[::3] = 0
    tp = [port.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    tr = [ref.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    rp = oracle_frame(port, sc, tp, port=port); rr = oracle_frame(ref, sc, tr, port=port)
    assert np.array_equal(rp["pairs"], rr["pairs"]) and len(rp["pairs"]) > 100
    assert rp["totals"] == rr["totals"] and rr["totals"]["hits"] > 500
    for key, r in rr["per_pair"].items():
        p = rp["per_pair"][key]
        assert np.array_equal(p.hit_ids, r.hit_ids) and np.array_equal(f32_bits(p.hit_seg), f32_bits(r.hit_seg))
        assert (p.rays_first, p.rays_second, p.colliding) == (r.rays_first, r.rays_second, r.colliding)
        if r.colliding:
            assert contacts_close(p.avg, r.avg, port.pair_matrix(sc.matrices[key[0]], sc.matrices[key[1]]))


def test_batch_frame_pairs_matches_per_pair(port, ref):
    """bench.py's CPU legs use the batch entry points; they must agree with the per-pair path and with each other."""
    from oracle import bind
    sc = scenes.scene_instances(scenes.torus(30, 14), 80, seed=5, neighbours=6.0)
    for orc in (port, ref):
        trees = [orc.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
        et = [trees[m] for m in sc.mesh_index]
        res = oracle_frame(orc, sc, trees, port=port)
        for th in (1, 3):
            r = bind.frame_pairs(orc, sc.matrices, et, res["pairs"], threads=th)
            assert (r["combos"], r["tri_tests"], r["colliding"]) == (res["totals"]["combos"], res["totals"]["tri_tests"], res["totals"]["colliding"])


def test_ray_tree_identical(port, ref):
    """Ray::IntersectOBBtree (Ray.cpp:136-236): same hit flag, back-face flag, distance, barycentrics and triangle, bit for bit."""
    rng = np.random.default_rng(31)
    mesh = scenes.torus(40, 20)
    tp = port.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids); tr = ref.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids)
    mats = _rand_mats(rng, 40, tscale=0.3)
    n_hit = 0
    for k in range(2000):
        m = mats[k % len(mats)]
        o = (rng.normal(size=3) * (0.2 if k % 3 else 3.0)).astype(np.float32)
        d = rng.normal(size=3); d = (d / np.linalg.norm(d)).astype(np.float32)
        if k % 50 == 0:
            o[:] = 0                                   # the un-centred branch (Ray.cpp:138)
        a = port.ray_tree(tp, m, o, d); b = ref.ray_tree(tr, m, o, d)
        assert a[0] == b[0] and a[1] == b[1] and a[4] == b[4], (k, a, b)
        assert np.array_equal(f32_bits(np.array([a[2]])), f32_bits(np.array([b[2]]))) and np.array_equal(f32_bits(a[3]), f32_bits(b[3])), (k, a, b)
        n_hit += a[0]
    assert n_hit > 300


def test_pair_delta_matches_reference(port, ref):
    """deltaVector of both entities (ShootUncollideRays.cpp:14-93, CollisionDetection.cpp:80-103).  The rays of a pair come out of
    std::unordered_map iteration in the reference and in merge order in the port, and the response is a float sum / max over them:
    so the two agree to 1e-4 of the vector's length end to end; on the SAME rays in the SAME order (the reference's) the port is bit-identical."""
    rng = np.random.default_rng(41)
    sc = scenes.scene_instances(scenes.torus(40, 20), 60, seed=23, neighbours=6.0)
    tp = [port.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    tr = [ref.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    pairs = oracle_frame(port, sc, tp, port=port)["pairs"]
    prev = sc.matrices.copy()
    prev[:, 12:15] += (rng.normal(size=(sc.n_entries, 3)) * 0.02).astype(np.float32)       # every entry moved a little since the last frame
    n_col = n_nonzero = 0
    for (i, j) in pairs.tolist():
        cp, p1, p2 = port.pair_delta(tp[sc.mesh_index[i]], sc.matrices[i], prev[i], tp[sc.mesh_index[j]], sc.matrices[j], prev[j])
        cr, r1, r2 = ref.pair_delta(tr[sc.mesh_index[i]], sc.matrices[i], prev[i], tr[sc.mesh_index[j]], sc.matrices[j], prev[j])
        assert cp == cr
        if not cr:
            assert not p1.any() and not p2.any() and not r1.any() and not r2.any()
            continue
        n_col += 1
        for a, b in ((p1, r1), (p2, r2)):
            assert np.linalg.norm(a.astype(np.float64) - b) <= 1e-4 * max(np.linalg.norm(b), 1e-30) + 1e-12, ((i, j), a, b)
        ta, tb = tr[sc.mesh_index[i]], tr[sc.mesh_index[j]]
        r1, r2 = ref.pair_rays(ta, sc.matrices[i], tb, sc.matrices[j])
        d_ref, _ = ref.shoot(ta, sc.matrices[i], tb, sc.matrices[j], r1, r2)
        d_port, _ = port.shoot(tp[sc.mesh_index[i]], sc.matrices[i], tp[sc.mesh_index[j]], sc.matrices[j], r1, r2)
        assert np.array_equal(f32_bits(d_ref), f32_bits(d_port)), ((i, j), d_ref, d_port)
        n_nonzero += bool(r1.any() or r2.any())
        # unmoved entries: the response is skipped (CollisionDetection.cpp:99-103)
        c0, z1, z2 = port.pair_delta(tp[sc.mesh_index[i]], sc.matrices[i], sc.matrices[i], tp[sc.mesh_index[j]], sc.matrices[j], sc.matrices[j])
        assert c0 and not z1.any() and not z2.any()
    assert n_col > 10 and n_nonzero > 5
