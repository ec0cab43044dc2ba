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
