"""CPU: the plain-C restatement (oracle/imr_oracle.c) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle on a box without /root/reference.  Bit-exact."""
import numpy as np
import pytest

import golden_io
from helpers import contacts_close, f32_bits, oracle_frame


def test_sat_surface_transform(port):
    z = golden_io.load("sat")
    a, b, m = z["a"], z["b"], z["mats"]
    n = a.shape[0]
    assert np.array_equal(np.array([port.sat(a[i], b[i], m[i]) for i in range(n)], np.uint8), z["verdict"])
    assert 0.05 < z["verdict"].mean() < 0.95
    assert np.array_equal(f32_bits(np.array([port.surface(a[i]) for i in range(n)], np.float32)), f32_bits(z["surf_a"]))
    assert np.array_equal(f32_bits(np.array([port.surface(b[i], m[i]) for i in range(n)], np.float32)), f32_bits(z["surf_b"]))
    assert np.array_equal(f32_bits(np.stack([port.box_transform(b[i], m[i]) for i in range(n)])), f32_bits(z["xform"]))


def test_tri_tri(port):
    z = golden_io.load("tri_tri")
    f, s = port.tri_tri(z["a"], z["b"], z["m"])
    assert np.array_equal(f, z["flags"]) and (f == 1).sum() > 50
    assert np.array_equal(f32_bits(s), f32_bits(z["seg"]))
    f, s = port.tri_tri(z["da"], z["db"], None)
    assert np.array_equal(f, z["dflags"]) and set(np.unique(f).tolist()) >= {0, 1, 3}
    assert np.array_equal(f32_bits(s), f32_bits(z["dseg"]))


def test_pair_matrix_and_sweep_axes(port):
    z = golden_io.load("pair_matrix")
    rel = np.stack([port.pair_matrix(z["a"][i], z["b"][i]) for i in range(z["a"].shape[0])])
    assert np.array_equal(f32_bits(rel), f32_bits(z["rel"]))
    assert np.array_equal(f32_bits(port.sweep_axes()), f32_bits(z["axes"]))


def test_obb_fit(port):
    z = golden_io.load("obb_fit")
    off = 0
    for k, cnt in enumerate(z["sizes"].tolist()):
        box = port.obb_from_points(z["points"][off:off + cnt]); off += cnt
        assert np.array_equal(f32_bits(box), f32_bits(z["boxes"][k])), f"cloud {k}"


@pytest.mark.parametrize("name", golden_io.tree_names())
def test_tree_build(port, name):
    z = golden_io.load("trees")
    gold, mesh = golden_io.golden_tree(z, name)
    ft = port.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids).flat
    for f in golden_io.TREE_FIELDS:
        g = getattr(gold, f); o = getattr(ft, f)
        if g.dtype == np.float32:
            assert np.array_equal(f32_bits(o), f32_bits(g)), f
        else:
            assert np.array_equal(o, g), f


@pytest.mark.parametrize("name", golden_io.frame_names())
def test_frame(port, name):
    sc, gold = golden_io.golden_frame(golden_io.load("frames"), name)
    trees = [port.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    res = oracle_frame(port, sc, trees, port=port)
    assert np.array_equal(res["pairs"], gold["pairs"])
    for k, key in enumerate(map(tuple, gold["pairs"].tolist())):
        r = res["per_pair"][key]
        summ = [r.n_combos, r.n_tri_tests, r.n_hits, r.n_coplanar, r.rays_first, r.rays_second, int(r.colliding)]
        assert summ == gold["summary"][k].tolist(), f"pair {key}"
        sel = gold["hit_pair"] == k
        # hits come out in traversal order, identical in both implementations
        assert np.array_equal(r.hit_ids, gold["hit_ids"][sel])
        assert np.array_equal(f32_bits(r.hit_seg), f32_bits(gold["hit_seg"][sel]))
        if r.colliding:   # contact averages: unordered_map iteration order leaks into the float sums (SURVEY trap 11)
            rel = port.pair_matrix(sc.matrices[key[0]], sc.matrices[key[1]])
            assert contacts_close(r.avg, gold["avg"][k], rel), f"pair {key}: {r.avg} vs {gold['avg'][k]}"


def test_ray_tree_golden(port):
    """Ray::IntersectOBBtree (Ray.cpp:136-236) against the reference's answers: bit-exact."""
    z = golden_io.load("response")
    from inmyroom_vulkan_b200 import scenes
    mesh = scenes.torus(20, 10)
    tree = port.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids)
    for k in range(z["ray.mats"].shape[0]):
        h, b, dist, bary, tri = port.ray_tree(tree, z["ray.mats"][k], z["ray.origins"][k], z["ray.dirs"][k])
        assert h == bool(z["ray.hit"][k]) and b == bool(z["ray.back"][k]) and (tri & 0xffffffff) == int(z["ray.tri"][k]), k
        assert f32_bits(np.array([dist]))[0] == f32_bits(z["ray.dist"][k:k + 1])[0] and np.array_equal(f32_bits(bary), f32_bits(z["ray.bary"][k])), k
    assert z["ray.hit"].sum() > 300


def test_response_golden(port):
    """ShootUncollideRays on the reference's rays in the reference's order: bit-exact; deltaVectors end to end (the port's ray order
    differs from std::unordered_map's): 1e-4 of the vector's length."""
    z = golden_io.load("response")
    sc, gold = golden_io.golden_frame(golden_io.load("frames"), "torus_instances")
    trees = [port.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    prev = z["frame.previous"]
    o1 = o2 = 0
    n_checked = 0
    for k, (i, j) in enumerate(gold["pairs"].tolist()):
        c1, c2 = z["frame.ray_counts"][k].tolist()
        r1 = z["frame.rays_first"][o1:o1 + c1]; r2 = z["frame.rays_second"][o2:o2 + c2]; o1 += c1; o2 += c2
        ta, tb = trees[sc.mesh_index[i]], trees[sc.mesh_index[j]]
        if c1 or c2:
            d, _ = port.shoot(ta, sc.matrices[i], tb, sc.matrices[j], r1, r2)
            assert np.array_equal(f32_bits(d), f32_bits(z["frame.shoot"][k])), (k, d, z["frame.shoot"][k])
        col, d1, d2 = port.pair_delta(ta, sc.matrices[i], prev[i], tb, sc.matrices[j], prev[j])
        assert int(col) == int(z["frame.colliding"][k])
        g = z["frame.delta"][k].astype(np.float64)
        for a, b in ((d1, g[:3]), (d2, g[3:])):
            if np.isnan(b).any():
                assert np.isnan(a).any()
            else:
                assert np.linalg.norm(a - b) <= 1e-4 * max(np.linalg.norm(b), 1e-30) + 1e-12, (k, a, b)
                n_checked += bool(b.any())
    assert n_checked > 20
