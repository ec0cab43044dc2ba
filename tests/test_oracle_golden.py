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
