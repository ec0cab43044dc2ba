"""GPU parity against the COMMITTED golden vectors (outputs of the unmodified reference, tests/golden/make_golden.py):
the device predicates and whole frames through the C ABI.  Bit-exact."""
import numpy as np
import pytest

import golden_io
from inmyroom_vulkan_b200.collision import CollisionDetection, OBBtree
from helpers import contacts_close, f32_bits, gpu_frame

pytestmark = pytest.mark.gpu


def test_sat_golden(gpu_ctx):
    z = golden_io.load("sat")
    v, sa, sb = gpu_ctx.test_sat(z["a"], z["b"], z["mats"])
    assert np.array_equal(v, z["verdict"])
    assert np.array_equal(f32_bits(sa), f32_bits(z["surf_a"])) and np.array_equal(f32_bits(sb), f32_bits(z["surf_b"]))


def test_tri_tri_golden(gpu_ctx):
    z = golden_io.load("tri_tri")
    f, s = gpu_ctx.test_tri_tri(z["a"], z["b"], z["m"])
    assert np.array_equal(f, z["flags"]) and np.array_equal(f32_bits(s), f32_bits(z["seg"]))
    f, s = gpu_ctx.test_tri_tri(z["da"], z["db"], None)
    assert np.array_equal(f, z["dflags"]) and np.array_equal(f32_bits(s), f32_bits(z["dseg"]))


def test_pair_matrix_golden(gpu_ctx):
    z = golden_io.load("pair_matrix")
    assert np.array_equal(f32_bits(gpu_ctx.test_pair_matrix(z["a"], z["b"])), f32_bits(z["rel"]))


@pytest.mark.parametrize("name", golden_io.frame_names())
def test_frame_golden(gpu_ctx, port, name):
    """Reference-identical trees (the port's build is pinned bit-exact to the reference's by test_oracle_golden.py)
    imported into the GPU; the frame must reproduce the reference's pairs, hits and segments."""
    sc, gold = golden_io.golden_frame(golden_io.load("frames"), name)
    trees = [OBBtree.from_flat(gpu_ctx, port.tree_build(m.positions, m.normals, m.vertex_ids).flat) for m in sc.meshes]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, sc, trees)
    assert set(map(tuple, bp.tolist())) == set(map(tuple, gold["pairs"].tolist()))
    summ = gold["summary"]
    assert st["n_combos"] == summ[:, 0].sum() and st["n_tri_tests"] == summ[:, 1].sum()
    assert st["n_hits"] == summ[:, 2].sum() and st["n_coplanar_hits"] == summ[:, 3].sum() and st["n_colliding"] == summ[:, 6].sum()
    gp = gold["pairs"][gold["hit_pair"]]
    g = {(int(a), int(b), int(i), int(j)): f32_bits(seg).tobytes() for (a, b), (i, j), seg in zip(gp.tolist(), gold["hit_ids"].tolist(), gold["hit_seg"])}
    o = {(int(bp[h["pair"]][0]), int(bp[h["pair"]][1]), int(h["tri_first"]), int(h["tri_second"])):
         f32_bits(np.concatenate([h["source"], h["target"], [h["weight"]]])).tobytes() for h in hits}
    assert o == g
    coll = {tuple(p) for p, s in zip(gold["pairs"].tolist(), summ) if s[6]}
    assert {(int(p["entry_first"]), int(p["entry_second"])) for p in ep} == coll
    # contact reduction against the reference's own numbers: ray counts exact, contact points within 1e-5 relative
    index = {tuple(p): k for k, p in enumerate(gold["pairs"].tolist())}
    for p in ep:
        k = index[(int(p["entry_first"]), int(p["entry_second"]))]
        assert (int(p["n_rays_first"]), int(p["n_rays_second"])) == (int(summ[k][4]), int(summ[k][5]))
        if summ[k][4] and summ[k][5]:
            rel = port.pair_matrix(sc.matrices[int(p["entry_first"])], sc.matrices[int(p["entry_second"])])
            assert contacts_close(np.concatenate([p["avg_first"], p["avg_second"]]), gold["avg"][k], rel)


def test_ray_tree_golden(gpu_ctx):
    """Ray::IntersectOBBtree on the device against the reference's answers (tests/golden/response.npz), with the reference's tree: bit-exact."""
    z = golden_io.load("response")
    gold_tree, mesh = golden_io.golden_tree(golden_io.load("trees"), "torus20x10")
    gt = OBBtree.from_flat(gpu_ctx, gold_tree)
    hit, back, dist, bary, tri = gpu_ctx.test_ray_tree(gt, z["ray.mats"], z["ray.origins"], z["ray.dirs"])
    assert np.array_equal(hit, z["ray.hit"].astype(bool))
    sel = hit
    assert np.array_equal(back[sel], z["ray.back"].astype(bool)[sel]) and np.array_equal(tri[sel], z["ray.tri"][sel])
    assert np.array_equal(f32_bits(dist[sel]), f32_bits(z["ray.dist"][sel])) and np.array_equal(f32_bits(bary[sel]), f32_bits(z["ray.bary"][sel]))


def test_delta_golden(gpu_ctx):
    """deltaVector of every colliding pair of the golden frame against the reference's (1e-4 of the vector's length: the rays of a pair
    are summed in a different order)."""
    z = golden_io.load("response")
    sc, gold = golden_io.golden_frame(golden_io.load("frames"), "torus_instances")
    sc.previous = z["frame.previous"]
    tz = golden_io.load("trees")
    gold_tree, _ = golden_io.golden_tree(tz, "torus20x10")
    trees = [OBBtree.from_flat(gpu_ctx, gold_tree) for _ in sc.meshes]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, sc, trees)
    want = {tuple(p): (z["frame.delta"][k], int(z["frame.colliding"][k])) for k, p in enumerate(gold["pairs"].tolist())}
    assert len(ep) == sum(c for _, c in want.values())
    n_checked = 0
    for p in ep:
        g, col = want[(int(p["entry_first"]), int(p["entry_second"]))]
        assert col
        for a, b in ((p["delta_first"], g[:3]), (p["delta_second"], g[3:])):
            a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
            if np.isnan(b).any():
                assert np.isnan(a).any()
            else:
                assert np.linalg.norm(a - b) <= 1e-4 * max(np.linalg.norm(b), 1e-30) + 1e-12, (a, b)
                n_checked += bool(b.any())
    assert n_checked > 20
