"""BASELINE config 1 on the reference's shipped assets: Sponza (163 nodes, 263,911 triangles) static vs the 8,448-triangle sphere of
environment.gltf at scale 1.5, ten single frames.  tests/golden/c1_sponza.npz holds the UNMODIFIED reference's answer on its own trees
(tests/golden/make_golden_c1.py); the assets travel in oracle/_ref/assets/ (git-ignored, copied from the reference checkout by the same
script / by __graft_entry__.build()), and the tests are skipped where they are absent.

  CPU: the library's glTF reader + the port oracle reproduce the golden answer (reader, CreateTriangleList, tree build, sweep, descent,
       tri-tri on real data, including Sponza's non-uniform node scale);
  GPU: imrcd_gltf_load in IMRCD_BUILD_REFERENCE mode + whole frames through the C ABI: the sweep's pair list, leaf combos, triangle-pair
       tests, hit sets and segments bit for bit, ray counts, contact points 1e-5; then the default Morton trees: same colliding set, hit
       set within the documented 3e-4.
"""
import os

import numpy as np
import pytest

import golden_io
from helpers import contacts_close, f32_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "oracle", "_ref", "assets")
SPONZA = os.path.join(ASSETS, "sponzaModel", "Sponza.gltf")
ENV = os.path.join(ASSETS, "environmentTest", "environment.gltf")


def _need_assets():
    if not (os.path.exists(SPONZA) and os.path.exists(ENV)):
        pytest.skip("oracle/_ref/assets is absent (made from the reference checkout by tests/golden/make_golden_c1.py)")


def _triangles_of(checker, prims):
    ps, ns, vs = [], [], []
    for pts, nrm, idx, mode, _ in prims:
        i = np.arange(len(pts), dtype=np.uint32) if idx is None else idx
        p, n, v = checker.triangle_list(pts, nrm, i, mode)
        ps.append(p); ns.append(n); vs.append(v)
    return np.concatenate(ps), np.concatenate(ns), np.concatenate(vs)


def test_port_on_the_shipped_sponza_matches_the_reference(port):
    _need_assets()
    from inmyroom_vulkan_b200.gltf import GltfFile
    z = golden_io.load("c1_sponza")
    with GltfFile(SPONZA) as g:
        assert g.n_meshes == 163
        trees = []
        for m in range(163):
            p, n, v = _triangles_of(port, g.primitives(m))
            assert len(p) == int(z["n_tri"][m])
            trees.append(port.tree_build(p, n, v))
    with GltfFile(ENV) as g:
        p, n, v = _triangles_of(port, g.primitives(0))
        assert len(p) == 8448 == int(z["n_tri"][163])
        sphere = port.tree_build(p, n, v)
    entry_trees = [trees[m] for m in z["node_mesh"]] + [sphere]
    cb = np.zeros(164, np.uint8); cb[163] = 1
    tot = np.zeros(6, np.int64)
    for k in range(len(z["poses"])):
        mats = np.concatenate([z["node_mat"], z["poses"][k][None]]).astype(np.float32)
        pairs, _ = port.broad(mats, entry_trees, cb)
        pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))] if len(pairs) else pairs.reshape(0, 2)
        assert np.array_equal(pairs.astype(np.uint32), z[f"p{k}.pairs"]), k
        tot[0] += len(pairs)
        summ = z[f"p{k}.summary"]
        hn, hi, hs = z[f"p{k}.hit_node"], z[f"p{k}.hit_ids"], z[f"p{k}.hit_seg"]
        for node in range(163):
            a, b = int(summ[node][0]), int(summ[node][1])
            r = port.pair(entry_trees[a], mats[a], entry_trees[b], mats[b])
            assert [r.n_combos, r.n_tri_tests, r.n_hits, r.n_coplanar, int(r.colliding), r.rays_first, r.rays_second] == summ[node][2:9].tolist(), (k, node)
            tot[1:] += [r.n_combos, r.n_tri_tests, r.n_hits, r.n_coplanar, int(r.colliding)]
            if r.n_hits:
                sel = hn == node
                want = {tuple(i): f32_bits(s).tobytes() for i, s in zip(hi[sel].tolist(), hs[sel])}
                got = {tuple(i): f32_bits(s).tobytes() for i, s in zip(np.asarray(r.hit_ids).reshape(-1, 2).tolist(), np.asarray(r.hit_seg, np.float32).reshape(r.n_hits, -1))}
                assert want == got, (k, node)
    assert np.array_equal(tot, z["totals"])
    assert tot[3] > 2000 and tot[4] == 0                     # thousands of hits, none coplanar


def _gpu_c1(gpu_ctx, build_mode):
    from inmyroom_vulkan_b200.collision import CollisionDetection
    from inmyroom_vulkan_b200.gltf import load_gltf
    z = golden_io.load("c1_sponza")
    trees = load_gltf(gpu_ctx, SPONZA, build_mode=build_mode)
    sphere = load_gltf(gpu_ctx, ENV, build_mode=build_mode)[0]
    assert len(trees) == 163
    ids = np.array([trees[m].mesh_id for m in z["node_mesh"]] + [sphere.mesh_id], np.uint32)
    cb = np.zeros(164, np.uint8); cb[163] = 1
    cd = CollisionDetection(ctx=gpu_ctx)
    out = []
    for k in range(len(z["poses"])):
        mats = np.concatenate([z["node_mat"], z["poses"][k][None]]).astype(np.float32)
        cd.Reset(); cd.add_entries(mats, ids, cb, np.arange(164, dtype=np.uint32)); cd.ExecuteCollisionDetection()
        ep, hits = cd.results(want_hits=True)
        out.append((mats, cd.stats(), cd.broad_pairs(), ep, hits))
    return z, out


@pytest.mark.gpu
def test_gpu_c1_reference_trees_bit_exact(gpu_ctx, port):
    _need_assets()
    from inmyroom_vulkan_b200.collision import IMRCD_BUILD_REFERENCE
    z, frames = _gpu_c1(gpu_ctx, IMRCD_BUILD_REFERENCE)
    tot_hits = tot_cop = 0
    for k, (mats, st, bp, ep, hits) in enumerate(frames):
        want_pairs = z[f"p{k}.pairs"]
        assert set(map(tuple, bp.tolist())) == set(map(tuple, want_pairs.tolist())) and len(bp) == len(want_pairs), k
        summ = z[f"p{k}.summary"]
        in_broad = summ[:, 9] == 1
        assert summ[~in_broad, 4].sum() == 0                  # what the sweep does not pair has no hits
        assert st["n_combos"] == summ[in_broad, 2].sum() and st["n_tri_tests"] == summ[in_broad, 3].sum(), k
        assert st["n_hits"] == summ[:, 4].sum() and st["n_coplanar_hits"] == summ[:, 5].sum() and st["n_colliding"] == summ[:, 6].sum(), k
        tot_hits += st["n_hits"]; tot_cop += st["n_coplanar_hits"]
        hn, hi, hs = z[f"p{k}.hit_node"], z[f"p{k}.hit_ids"], z[f"p{k}.hit_seg"]
        first_is_node = {int(s[0]) if s[0] != 163 else int(s[1]): s[0] != 163 for s in summ}
        want = {}
        for node, ids_, seg in zip(hn.tolist(), hi.tolist(), hs):
            pair = (node, 163) if first_is_node[node] else (163, node)
            want[(pair, tuple(ids_))] = f32_bits(seg).tobytes()
        got = {(tuple(bp[h["pair"]].tolist()), (int(h["tri_first"]), int(h["tri_second"]))): f32_bits(np.concatenate([h["source"], h["target"], [h["weight"]]])).tobytes() for h in hits}
        assert got == want, k
        for p in ep:                                            # ray counts exact, contact points 1e-5 (CreateUncollideRays.cpp:131-198)
            a, b = int(p["entry_first"]), int(p["entry_second"])
            node = a if a != 163 else b
            assert (int(p["n_rays_first"]), int(p["n_rays_second"])) == (int(summ[node][7]), int(summ[node][8])), (k, node)
            if summ[node][7] and summ[node][8]:
                assert contacts_close(np.concatenate([p["avg_first"], p["avg_second"]]), z[f"p{k}.avg"][node], port.pair_matrix(mats[a], mats[b])), (k, node)
    assert tot_hits == int(z["totals"][3]) and tot_cop == 0


@pytest.mark.gpu
def test_gpu_c1_morton_trees_same_collisions(gpu_ctx):
    """Which entity of a pair is "first" follows the root boxes' U-minima (SweepAndPrune.cpp:63), i.e. it depends on the tree: pairs and hits
    are compared as (node, node's triangle, sphere's triangle) whatever the orientation."""
    _need_assets()
    from inmyroom_vulkan_b200.collision import IMRCD_BUILD_MORTON
    z, frames = _gpu_c1(gpu_ctx, IMRCD_BUILD_MORTON)
    n_diff = n_hits = 0
    for k, (mats, st, bp, ep, hits) in enumerate(frames):
        summ = z[f"p{k}.summary"]
        want_coll = {min(int(s[0]), int(s[1])) for s in summ if s[6]}
        assert {min(int(p["entry_first"]), int(p["entry_second"])) for p in ep} == want_coll, k
        hn, hi = z[f"p{k}.hit_node"], z[f"p{k}.hit_ids"]
        node_first = {min(int(s[0]), int(s[1])): s[0] != 163 for s in summ}
        want = {(n, i[0], i[1]) if node_first[n] else (n, i[1], i[0]) for n, i in zip(hn.tolist(), hi.tolist())}
        got = set()
        for h in hits:
            a, b = bp[h["pair"]].tolist()
            got.add((a, int(h["tri_first"]), int(h["tri_second"])) if b == 163 else (b, int(h["tri_second"]), int(h["tri_first"])))
        n_diff += len(got ^ want); n_hits += len(want)
    assert n_hits == int(z["totals"][3]) and n_diff <= 3e-4 * n_hits + 1, (n_diff, n_hits)
