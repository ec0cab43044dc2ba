"""Shared helpers for the parity tests: run a whole frame through an oracle."""
from __future__ import annotations

import numpy as np


def f32_bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_entity_pairs(a, b, rtol=1e-6):
    """Two runs of the same frame: identical colliding pairs, hit and ray counts; contact points equal up to FP32 rounding.
    The contact averages are sums over unordered sets (std::unordered_map iteration in the reference, CreateUncollideRays.cpp:138,185-198;
    order-free FP64 atomics here), so the last bit of a component may differ between runs; everything integral is exact."""
    assert len(a) == len(b)
    for f in ("entry_first", "entry_second", "entity_first", "entity_second", "n_hits", "n_rays_first", "n_rays_second", "flags"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("avg_first", "avg_second"):
        x = np.asarray(a[f], np.float64); y = np.asarray(b[f], np.float64)
        ok = np.linalg.norm(x - y, axis=1) <= rtol * np.maximum(np.linalg.norm(y, axis=1), 1e-30)
        ok |= np.isnan(x).any(1) & np.isnan(y).any(1)           # 0 rays on a side: NaN like the reference
        assert ok.all(), (f, x[~ok][:3], y[~ok][:3])


def contacts_close(avg, gold, rel, rtol=1e-5):
    """Contact points agree within `rtol` relative (north star): average_point_first lives in first's model space;
    average_point_second is inverse(rel) * (a first-space point) (CreateUncollideRays.cpp:191-198), so its rounding noise
    scales with the FIRST-space magnitude -- compare it after mapping back with rel, relative to the vector norms."""
    avg = np.asarray(avg, np.float64); gold = np.asarray(gold, np.float64)
    M = np.asarray(rel, np.float64).reshape(4, 4).T           # column-major -> math layout
    def fwd(p):
        return M[:3, :3] @ p + M[:3, 3]
    a_ok = np.linalg.norm(avg[:3] - gold[:3]) <= rtol * max(np.linalg.norm(gold[:3]), 1e-30)
    b1, b2 = fwd(avg[3:]), fwd(gold[3:])
    b_ok = np.linalg.norm(b1 - b2) <= 4 * rtol * max(np.linalg.norm(b2), np.linalg.norm(M[:3, 3]), 1e-30)
    return bool(a_ok and b_ok)


def oracle_frame(orc, scene, trees, port=None, max_pairs=None):
    """Broad + mid + narrow through an oracle.  `trees` = one oracle Tree per scene mesh.
    Returns dict(pairs=(k,2) ordered entry pairs, per_pair={(a,b): PairResult}, totals)."""
    entry_trees = [trees[m] for m in scene.mesh_index]
    if getattr(orc, "kind", "") == "reference" and scene.n_entries > 65534:
        raise ValueError("reference broad phase cannot take > 65534 entries")
    pairs, _ = orc.broad(scene.matrices, entry_trees, scene.should_callback)
    # Orientation = order on the U axis (SweepAndPrune.cpp:63).  For EXACTLY equal U-minima the reference's
    # std::sort leaves the order unspecified; the C ABI (and the port oracle) define it as "lower entry index
    # first".  Normalise the real reference's output to that rule before comparing (documented in DESIGN.md).
    if len(pairs):
        from oracle import bind as _bind
        _port = port or _bind.PortOracle()
        rb = np.stack([t.root_box for t in entry_trees]).astype(np.float32)
        umin = _port.extents(scene.matrices, rb)[:, 0]
        tie = (umin[pairs[:, 0]] == umin[pairs[:, 1]]) & (pairs[:, 0] > pairs[:, 1])
        pairs[tie] = pairs[tie][:, ::-1]
    order = np.lexsort((pairs[:, 1], pairs[:, 0])) if len(pairs) else np.zeros(0, np.int64)
    pairs = pairs[order]
    per_pair = {}
    tot = dict(combos=0, tri_tests=0, hits=0, coplanar=0, colliding=0)
    for k, (a, b) in enumerate(pairs.tolist()):
        if max_pairs is not None and k >= max_pairs:
            break
        r = orc.pair(entry_trees[a], scene.matrices[a], entry_trees[b], scene.matrices[b])
        per_pair[(a, b)] = r
        tot["combos"] += r.n_combos; tot["tri_tests"] += r.n_tri_tests; tot["hits"] += r.n_hits
        tot["coplanar"] += r.n_coplanar; tot["colliding"] += int(r.colliding)
    return dict(pairs=pairs, per_pair=per_pair, totals=tot)


def gpu_frame(cd, scene, gpu_trees):
    """Run one frame through the C ABI; returns (stats, broad_pairs, entity_pairs, hits)."""
    cd.Reset()
    mesh_ids = np.array([gpu_trees[m].mesh_id for m in scene.mesh_index], np.uint32)
    cd.add_entries(scene.matrices, mesh_ids, scene.should_callback, scene.entities, scene.previous)
    cd.ExecuteCollisionDetection()
    st = cd.stats()
    bp = cd.broad_pairs()
    ep, hits = cd.results(want_hits=True)
    return st, bp, ep, hits


def compare_frame(orc_res, st, bp, ep, hits, check_hits_bits=True, check_contacts=True, rel_of=None):
    """Assert the GPU frame equals the oracle frame: ordered pair set, per-pair hit sets (bit-exact segments),
    colliding-entity set, counters."""
    o_pairs = set(map(tuple, orc_res["pairs"].tolist()))
    g_pairs = set(map(tuple, bp.tolist()))
    assert len(bp) == len(g_pairs), "GPU emitted a duplicate broad-phase pair"
    assert g_pairs == o_pairs, f"broad-phase pair sets differ: only-gpu={sorted(g_pairs - o_pairs)[:5]} only-oracle={sorted(o_pairs - g_pairs)[:5]}"
    tot = orc_res["totals"]
    if len(orc_res["per_pair"]) == len(o_pairs):
        assert st["n_combos"] == tot["combos"]
        assert st["n_tri_tests"] == tot["tri_tests"]
        assert st["n_hits"] == tot["hits"]
        assert st["n_coplanar_hits"] == tot["coplanar"]
        assert st["n_colliding"] == tot["colliding"]
    # per pair hit sets
    g_by_pair = {}
    for h in hits:
        key = tuple(bp[h["pair"]].tolist())
        g_by_pair.setdefault(key, []).append(h)
    for key, r in orc_res["per_pair"].items():
        gh = g_by_pair.get(key, [])
        assert len(gh) == r.n_hits, f"pair {key}: {len(gh)} GPU hits vs {r.n_hits} oracle hits"
        if r.n_hits == 0:
            continue
        o_rec = {(int(a), int(b)): f32_bits(seg).tobytes() for (a, b), seg in zip(r.hit_ids.tolist(), r.hit_seg)}
        g_rec = {(int(h["tri_first"]), int(h["tri_second"])): f32_bits(np.concatenate([h["source"], h["target"], [h["weight"]]])).tobytes() for h in gh}
        assert set(o_rec) == set(g_rec), f"pair {key}: triangle-pair sets differ"
        if check_hits_bits:
            bad = [k for k in o_rec if o_rec[k] != g_rec[k]]
            assert not bad, f"pair {key}: {len(bad)} hit segments differ in bits, e.g. {bad[:3]}"
    o_coll = {k for k, r in orc_res["per_pair"].items() if r.colliding}
    g_coll = {(int(p["entry_first"]), int(p["entry_second"])) for p in ep}
    # contact reduction (CreateUncollideRays.cpp:117-198): ray counts exact, contact points within 1e-5 relative
    if check_contacts:
        for p in ep:
            key = (int(p["entry_first"]), int(p["entry_second"]))
            r = orc_res["per_pair"].get(key)
            if r is None:
                continue
            assert (int(p["n_rays_first"]), int(p["n_rays_second"])) == (r.rays_first, r.rays_second), f"pair {key}: ray counts {p['n_rays_first']},{p['n_rays_second']} vs {r.rays_first},{r.rays_second}"
            assert int(p["n_hits"]) == r.n_hits
            if rel_of is not None and r.rays_first and r.rays_second:
                assert contacts_close(np.concatenate([p["avg_first"], p["avg_second"]]), r.avg, rel_of(key)), f"pair {key}: contact points {p['avg_first']} {p['avg_second']} vs {r.avg}"
    if len(orc_res["per_pair"]) == len(o_pairs):
        assert g_coll == o_coll, "colliding-entity sets differ"
    else:
        assert o_coll <= g_coll
