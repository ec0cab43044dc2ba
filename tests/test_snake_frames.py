"""Frames of the REAL SnakeGame as parity vectors (tests/golden/snake_frames.npz, made by `make_golden.py snake` from the game DLL running
under the reference's ECS, general components and CollisionDetection): the entries exactly as ModelCollisionComp::Update makes them
(armatures scaled x4, animated bones, bodies at rest on the floor, snakes running into pillars, apples and each other) and the engine's
verdict -- every colliding (entity, other) with the deltaVector handed to the components (CollisionDetection.cpp:60-103).
CPU: the port oracle against it; GPU: the library (trees built on the device in IMRCD_BUILD_REFERENCE mode, frames through the Python
mirror of the reference interface) against it."""
import numpy as np
import pytest

import golden_io

Z = golden_io.load("snake_frames")
FRAMES = Z["frames"].tolist()


def _meshes():
    return [(Z[f"mesh{k}.points"], Z[f"mesh{k}.normals"], Z[f"mesh{k}.indices"]) for k in range(int(Z["n_meshes"][0]))]


def _want(f):
    return {(int(a), int(b)): d for (a, b), d in zip(Z[f"f{f}.pairs"].tolist(), Z[f"f{f}.deltas"])}


def _check(got, want, rel, floor):
    assert set(got) == set(want), (sorted(set(want) - set(got))[:4], sorted(set(got) - set(want))[:4])
    for k, w in want.items():
        g = np.asarray(got[k], np.float64); w = np.asarray(w, np.float64)
        if np.isnan(w).any():
            assert np.isnan(g).any(), k
            continue
        assert np.linalg.norm(g - w) <= rel * np.linalg.norm(w) + floor, (k, g, w)


def test_fixture_is_a_real_game():
    rows = sum(len(Z[f"f{f}.pairs"]) for f in FRAMES)
    assert len(FRAMES) >= 20 and rows >= 500
    f = FRAMES[-1]
    assert len(Z[f"f{f}.entity"]) == 23 + 12 and (Z[f"f{f}.callback"] == 1).sum() == 12          # the map's solids + one collision sphere per snake
    moving = (Z[f"f{f}.cur"] != Z[f"f{f}.prev"]).any(1)
    assert moving.sum() == 12                                                                       # only the snakes move


def test_port_reproduces_the_engines_verdict(port):
    trees = []
    for pts, nrm, idx in _meshes():
        p, n, v = port.triangle_list(pts, nrm, idx, 4)
        trees.append(port.tree_build(p, n, v))
    for f in FRAMES:
        ent = Z[f"f{f}.entity"]; cur = Z[f"f{f}.cur"]; prev = Z[f"f{f}.prev"]
        et = [trees[m] for m in Z[f"f{f}.mesh"]]
        pairs, _ = port.broad(cur, et, Z[f"f{f}.callback"])
        got = {}
        for a, b in pairs.tolist():
            col, d1, d2 = port.pair_delta(et[a], cur[a], prev[a], et[b], cur[b], prev[b])
            if col:
                got[(int(ent[a]), int(ent[b]))] = d1; got[(int(ent[b]), int(ent[a]))] = d2
        # same trees (the port's build is bit-identical), same rays in the same order: only the order of the contact-point sums differs
        _check(got, _want(f), rel=2e-5, floor=2e-7)


@pytest.mark.gpu
def test_device_reproduces_the_engines_verdict(gpu_ctx):
    from inmyroom_vulkan_b200.collision import IMRCD_BUILD_REFERENCE, CollisionDetection, OBBtree
    trees = [OBBtree.from_primitives(gpu_ctx, [(pts, nrm, idx, 4)], build_mode=IMRCD_BUILD_REFERENCE) for pts, nrm, idx in _meshes()]
    cd = CollisionDetection(ctx=gpu_ctx)
    n_rows = 0
    for f in FRAMES:
        cd.Reset()
        mesh_ids = np.array([trees[m].mesh_id for m in Z[f"f{f}.mesh"]], np.uint32)
        cd.add_entries(Z[f"f{f}.cur"], mesh_ids, Z[f"f{f}.callback"], Z[f"f{f}.entity"], Z[f"f{f}.prev"])
        cd.ExecuteCollisionDetection()
        ep, _ = cd.results(want_hits=False)
        got = {}
        for p in ep:
            got[(int(p["entity_first"]), int(p["entity_second"]))] = p["delta_first"]
            got[(int(p["entity_second"]), int(p["entity_first"]))] = p["delta_second"]
        # deltaVectors of resting contacts are ~1e-3 units from ray hits at coordinates of ~10: a few FP32 ulps of the coordinates
        # (2e-5) is the precision such a delta has once the ray origins are sums in another order (tests/test_gpu_snake_game.py)
        _check(got, _want(f), rel=1e-4, floor=2e-5)
        n_rows += len(got)
    assert n_rows >= 500
