"""BASELINE config 1 on the reference's SHIPPED assets: Sponza (163 mesh nodes, 262,267 collision triangles) static vs one dynamic
`environment.gltf` mesh 0 (8,448 triangles, scale 1.5), single frames on the CPU.  Writes tests/golden/c1_sponza.npz from the UNMODIFIED
reference (oracle/_ref/libimr_ref.so + its own tinygltf reader, libimr_ref_gltf.so) and copies the two assets (data, not source) into
oracle/_ref/assets/, which is git-ignored and travels to the GPU box with the compiled reference.  Run here, in the build container:

    python tests/golden/make_golden_c1.py

What is stored: the 163 node matrices as the engine builds them (GameImporter.cpp:530-541: rotation (w, x, -y, -z), translation
(x, -y, -z); NodeDataCompEntity.cpp:63-69: T * R * S), ten poses of the sphere, and per pose the reference's own answer with its own
trees: the sweep's pair list (SweepAndPrune.cpp:15-88) and, for EVERY node paired with the sphere (the survey's known-answer set-up,
SURVEY 8c), leaf combos, triangle-pair tests, coplanar count, the non-coplanar hits (node, triA, triB in ORIGINAL triangle order) with
their segments, ray counts and contact points.  The survey's own probe (`mt19937(7)`, 117,699 combos / 2,394 hits) was a throw-away
program whose pose recipe is not recorded; this file fixes a recipe (numpy default_rng(7), below) and stores the reference's answer to it.
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from inmyroom_vulkan_b200 import scenes  # noqa: E402
from oracle import bind  # noqa: E402

REF_ASSETS = "/root/reference/inMyRoom_vulkan/testGames/Sponza"
FILES = ["sponzaModel/Sponza.gltf", "sponzaModel/Sponza.bin", "environmentTest/environment.gltf", "environmentTest/environment.bin"]
ASSET_DIR = os.path.join(ROOT, "oracle", "_ref", "assets")
SPHERE_SCALE = 1.5
LO = np.array([-56.97, -54.20, -23.26]); HI = np.array([52.24, 0.10, 25.68])     # Sponza's root-centre bounds in ENGINE space (y, z negated;
                                                                                  # SURVEY 8d, C1 quotes them before the flip)


def copy_assets():
    for f in FILES:
        dst = os.path.join(ASSET_DIR, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(os.path.join(REF_ASSETS, f)):
            shutil.copyfile(os.path.join(REF_ASSETS, f), dst)
    return ASSET_DIR


def node_matrices(gltf_path):
    """(mesh index, mat4 column-major f32) of every mesh node, the engine's conventions (GameImporter.cpp:530-541)."""
    g = json.load(open(gltf_path))
    t, q, s, mesh = [], [], [], []
    for n in g["nodes"]:
        if "mesh" not in n:
            continue
        assert "children" not in n and "matrix" not in n
        tr = n.get("translation", [0, 0, 0]); ro = n.get("rotation", [0, 0, 0, 1]); sc = n.get("scale", [1, 1, 1])
        t.append([tr[0], -tr[1], -tr[2]]); q.append([ro[0], -ro[1], -ro[2], ro[3]]); s.append(sc); mesh.append(n["mesh"])      # (x, y, z, w) for scenes.trs_matrices
    q = np.asarray(q, np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)          # NodeDataCompEntity.cpp:38 normalises
    return np.asarray(mesh, np.uint32), scenes.trs_matrices(np.asarray(t, np.float32), q, np.asarray(s, np.float32))


def triangles_of(orc, prims):
    ps, ns, vs = [], [], []
    for pts, nrm, idx, mode, _ in prims:
        i = np.arange(len(pts), dtype=np.uint32) if idx is None else idx
        p, n, v = orc.triangle_list(pts, nrm, i, mode)
        ps.append(p); ns.append(n); vs.append(v)
    return np.concatenate(ps), np.concatenate(ns), np.concatenate(vs)


def sphere_poses(n=10, seed=7):
    rng = np.random.default_rng(seed)
    t = (LO + (HI - LO) * rng.random((n, 3))).astype(np.float32)
    q = scenes.random_quaternions(rng, n)
    return scenes.trs_matrices(t, q, np.full((n, 3), SPHERE_SCALE, np.float32))


def main():
    assets = copy_assets()
    bind.build("ref")
    orc = bind.RefOracle()
    port = bind.PortOracle()
    sponza = bind.RefGltf(os.path.join(assets, FILES[0]))
    env = bind.RefGltf(os.path.join(assets, FILES[2]))
    node_mesh, node_mat = node_matrices(os.path.join(assets, FILES[0]))
    assert sponza.n_meshes == 163 and len(node_mesh) == 163
    trees, n_tri = [], []
    for m in range(sponza.n_meshes):
        p, n, v = triangles_of(orc, sponza.primitives(m))
        trees.append(orc.tree_build(p, n, v)); n_tri.append(len(p))
    sp, sn, sv = triangles_of(orc, env.primitives(0))
    sphere = orc.tree_build(sp, sn, sv)
    print(f"Sponza: {sum(n_tri)} collision triangles in {len(trees)} meshes (min {min(n_tri)}, max {max(n_tri)}); sphere: {len(sp)}")
    poses = sphere_poses()
    out = dict(node_mesh=node_mesh, node_mat=node_mat, poses=poses, n_tri=np.asarray(n_tri + [len(sp)], np.uint32),
               sphere_scale=np.float32(SPHERE_SCALE))
    entry_trees = [trees[m] for m in node_mesh] + [sphere]
    cb = np.zeros(164, np.uint8); cb[163] = 1
    tot = dict(combos=0, tests=0, hits=0, coplanar=0, colliding=0, pairs=0)
    for k, pose in enumerate(poses):
        mats = np.concatenate([node_mat, pose[None]]).astype(np.float32)
        pairs, _ = orc.broad(mats, entry_trees, cb)
        # orientation on exactly equal U-minima: "lower entry index first" 
        rb = np.stack([t.root_box for t in entry_trees]).astype(np.float32)
        umin = port.extents(mats, rb)[:, 0]
        tie = (umin[pairs[:, 0]] == umin[pairs[:, 1]]) & (pairs[:, 0] > pairs[:, 1])
        pairs[tie] = pairs[tie][:, ::-1]
        pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
        broad_set = set(map(tuple, pairs.tolist()))
        summ, hit_node, hit_ids, hit_seg, avg = [], [], [], [], []
        for node in range(163):
            # the pair as the sweep orients it (earlier on U first); a node the sweep does not pair with the sphere is tested "node first"
            first_is_node = (node, 163) in broad_set or (163, node) not in broad_set
            a, b = (node, 163) if first_is_node else (163, node)
            r = orc.pair(entry_trees[a], mats[a], entry_trees[b], mats[b])
            summ.append([a, b, r.n_combos, r.n_tri_tests, r.n_hits, r.n_coplanar, int(r.colliding), r.rays_first, r.rays_second, int((a, b) in broad_set)])
            if r.n_hits:
                hit_node.append(np.full(r.n_hits, node, np.uint32)); hit_ids.append(np.asarray(r.hit_ids, np.uint32).reshape(-1, 2)); hit_seg.append(np.asarray(r.hit_seg, np.float32).reshape(r.n_hits, -1))
            avg.append(np.asarray(r.avg, np.float32) if r.colliding else np.full(6, np.nan, np.float32))
            tot["combos"] += r.n_combos; tot["tests"] += r.n_tri_tests; tot["hits"] += r.n_hits; tot["coplanar"] += r.n_coplanar; tot["colliding"] += int(r.colliding)
        tot["pairs"] += len(pairs)
        out[f"p{k}.pairs"] = pairs.astype(np.uint32)
        out[f"p{k}.summary"] = np.asarray(summ, np.int64)
        out[f"p{k}.hit_node"] = np.concatenate(hit_node) if hit_node else np.zeros(0, np.uint32)
        out[f"p{k}.hit_ids"] = np.concatenate(hit_ids) if hit_ids else np.zeros((0, 2), np.uint32)
        out[f"p{k}.hit_seg"] = np.concatenate(hit_seg) if hit_seg else np.zeros((0, 7), np.float32)
        out[f"p{k}.avg"] = np.stack(avg)
    out["totals"] = np.asarray([tot["pairs"], tot["combos"], tot["tests"], tot["hits"], tot["coplanar"], tot["colliding"]], np.int64)
    print("10 poses, every node against the sphere:", tot)
    np.savez_compressed(os.path.join(HERE, "c1_sponza.npz"), **out)
    print("wrote", os.path.join(HERE, "c1_sponza.npz"), os.path.getsize(os.path.join(HERE, "c1_sponza.npz")), "bytes")


if __name__ == "__main__":
    main()
