"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libimr_ref.so, compiled in place from
/root/reference by oracle/Makefile).  Run here, in the build container; the fixtures are committed and travel to the
GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Every array below is an output of the reference's own translation units (IEEE build: -O2 -ffp-contract=off) on the
stored inputs.  Floats are compared by bit pattern in the tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from inmyroom_vulkan_b200 import scenes  # noqa: E402
from oracle import bind  # noqa: E402


def rand_mats(rng, n, tscale=1.0):
    s = rng.random((n, 3)) * 1.5 + 0.25
    return scenes.trs_matrices(rng.normal(size=(n, 3)) * tscale, scenes.random_quaternions(rng, n), s)


def degenerate_tris(rng, count=800):
    A, B = [], []
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    for k in range(count):
        t = base * np.float32(0.5 + rng.random()) + rng.normal(size=3).astype(np.float32) * np.float32(k % 3 == 0)
        kind = k % 8
        if kind == 0:
            u = t + np.array([0.2, 0.1, 0], np.float32)
        elif kind == 1:
            u = t + np.array([5, 5, 0], np.float32)
        elif kind == 2:
            u = t.copy(); u[2] = t[2] + np.array([0, 0, 1], np.float32)
        elif kind == 3:
            u = t.copy(); u[1] += np.array([0, 0.3, 1], np.float32); u[2] += np.array([0.4, 0, -1], np.float32)
        elif kind == 4:
            u = np.stack([t[0], t[0], t[1]])
        elif kind == 5:
            u = t + np.array([0.1, 0.1, 5e-7], np.float32); u[2] += np.array([0, 0, 1], np.float32)
        elif kind == 6:
            u = t.copy()
        else:
            u = np.array([[0.2, 0.2, -1], [0.3, 0.2, 1], [0.2, 0.3, 1]], np.float32) * np.float32(0.5 + rng.random())
        A.append(t.reshape(9)); B.append(u.reshape(9))
    return np.array(A, np.float32), np.array(B, np.float32)


def make_response(ref):
    """response.npz: Ray::IntersectOBBtree on one tree, and the response stage (rays in the reference's order, the delta
    ShootUncollideRays returns for them, the deltaVectors of CollisionDetection.cpp:80-103) on the torus_instances frame."""
    import golden_io
    rng = np.random.default_rng(20261018)
    out = {}
    mesh = scenes.torus(20, 10)
    tree = ref.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids)
    n = 1500
    mats = rand_mats(rng, n, 0.3)
    o = (rng.normal(size=(n, 3)) * np.where(np.arange(n) % 3 == 0, 3.0, 0.2)[:, None]).astype(np.float32); o[::50] = 0
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    res = [ref.ray_tree(tree, mats[k], o[k], d[k]) for k in range(n)]
    out["ray.mats"] = mats; out["ray.origins"] = o; out["ray.dirs"] = d
    out["ray.hit"] = np.array([r[0] for r in res], np.uint8); out["ray.back"] = np.array([r[1] for r in res], np.uint8)
    out["ray.dist"] = np.array([r[2] for r in res], np.float32); out["ray.bary"] = np.stack([r[3] for r in res]).astype(np.float32)
    out["ray.tri"] = np.array([r[4] & 0xffffffff for r in res], np.uint32)
    sc, gold = golden_io.golden_frame(golden_io.load("frames"), "torus_instances")
    trees = [ref.tree_build(m.positions, m.normals, m.vertex_ids) for m in sc.meshes]
    prev = sc.matrices.copy()
    prev[:, 12:15] += (rng.normal(size=(sc.n_entries, 3)) * 0.03).astype(np.float32)
    prev[::5] = sc.matrices[::5]
    rays1, rays2, cnt, shoot, delta, col = [], [], [], [], [], []
    for (i, j) in gold["pairs"].tolist():
        ta, tb = trees[sc.mesh_index[i]], trees[sc.mesh_index[j]]
        r1, r2 = ref.pair_rays(ta, sc.matrices[i], tb, sc.matrices[j])
        dd, _ = ref.shoot(ta, sc.matrices[i], tb, sc.matrices[j], r1, r2) if (len(r1) or len(r2)) else (np.zeros(3, np.float32), None)
        c, d1, d2 = ref.pair_delta(ta, sc.matrices[i], prev[i], tb, sc.matrices[j], prev[j])
        rays1.append(r1); rays2.append(r2); cnt.append([len(r1), len(r2)]); shoot.append(dd); delta.append(np.concatenate([d1, d2])); col.append(int(c))
    out["frame.previous"] = prev
    out["frame.rays_first"] = np.concatenate(rays1).astype(np.float32).reshape(-1, 6); out["frame.rays_second"] = np.concatenate(rays2).astype(np.float32).reshape(-1, 6)
    out["frame.ray_counts"] = np.array(cnt, np.int64); out["frame.shoot"] = np.array(shoot, np.float32); out["frame.delta"] = np.array(delta, np.float32)
    out["frame.colliding"] = np.array(col, np.uint8)
    np.savez_compressed(os.path.join(HERE, "response.npz"), **out)
    print("response: rays hit", int(out["ray.hit"].sum()), "of", n, "; colliding pairs", int(sum(col)), "non-zero deltas", int((np.abs(np.nan_to_num(out["frame.delta"])).sum(1) > 0).sum()))


def make_primitives(ref):
    """primitives.npz: Triangle::CreateTriangleList (Triangle.cpp:9-62,259-280) on every draw mode, with and without normals."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_primitives import _cases
    out = {}
    names = []
    for name, pts, nrm, idx, mode in _cases():
        p, n, v = ref.triangle_list(pts, nrm, idx, mode)
        names.append(name)
        out[f"{name}.points"] = pts; out[f"{name}.indices"] = idx; out[f"{name}.mode"] = np.array([mode])
        if nrm is not None:
            out[f"{name}.normals"] = nrm
        out[f"{name}.pos"] = p; out[f"{name}.nrm"] = n; out[f"{name}.vid"] = v
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)
    print("primitives:", len(names), "cases,", sum(len(out[f"{k}.pos"]) for k in names), "triangles")


def make_gltf(ref):
    """gltf_scene.glb + gltf.npz: a file with every loader situation (tests/gltf_writer.py), read by the reference's own reader (tinygltf) in the
    engine's primitive order and extraction (oracle/ref_gltf_shim.cpp), through the reference's Triangle::CreateTriangleList."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gltf_writer
    path = gltf_writer.write(os.path.join(HERE, "gltf_scene.glb"), gltf_writer.sample_meshes(), "glb")
    g = bind.RefGltf(path)
    out = {"n_meshes": np.array([g.n_meshes])}
    for m in range(g.n_meshes):
        ps, ns, vs = [np.zeros((0, 9), np.float32)], [np.zeros((0, 9), np.float32)], [np.zeros((0, 3), np.uint32)]
        for pts, nrm, idx, mode, _ in g.primitives(m):
            p, n, v = ref.triangle_list(pts, nrm, np.arange(len(pts), dtype=np.uint32) if idx is None else idx, mode)
            ps.append(p); ns.append(n); vs.append(v)
        out[f"m{m}.pos"] = np.concatenate(ps); out[f"m{m}.nrm"] = np.concatenate(ns); out[f"m{m}.vid"] = np.concatenate(vs)
    np.savez_compressed(os.path.join(HERE, "gltf.npz"), **out)
    print("gltf:", g.n_meshes, "meshes,", [len(out[f"m{m}.pos"]) for m in range(g.n_meshes)], "triangles,", os.path.getsize(path), "bytes")


def make_snake():
    """snake_frames.npz: frames of the REAL SnakeGame (its game DLL under the reference's ECS, general components and CollisionDetection,
    oracle/_ref/snake_harness_record, 12 snakes, fixed 1/60 s clock): the scene's meshes, the entries ModelCollisionComp::Update made in
    a selection of frames, and the engine's verdict on each -- every colliding (entity, other) with the deltaVector it was handed."""
    import subprocess
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import snake_dump
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "oracle"), "harness"], check=True)
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    with tempfile.TemporaryDirectory() as d:
        dump = os.path.join(d, "dump.txt")
        subprocess.run([os.path.join(ref_dir, "snake_harness_record"), "200", "12", os.path.join(ref_dir, "libsnake_game.so")],
                       check=True, env=dict(os.environ, SNAKE_DUMP=dump), stdout=subprocess.DEVNULL)
        meshes, frames = snake_dump.load(dump)
    busiest = sorted(range(len(frames)), key=lambda f: -len(frames[f]["callbacks"]))[:6]
    keep = sorted(set([0, 1, 2, 5, 10, 20, 35, 50, 75, 100, 125, 150, 175, 199] + busiest))
    out = {"n_meshes": np.array([len(meshes)]), "frames": np.array(keep)}
    for k, (pts, nrm, idx) in enumerate(meshes):
        out[f"mesh{k}.points"] = pts; out[f"mesh{k}.normals"] = nrm; out[f"mesh{k}.indices"] = idx
    n_rows = 0
    for f in keep:
        fr = frames[f]
        for key in ("entity", "mesh", "callback", "cur", "prev"):
            out[f"f{f}.{key}"] = fr[key]
        rows = fr["callbacks"]
        out[f"f{f}.pairs"] = np.array([[a, b] for a, b, _ in rows], np.uint32).reshape(-1, 2)
        out[f"f{f}.deltas"] = np.array([d for _, _, d in rows], np.float32).reshape(-1, 3)
        n_rows += len(rows)
    np.savez_compressed(os.path.join(HERE, "snake_frames.npz"), **out)
    print("snake:", len(keep), "frames,", n_rows, "colliding (entity, other) rows,", os.path.getsize(os.path.join(HERE, "snake_frames.npz")), "bytes")


def main():
    bind.build("ref")
    ref = bind.RefOracle()
    if len(sys.argv) > 1 and sys.argv[1] == "snake":
        return make_snake()
    if len(sys.argv) > 1 and sys.argv[1] == "gltf":
        return make_gltf(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "response":
        return make_response(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "primitives":
        return make_primitives(ref)
    rng = np.random.default_rng(20261017)

    # ---- predicates ----
    n = 1500
    a = rng.normal(size=(n, 12)).astype(np.float32); b = rng.normal(size=(n, 12)).astype(np.float32)
    b[:, :3] *= 2.5
    b[:100, 3:] = a[:100, 3:]              # parallel axes: zero cross products
    a[100:140, 3:6] = 0                    # degenerate (flat) boxes
    mats = rand_mats(rng, n)
    verdict = np.array([ref.sat(a[i], b[i], mats[i]) for i in range(n)], np.uint8)
    surf_a = np.array([ref.surface(a[i]) for i in range(n)], np.float32)
    surf_b = np.array([ref.surface(b[i], mats[i]) for i in range(n)], np.float32)
    xform = np.stack([ref.box_transform(b[i], mats[i]) for i in range(n)])
    np.savez_compressed(os.path.join(HERE, "sat.npz"), a=a, b=b, mats=mats, verdict=verdict, surf_a=surf_a, surf_b=surf_b, xform=xform)

    n = 4000
    ta = rng.normal(size=(n, 9)).astype(np.float32); tb = (rng.normal(size=(n, 9)) * 0.8).astype(np.float32)
    m = rand_mats(rng, 1, 0.2)[0]
    f, s = ref.tri_tri(ta, tb, m)
    da, db = degenerate_tris(rng)
    df, ds = ref.tri_tri(da, db, None)
    np.savez_compressed(os.path.join(HERE, "tri_tri.npz"), a=ta, b=tb, m=m, flags=f, seg=s, da=da, db=db, dflags=df, dseg=ds)

    n = 400
    ma = rand_mats(rng, n, 20.0); mb = rand_mats(rng, n, 20.0)
    ma[:20] = scenes.trs_matrices(np.zeros((20, 3)), np.tile([[0.70710678, 0, 0, 0.70710678]], (20, 1)),
                                  np.tile([[0.0399999991, 0.0400000028, 0.0400000028]], (20, 1)))
    rel = np.stack([ref.pair_matrix(ma[i], mb[i]) for i in range(n)])
    np.savez_compressed(os.path.join(HERE, "pair_matrix.npz"), a=ma, b=mb, rel=rel, axes=ref.sweep_axes())

    # ---- OBB fit (OBB.cpp:33-166 incl. the rows-of-V quirk, SURVEY finding 3) ----
    clouds, boxes, sizes = [], [], []
    for k in range(40):
        cnt = int(rng.integers(3, 60))
        p = (rng.normal(size=(cnt, 3)) * (rng.random(3) * 3 + 0.1) + rng.normal(size=3) * 5).astype(np.float32)
        if k % 10 == 0:
            p[:] = p[0]                   # a single unique point
        if k % 10 == 1:
            p[:, 2] = p[0, 2]             # planar cloud
        clouds.append(p); sizes.append(cnt); boxes.append(ref.obb_from_points(p))
    np.savez_compressed(os.path.join(HERE, "obb_fit.npz"), points=np.concatenate(clouds), sizes=np.array(sizes, np.int64), boxes=np.stack(boxes))

    # ---- trees (OBBtree.cpp:321) ----
    tree_meshes = {"torus20x10": scenes.torus(20, 10), "box3": scenes.box_mesh(1, 2, 3, sub=3), "sphere12x9": scenes.uv_sphere(12, 9),
                   "tiny4": None, "one": None}
    bm = scenes.box_mesh(1, 1, 1, sub=1)
    tree_meshes["tiny4"] = scenes.Mesh(bm.positions[:4].copy(), bm.normals[:4].copy(), bm.vertex_ids[:4].copy(), "tiny4")
    tree_meshes["one"] = scenes.Mesh(bm.positions[:1].copy(), bm.normals[:1].copy(), bm.vertex_ids[:1].copy(), "one")
    out = {}
    for name, msh in tree_meshes.items():
        ft = ref.tree_build(msh.positions, msh.normals, msh.vertex_ids).flat
        for fld in ("boxes", "left", "right", "tri_off", "tri_cnt", "tri_pos", "tri_nrm", "tri_vid", "tri_orig"):
            out[f"{name}.{fld}"] = getattr(ft, fld)
        out[f"{name}.in_pos"] = msh.positions; out[f"{name}.in_nrm"] = msh.normals; out[f"{name}.in_vid"] = msh.vertex_ids
    np.savez_compressed(os.path.join(HERE, "trees.npz"), **out)

    # ---- whole frames: broad + mid + narrow with the reference's own trees ----
    from helpers import oracle_frame
    port = None
    try:
        bind.build("port"); port = bind.PortOracle()
    except Exception:
        pass
    frames = {}
    static = scenes.atrium_static(detail=1)
    keep = [0, 3, 7, 8, 64, 120, 125, 130]
    small_static = ([scenes.Mesh(static[0][i].positions[::7].copy(), static[0][i].normals[::7].copy(), static[0][i].vertex_ids[::7].copy(), static[0][i].name)
                     for i in keep], static[1][keep])
    cases = {
        "torus_instances": scenes.scene_instances(scenes.torus(20, 10), 40, seed=77, neighbours=6.0),
        "static_vs_bodies": scenes.scene_static_vs_bodies(scenes.uv_sphere(12, 9), 240, seed=5, body_scale=(1.0, 3.0), static=small_static),
    }
    for name, sc in cases.items():
        trees = [ref.tree_build(msh.positions, msh.normals, msh.vertex_ids) for msh in sc.meshes]
        res = oracle_frame(ref, sc, trees, port=port)
        pairs = res["pairs"]
        hit_pair, hit_ids, hit_seg, summ, avg = [], [], [], [], []
        for k, (pa, pb) in enumerate(pairs.tolist()):
            r = res["per_pair"][(pa, pb)]
            hit_pair.append(np.full(r.n_hits, k, np.uint32)); hit_ids.append(r.hit_ids.reshape(-1, 2)); hit_seg.append(r.hit_seg.reshape(-1, 7))
            summ.append([r.n_combos, r.n_tri_tests, r.n_hits, r.n_coplanar, r.rays_first, r.rays_second, int(r.colliding)])
            avg.append(r.avg)
        frames[f"{name}.matrices"] = sc.matrices; frames[f"{name}.mesh_index"] = sc.mesh_index
        frames[f"{name}.should_callback"] = sc.should_callback; frames[f"{name}.entities"] = sc.entities
        frames[f"{name}.n_meshes"] = np.array([len(sc.meshes)])
        for mi, msh in enumerate(sc.meshes):
            frames[f"{name}.mesh{mi}.pos"] = msh.positions; frames[f"{name}.mesh{mi}.nrm"] = msh.normals; frames[f"{name}.mesh{mi}.vid"] = msh.vertex_ids
        frames[f"{name}.pairs"] = pairs
        frames[f"{name}.hit_pair"] = np.concatenate(hit_pair) if hit_pair else np.zeros(0, np.uint32)
        frames[f"{name}.hit_ids"] = np.concatenate(hit_ids) if hit_ids else np.zeros((0, 2), np.uint32)
        frames[f"{name}.hit_seg"] = np.concatenate(hit_seg) if hit_seg else np.zeros((0, 7), np.float32)
        frames[f"{name}.summary"] = np.array(summ, np.uint64).reshape(-1, 7)
        frames[f"{name}.avg"] = np.array(avg, np.float32).reshape(-1, 6)
        print(name, "entries", sc.n_entries, "pairs", len(pairs), "totals", res["totals"])
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **frames)
    make_response(ref)
    make_primitives(ref)
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
