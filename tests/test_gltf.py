""".gltf / .glb -> collision trees without the engine's extractor (imrcd_gltf_*, include/imrcd.h).

CPU: the library's file reader against the reference's own reader (tinygltf) + the engine's primitive order and extraction
(oracle/ref_gltf_shim.cpp), on files in all three containers; against the committed golden (outputs of the reference) where /root/reference is
absent; error behaviour.  GPU: imrcd_gltf_load's trees against the unmodified reference building from the same file, bit for bit."""
import os

import numpy as np
import pytest

import gltf_writer
import golden_io
from helpers import f32_bits
from inmyroom_vulkan_b200.gltf import GltfFile

CONTAINERS = {"bin": "scene.gltf", "uri": "scene.gltf", "glb": "scene.glb"}


@pytest.fixture(scope="module")
def refgltf():
    from oracle import bind
    if os.path.isdir("/root/reference/tinygltf"):
        bind.build("ref")
    if not os.path.exists(bind.REF_GLTF_SO):
        pytest.skip("oracle/_ref/libimr_ref_gltf.so not available")
    return bind.RefGltf


def _same_primitives(got, want):
    assert len(got) == len(want)
    for (gp, gn, gi, gm, gs), (wp, wn, wi, wm, ws) in zip(got, want):
        assert (gm, gs) == (wm, ws)
        assert np.array_equal(f32_bits(gp), f32_bits(wp))
        assert (gn is None) == (wn is None) and (gn is None or np.array_equal(f32_bits(gn), f32_bits(wn)))
        assert (gi is None) == (wi is None) and (gi is None or np.array_equal(gi, wi))


@pytest.mark.parametrize("container", list(CONTAINERS))
def test_reader_matches_the_reference_reader(tmp_path, refgltf, container):
    path = gltf_writer.write(str(tmp_path / CONTAINERS[container]), gltf_writer.sample_meshes(), container)
    want = refgltf(path)
    with GltfFile(path) as g:
        assert g.n_meshes == want.n_meshes == len(gltf_writer.sample_meshes())
        for m in range(g.n_meshes):
            _same_primitives(g.primitives(m), want.primitives(m))
        order = [p[4] for p in g.primitives(0)]
    assert order == [3, 2, 1, 0]                       # what "triangles first" really does to an all-triangles mesh (MeshesOfNodes.cpp:41-43)


@pytest.mark.parametrize("name", ["Cube/Cube.gltf", "box01.glb", "BoundsChecking/integer-out-of-bounds.gltf", "BoundsChecking/invalid-buffer-index.gltf",
                                  "BoundsChecking/invalid-buffer-view-index.gltf", "BoundsChecking/invalid-primitive-indices.gltf", "regression/unassigned-skeleton.gltf"])
def test_reader_on_the_reference_readers_sample_models(refgltf, name):
    path = os.path.join("/root/reference/tinygltf/models", name)
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    # the reference's reader has undefined behaviour on some of its own bounds-checking models (an out-of-range buffer index is read
    # before it is checked: sometimes an error, sometimes a crash), so it is asked in a child process first; a crash counts as a refusal
    import subprocess
    import sys
    probe = subprocess.run([sys.executable, "-c", "import sys; from oracle import bind\ntry:\n    bind.RefGltf(sys.argv[1])\nexcept ValueError:\n    sys.exit(3)", path],
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), capture_output=True)
    if probe.returncode != 0:
        with pytest.raises(ValueError):
            GltfFile(path)
        return
    want = refgltf(path)
    with GltfFile(path) as g:
        assert g.n_meshes == want.n_meshes
        for m in range(g.n_meshes):
            _same_primitives(g.primitives(m), want.primitives(m))


def _triangles_of(checker, prims):
    ps, ns, vs = [], [], []
    for pts, nrm, idx, mode, _ in prims:
        i = np.arange(len(pts), dtype=np.uint32) if idx is None else idx
        p, n, v = checker.triangle_list(pts, nrm, i, mode)
        ps.append(p); ns.append(n); vs.append(v)
    if not ps:
        return np.zeros((0, 9), np.float32), np.zeros((0, 9), np.float32), np.zeros((0, 3), np.uint32)
    return np.concatenate(ps), np.concatenate(ns), np.concatenate(vs)


def _nan_equal_bits(a, b):
    nan = np.isnan(b)
    return np.array_equal(np.isnan(a), nan) and np.array_equal(f32_bits(a)[~nan], f32_bits(b)[~nan])


def test_reader_and_port_match_golden(port):
    """tests/golden/gltf_scene.glb + gltf.npz (made by make_golden.py gltf: the reference's reader, order, CreateTriangleList): pins the
    reader where /root/reference is absent."""
    z = golden_io.load("gltf")
    with GltfFile(os.path.join(golden_io.GOLDEN, "gltf_scene.glb")) as g:
        assert g.n_meshes == int(z["n_meshes"][0])
        for m in range(g.n_meshes):
            p, n, v = _triangles_of(port, g.primitives(m))
            assert np.array_equal(f32_bits(p), f32_bits(z[f"m{m}.pos"])) and _nan_equal_bits(n, z[f"m{m}.nrm"]) and np.array_equal(v, z[f"m{m}.vid"]), m


def test_strided_views_and_u8_indices(tmp_path):
    """Product-only superset (the reference reads tightly packed accessors and asserts on u8 indices): against the writer's own arrays."""
    pts, nrm, idx = gltf_writer._indexed(gltf_writer.scenes.box_mesh(1, 1, 2, sub=2))
    assert len(pts) < 256
    path = gltf_writer.write(str(tmp_path / "s.gltf"), [[dict(points=pts, normals=nrm, indices=idx.astype(np.uint8), mode=4, interleave=True)]], "uri")
    with GltfFile(path) as g:
        (p, n, i, mode, _), = g.primitives(0)
    flip = np.array([1, -1, -1], np.float32)
    assert np.array_equal(f32_bits(p), f32_bits(pts * flip)) and np.array_equal(f32_bits(n), f32_bits(nrm * flip)) and np.array_equal(i, idx) and mode == 4


def test_indexed_primitive_that_draws_nothing(tmp_path):
    """An accessor of zero indices: no triangles.  (The reference dereferences max_element of the empty list, Triangle.cpp:244 -- a crash,
    not a behaviour to reproduce.)"""
    pts, nrm, _ = gltf_writer._indexed(gltf_writer.scenes.box_mesh(1, 1, 1, sub=1))
    path = gltf_writer.write(str(tmp_path / "e.glb"), [[dict(points=pts, normals=nrm, indices=np.zeros(0, np.uint16), mode=4)]], "glb")
    with GltfFile(path) as g:
        (p, n, i, mode, _), = g.primitives(0)
    assert len(p) == len(pts) and i is not None and len(i) == 0


@pytest.mark.parametrize("damage", ["json", "truncated_bin", "missing_bin", "sparse", "bad_accessor", "index_range", "glb_header", "int_position"])
def test_malformed_files_are_refused(tmp_path, damage):
    import json
    pts, nrm, idx = gltf_writer._indexed(gltf_writer.scenes.box_mesh(1, 1, 1, sub=1))
    path = gltf_writer.write(str(tmp_path / "m.gltf"), [[dict(points=pts, normals=nrm, indices=idx.astype(np.uint16), mode=4)]], "glb" if damage == "glb_header" else "bin")
    binp = str(tmp_path / "m data.bin")
    if damage == "glb_header":
        raw = bytearray(open(path, "rb").read()); raw[8:12] = (len(raw) + 64).to_bytes(4, "little"); open(path, "wb").write(raw)
    elif damage == "json":
        open(path, "w").write(open(path).read()[:-20])
    elif damage == "truncated_bin":
        open(binp, "wb").write(open(binp, "rb").read()[:40])
    elif damage == "missing_bin":
        os.remove(binp)
    else:
        doc = json.load(open(path))
        if damage == "sparse":
            doc["accessors"][0]["sparse"] = {"count": 1, "indices": {"bufferView": 0, "componentType": 5123}, "values": {"bufferView": 0}}
        elif damage == "bad_accessor":
            doc["meshes"][0]["primitives"][0]["attributes"]["POSITION"] = 99
        elif damage == "index_range":
            doc["accessors"][0]["count"] = 3            # fewer points than the indices address
        elif damage == "int_position":
            doc["accessors"][0]["componentType"] = 5125
        json.dump(doc, open(path, "w"))
    with pytest.raises(ValueError) as e:
        GltfFile(path)
    assert str(e.value)


# ------------------------------------------------------------------------------------------------ GPU
def _assert_same_tree(flat, gold, in_pos, in_vid):
    """tri_orig: the checker recovers it by matching triangle contents (the reference does not track it), which is ambiguous where a mesh
    draws the same triangle twice (sample mesh 1 does); it is checked through what it points at instead."""
    orig = np.asarray(flat.tri_orig)
    assert sorted(orig.tolist()) == list(range(len(in_pos)))
    assert np.array_equal(f32_bits(in_pos[orig]), f32_bits(flat.tri_pos)) and np.array_equal(in_vid[orig], flat.tri_vid)
    for f in golden_io.TREE_FIELDS[:-1]:
        g = np.asarray(getattr(gold, f)); o = np.asarray(getattr(flat, f))
        assert o.shape == g.shape, f"{f}: shape {o.shape} vs {g.shape}"
        if g.dtype == np.float32:
            assert _nan_equal_bits(o, g), f"{f} differs in bits"
        else:
            assert np.array_equal(o, g), f


@pytest.mark.gpu
@pytest.mark.parametrize("container", list(CONTAINERS))
def test_trees_from_file_match_the_reference(tmp_path, gpu_ctx, oracle, container):
    """imrcd_gltf_load in IMRCD_BUILD_REFERENCE mode: one tree per mesh, each bit-identical to the reference's OBBtree built from the
    triangles the reference's loader makes of the same file (golden primitives where the reference's reader is absent)."""
    from inmyroom_vulkan_b200.collision import IMRCD_BUILD_REFERENCE
    from inmyroom_vulkan_b200.gltf import load_gltf
    from oracle import bind
    path = gltf_writer.write(str(tmp_path / CONTAINERS[container]), gltf_writer.sample_meshes(), container)
    want = bind.RefGltf(path) if os.path.exists(bind.REF_GLTF_SO) else GltfFile(path)
    trees = load_gltf(gpu_ctx, path, build_mode=IMRCD_BUILD_REFERENCE)
    assert len(trees) == want.n_meshes
    for m, tree in enumerate(trees):
        p, n, v = _triangles_of(oracle, want.primitives(m))
        flat = tree.export()
        assert flat.tri_pos.shape[0] == len(p)
        if len(p) == 0:
            continue
        _assert_same_tree(flat, oracle.tree_build(p, n, v).flat, p, v)


@pytest.mark.gpu
def test_golden_file_morton_trees_hold_the_reference_triangles(gpu_ctx):
    from inmyroom_vulkan_b200.gltf import load_gltf
    z = golden_io.load("gltf")
    trees = load_gltf(gpu_ctx, os.path.join(golden_io.GOLDEN, "gltf_scene.glb"))
    assert len(trees) == int(z["n_meshes"][0])
    for m, tree in enumerate(trees):
        flat = tree.export()
        back = np.argsort(flat.tri_orig)
        assert np.array_equal(f32_bits(flat.tri_pos[back]), f32_bits(z[f"m{m}.pos"])) and _nan_equal_bits(flat.tri_nrm[back], z[f"m{m}.nrm"])
        assert np.array_equal(flat.tri_vid[back], z[f"m{m}.vid"])


@pytest.mark.gpu
def test_loaded_meshes_collide_like_the_flat_path(tmp_path, gpu_ctx, port):
    """Trees that came from a file go through a frame exactly like trees that came from imrcd_mesh_create."""
    from inmyroom_vulkan_b200 import scenes
    from inmyroom_vulkan_b200.collision import CollisionDetection
    from inmyroom_vulkan_b200.gltf import load_gltf
    from helpers import compare_frame, gpu_frame, oracle_frame
    mesh = scenes.torus(24, 12)
    pts, nrm, idx = gltf_writer._indexed(mesh)
    flip = np.array([1, -1, -1], np.float32)                       # the loader flips y and z; write the file in glTF axes
    path = gltf_writer.write(str(tmp_path / "t.glb"), [[dict(points=pts * flip, normals=nrm * flip, indices=idx, mode=4)]], "glb")
    tree, = load_gltf(gpu_ctx, path)
    scene = scenes.scene_instances(mesh, 48, seed=5, neighbours=6.0)
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, scene, [tree])
    o_tree = port.tree_import(tree.export())
    ores = oracle_frame(port, scene, [o_tree], port=port)
    compare_frame(ores, st, bp, ep, hits, rel_of=lambda k: port.pair_matrix(scene.matrices[k[0]], scene.matrices[k[1]]))
    assert st["n_hits"] > 0


def test_mutated_files_never_crash_the_reader(tmp_path):
    """Two hundred random byte / field mutations of valid files: every one is either read or refused with a message (the reader was also
    run over 3000 such files under ASan + UBSan while it was written)."""
    import json
    import random
    rng = random.Random(11)
    ms = gltf_writer.sample_meshes()
    seeds = [gltf_writer.write(str(tmp_path / "a.glb"), ms[:3], "glb"), gltf_writer.write(str(tmp_path / "b.gltf"), ms[:2], "uri")]
    n_ok = n_refused = 0
    for k in range(200):
        src = rng.choice(seeds); raw = bytearray(open(src, "rb").read())
        if src.endswith(".glb") or rng.random() < 0.5:
            for _ in range(rng.choice([1, 1, 2, 4, 16])):
                i = rng.randrange(min(len(raw), 6000 if rng.random() < 0.7 else len(raw)))
                raw[i] = rng.randrange(256) if rng.random() < 0.5 else raw[i] ^ (1 << rng.randrange(8))
            if rng.random() < 0.2:
                raw = raw[:rng.randrange(len(raw))]
        else:
            doc = json.loads(raw.decode())
            for _ in range(rng.choice([1, 2, 3])):
                sec = rng.choice(["accessors", "bufferViews", "buffers"])
                obj = rng.choice(doc[sec])
                keys = [key for key, v in obj.items() if isinstance(v, (int, float))]
                if keys:
                    obj[rng.choice(keys)] = rng.choice([-1, 0, 1, 2, 3, 7, 255, 65536, 2**31 - 1, 2**32, 2**40, 1e300, 1.5, 5120, 5121, 5123, 5125, 5126, 5130])
            raw = json.dumps(doc).encode()
        path = str(tmp_path / ("m" + os.path.splitext(src)[1]))
        open(path, "wb").write(raw)
        try:
            with GltfFile(path) as g:
                for m in range(g.n_meshes):
                    for p, n, i, mode, _ in g.primitives(m):
                        assert i is None or len(i) == 0 or int(i.max()) < len(p)
            n_ok += 1
        except ValueError as e:
            assert str(e)
            n_refused += 1
    assert n_ok > 20 and n_refused > 20
