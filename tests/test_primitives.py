"""the engine's way into the tree constructor -- glTF primitives (points, optional normals, optional indices, draw mode)
through Triangle::CreateTriangleList (IMR/src/Geometry/Triangle.cpp:9-62,214-280).  CPU: the port against the unmodified reference on every
draw mode; GPU: the device-side assembly (imrcd_mesh_begin / add_primitive / end) against the oracle, bit for bit."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from helpers import f32_bits

MODES = {"points": 0, "lines": 1, "line_strip": 3, "triangles": 4, "triangle_strip": 5, "triangle_fan": 6}


def _indexed(mesh):
    """An indexed vertex buffer (points, per-vertex normals, u32 indices) out of a flat triangle mesh."""
    vid = mesh.vertex_ids.reshape(-1)
    nv = int(vid.max()) + 1
    pts = np.zeros((nv, 3), np.float32); nrm = np.zeros((nv, 3), np.float32)
    pts[vid] = mesh.positions.reshape(-1, 3); nrm[vid] = mesh.normals.reshape(-1, 3)
    return pts, nrm, vid.astype(np.uint32)


def _cases():
    rng = np.random.default_rng(12)
    pts, nrm, idx = _indexed(scenes.torus(24, 12))
    strip = np.arange(40, dtype=np.uint32) * 3 % len(pts)
    out = []
    for name, mode in MODES.items():
        for with_normals in (True, False):
            i = idx if mode == 4 else (strip if mode in (3, 5, 6) else idx[:61])
            out.append((f"{name}-{'n' if with_normals else 'fn'}", pts, nrm if with_normals else None, i, mode))
    out.append(("line_loop", pts, nrm, idx[:30], 2))                        # not handled by the reference's switch: no triangles
    out.append(("random-soup", rng.normal(size=(50, 3)).astype(np.float32), None, rng.integers(0, 50, 90).astype(np.uint32), 4))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_port_triangle_list_matches_reference(port, ref, case):
    _, pts, nrm, idx, mode = case
    a = port.triangle_list(pts, nrm, idx, mode); b = ref.triangle_list(pts, nrm, idx, mode)
    assert a[0].shape == b[0].shape
    assert np.array_equal(f32_bits(a[0]), f32_bits(b[0])) and np.array_equal(f32_bits(a[1]), f32_bits(b[1])) and np.array_equal(a[2], b[2])


def _nan_equal_bits(a, b):
    nan = np.isnan(b)
    return np.array_equal(np.isnan(a), nan) and np.array_equal(f32_bits(a)[~nan], f32_bits(b)[~nan])


def test_port_triangle_list_matches_golden(port):
    """The committed vectors (tests/golden/primitives.npz, outputs of the reference) pin the port where /root/reference is absent."""
    import golden_io
    z = golden_io.load("primitives")
    for name in z["names"].tolist():
        nrm = z[f"{name}.normals"] if f"{name}.normals" in z.files else None
        p, n, v = port.triangle_list(z[f"{name}.points"], nrm, z[f"{name}.indices"], int(z[f"{name}.mode"][0]))
        assert np.array_equal(f32_bits(p), f32_bits(z[f"{name}.pos"])) and _nan_equal_bits(n, z[f"{name}.nrm"]) and np.array_equal(v, z[f"{name}.vid"]), name


@pytest.mark.gpu
def test_device_assembly_matches_golden(gpu_ctx):
    from inmyroom_vulkan_b200.collision import OBBtree
    import golden_io
    z = golden_io.load("primitives")
    names = [n for n in z["names"].tolist() if len(z[f"{n}.pos"])]
    prims = [(z[f"{n}.points"], z[f"{n}.normals"] if f"{n}.normals" in z.files else None, z[f"{n}.indices"], int(z[f"{n}.mode"][0])) for n in names]
    flat = OBBtree.from_primitives(gpu_ctx, prims).export()
    back = np.argsort(flat.tri_orig)
    want_p = np.concatenate([z[f"{n}.pos"] for n in names]); want_n = np.concatenate([z[f"{n}.nrm"] for n in names]); want_v = np.concatenate([z[f"{n}.vid"] for n in names])
    assert np.array_equal(f32_bits(flat.tri_pos[back]), f32_bits(want_p)) and _nan_equal_bits(flat.tri_nrm[back], want_n) and np.array_equal(flat.tri_vid[back], want_v)


@pytest.mark.gpu
def test_device_assembly_matches_oracle(gpu_ctx, oracle):
    """A mesh of several primitives (all draw modes, vec3 and vec4 strides, with and without normals, with and without indices):
    the triangles the device tree holds are the oracle's list, in the order given, bit for bit."""
    from inmyroom_vulkan_b200.collision import OBBtree
    prims, want_p, want_n, want_v = [], [], [], []
    for k, (_, pts, nrm, idx, mode) in enumerate(_cases()):
        if k % 3 == 1:                                                       # the engine keeps vec4 points (PrimitivesOfMeshes.cpp:630-635)
            p4 = np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1); n4 = None if nrm is None else np.concatenate([nrm, np.zeros((len(nrm), 1), np.float32)], 1)
            prims.append((p4, n4, idx, mode))
        else:
            prims.append((pts, nrm, idx, mode))
        p, n, v = oracle.triangle_list(pts, nrm, idx, mode)
        want_p.append(p); want_n.append(n); want_v.append(v)
    pts, nrm, _ = _indexed(scenes.uv_sphere(10, 7))
    prims.append((pts[:60], nrm[:60], None, 4))                              # non-indexed: 0 .. n_points-1
    p, n, v = oracle.triangle_list(pts[:60], nrm[:60], np.arange(60, dtype=np.uint32), 4)
    want_p.append(p); want_n.append(n); want_v.append(v)
    want_p = np.concatenate(want_p); want_n = np.concatenate(want_n); want_v = np.concatenate(want_v)
    for build_mode in (0, 1):
        tree = OBBtree.from_primitives(gpu_ctx, prims, build_mode=build_mode)
        flat = tree.export()
        assert flat.tri_pos.shape[0] == len(want_p) and sorted(flat.tri_orig.tolist()) == list(range(len(want_p)))
        back = np.argsort(flat.tri_orig)                                     # leaf order -> the order given
        assert np.array_equal(f32_bits(flat.tri_pos[back]), f32_bits(want_p))
        got_n = flat.tri_nrm[back]                                          # degenerate triangles (points, lines) without normals: the face normal is
        nan = np.isnan(want_n)                                               # 0 * inf = NaN on both sides; only its payload is platform-defined
        assert np.array_equal(np.isnan(got_n), nan) and nan.any()
        assert np.array_equal(f32_bits(got_n)[~nan], f32_bits(want_n)[~nan])
        assert np.array_equal(flat.tri_vid[back], want_v)


@pytest.mark.gpu
def test_tree_from_primitives_equals_tree_from_triangles(gpu_ctx):
    """The same mesh through imrcd_mesh_create (flat triangles) and through the primitive recording: identical trees."""
    from inmyroom_vulkan_b200.collision import OBBtree
    mesh = scenes.torus(40, 20)
    pts, nrm, idx = _indexed(mesh)
    a = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids).export()
    b = OBBtree.from_primitives(gpu_ctx, [(pts, nrm, idx, 4)]).export()
    for f in ("boxes", "tri_pos", "tri_nrm"):
        assert np.array_equal(f32_bits(getattr(a, f)), f32_bits(getattr(b, f))), f
    for f in ("left", "right", "tri_off", "tri_cnt", "tri_vid", "tri_orig"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
