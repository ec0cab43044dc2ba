"""Synthetic meshes and scenes of the shapes BASELINE.json names (numpy only, deterministic).

All outputs are float32 / uint32 arrays in the layout the C ABI takes:
positions (n_tri, 9) = p0 p1 p2, normals (n_tri, 9), vertex_ids (n_tri, 3);
matrices (n, 16) column-major like glm::mat4 (ECStypes.h:149-156).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    positions: np.ndarray   # (n,9) f32
    normals: np.ndarray     # (n,9) f32
    vertex_ids: np.ndarray  # (n,3) u32
    name: str = ""

    @property
    def n_tri(self) -> int:
        return int(self.positions.shape[0])

    @property
    def radius(self) -> float:
        p = self.positions.reshape(-1, 3)
        return float(np.sqrt((p.astype(np.float64) ** 2).sum(1).max()))


def _indexed(verts, norms, faces, name) -> Mesh:
    verts = np.asarray(verts, np.float32); norms = np.asarray(norms, np.float32); faces = np.asarray(faces, np.uint32)
    pos = verts[faces].reshape(-1, 9).astype(np.float32)
    nrm = norms[faces].reshape(-1, 9).astype(np.float32)
    return Mesh(np.ascontiguousarray(pos), np.ascontiguousarray(nrm), np.ascontiguousarray(faces), name)


def torus(nu: int = 100, nv: int = 50, R: float = 1.0, r: float = 0.35) -> Mesh:
    """nu x nv torus grid -> 2*nu*nv triangles (100 x 50 = 10,000: the config-2 rigid mesh)."""
    u = np.arange(nu) * (2 * np.pi / nu); v = np.arange(nv) * (2 * np.pi / nv)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    x = (R + r * np.cos(vv)) * np.cos(uu); y = (R + r * np.cos(vv)) * np.sin(uu); z = r * np.sin(vv)
    verts = np.stack([x, y, z], -1).reshape(-1, 3)
    nx = np.cos(vv) * np.cos(uu); ny = np.cos(vv) * np.sin(uu); nz = np.sin(vv)
    norms = np.stack([nx, ny, nz], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = i * nv + j; b = ((i + 1) % nu) * nv + j; c = ((i + 1) % nu) * nv + (j + 1) % nv; d = i * nv + (j + 1) % nv
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return _indexed(verts, norms, faces, f"torus{nu}x{nv}")


def uv_sphere(n_lon: int = 66, n_lat: int = 65, radius: float = 1.0) -> Mesh:
    """UV sphere; 66 x 65 gives 2*66*64 = 8,448 triangles (same count as environment.gltf's spheres)."""
    lat = np.linspace(0, np.pi, n_lat + 1)[1:-1]
    lon = np.arange(n_lon) * (2 * np.pi / n_lon)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    ring = np.stack([np.sin(la) * np.cos(lo), np.cos(la), np.sin(la) * np.sin(lo)], -1).reshape(-1, 3)
    verts = np.concatenate([[[0, 1, 0]], ring, [[0, -1, 0]]]).astype(np.float64)
    nr = n_lat - 1
    faces = []
    for k in range(n_lon):
        faces.append([0, 1 + (k + 1) % n_lon, 1 + k])
    for rI in range(nr - 1):
        for k in range(n_lon):
            a = 1 + rI * n_lon + k; b = 1 + rI * n_lon + (k + 1) % n_lon
            c = 1 + (rI + 1) * n_lon + (k + 1) % n_lon; d = 1 + (rI + 1) * n_lon + k
            faces.append([a, b, c]); faces.append([a, c, d])
    south = 1 + nr * n_lon
    for k in range(n_lon):
        faces.append([south, 1 + (nr - 1) * n_lon + k, 1 + (nr - 1) * n_lon + (k + 1) % n_lon])
    return _indexed(verts * radius, verts, np.array(faces), f"sphere{n_lon}x{n_lat}")


def box_mesh(hx=1.0, hy=1.0, hz=1.0, sub: int = 1) -> Mesh:
    """Axis-aligned box with each face subdivided sub x sub (12*sub^2 triangles)."""
    verts, norms, faces = [], [], []
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            a1, a2 = (axis + 1) % 3, (axis + 2) % 3
            base = len(verts)
            for i in range(sub + 1):
                for j in range(sub + 1):
                    p = [0.0, 0.0, 0.0]
                    p[axis] = sgn; p[a1] = -1 + 2 * i / sub; p[a2] = -1 + 2 * j / sub
                    verts.append([p[0] * hx, p[1] * hy, p[2] * hz])
                    n = [0.0, 0.0, 0.0]; n[axis] = sgn; norms.append(n)
            for i in range(sub):
                for j in range(sub):
                    a = base + i * (sub + 1) + j; b = a + 1; c = a + (sub + 1) + 1; d = a + (sub + 1)
                    if sgn > 0:
                        faces += [[a, d, c], [a, c, b]]
                    else:
                        faces += [[a, b, c], [a, c, d]]
    return _indexed(verts, norms, faces, f"box{sub}")


def grid_sheet(nx: int, nz: int, sx: float, sz: float, bump: float = 0.0, seed: int = 0) -> Mesh:
    """Height-field sheet in the XZ plane, 2*nx*nz triangles (floors, curtains, displaced-grid build input)."""
    x = np.linspace(-sx, sx, nx + 1); z = np.linspace(-sz, sz, nz + 1)
    xx, zz = np.meshgrid(x, z, indexing="ij")
    yy = bump * (np.sin(3.1 * xx / max(sx, 1e-9)) * np.cos(2.3 * zz / max(sz, 1e-9)))
    verts = np.stack([xx, yy, zz], -1).reshape(-1, 3)
    # analytic-ish normals by finite differences
    dydx = np.gradient(yy, x, axis=0) if nx > 0 else np.zeros_like(yy)
    dydz = np.gradient(yy, z, axis=1) if nz > 0 else np.zeros_like(yy)
    nrm = np.stack([-dydx, np.ones_like(yy), -dydz], -1).reshape(-1, 3)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    i, j = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
    a = i * (nz + 1) + j; b = a + 1; c = a + (nz + 1) + 1; d = a + (nz + 1)
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return _indexed(verts, nrm, faces, f"sheet{nx}x{nz}")


def cylinder(n_seg: int = 32, n_h: int = 8, radius: float = 1.0, half_h: float = 1.0) -> Mesh:
    th = np.arange(n_seg) * (2 * np.pi / n_seg); h = np.linspace(-half_h, half_h, n_h + 1)
    tt, hh = np.meshgrid(th, h, indexing="ij")
    verts = np.stack([radius * np.cos(tt), hh, radius * np.sin(tt)], -1).reshape(-1, 3)
    norms = np.stack([np.cos(tt), np.zeros_like(tt), np.sin(tt)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n_seg), np.arange(n_h), indexing="ij")
    a = i * (n_h + 1) + j; b = ((i + 1) % n_seg) * (n_h + 1) + j; c = b + 1; d = a + 1
    faces = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return _indexed(verts, norms, faces, f"cyl{n_seg}x{n_h}")


# ---------------------------------------------------------------------------
# poses
# ---------------------------------------------------------------------------
def random_quaternions(rng: np.random.Generator, n: int) -> np.ndarray:
    """Uniform on SO(3) (Shoemake)."""
    u = rng.random((n, 3))
    q = np.stack([np.sqrt(1 - u[:, 0]) * np.sin(2 * np.pi * u[:, 1]), np.sqrt(1 - u[:, 0]) * np.cos(2 * np.pi * u[:, 1]),
                  np.sqrt(u[:, 0]) * np.sin(2 * np.pi * u[:, 2]), np.sqrt(u[:, 0]) * np.cos(2 * np.pi * u[:, 2])], -1)
    return q  # x y z w


def trs_matrices(t: np.ndarray, q: np.ndarray, s: np.ndarray) -> np.ndarray:
    """(n,16) column-major float32 TRS matrices."""
    t = np.asarray(t, np.float64).reshape(-1, 3); q = np.asarray(q, np.float64).reshape(-1, 4)
    s = np.asarray(s, np.float64)
    n = t.shape[0]
    if s.ndim == 0:
        s = np.full((n, 3), float(s))
    elif s.ndim == 1 and s.shape[0] == n:
        s = np.repeat(s[:, None], 3, 1)
    s = s.reshape(-1, 3)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    M = np.zeros((n, 4, 4))
    M[:, :3, :3] = R * s[:, None, :]
    M[:, :3, 3] = t
    M[:, 3, 3] = 1.0
    # column-major flatten: element (r,c) at 4*c + r
    return np.ascontiguousarray(M.transpose(0, 2, 1).reshape(n, 16).astype(np.float32))


@dataclass
class Scene:
    """A frame's worth of collision entries over a set of meshes."""
    meshes: list                 # list[Mesh]
    mesh_index: np.ndarray       # (n,) u32 : which mesh each entry uses
    matrices: np.ndarray         # (n,16) f32 current global matrices
    should_callback: np.ndarray  # (n,) u8
    entities: np.ndarray         # (n,) u32
    name: str = ""
    previous: np.ndarray | None = None
    meta: dict = field(default_factory=dict)

    @property
    def n_entries(self) -> int:
        return int(self.matrices.shape[0])


def scene_instances(mesh: Mesh, n: int, seed: int = 1234, neighbours: float = 8.0, scale: float = 1.0) -> Scene:
    """Config 2: n random-pose instances of one rigid mesh in a cube sized so that each body's
    bounding sphere overlaps about `neighbours` others."""
    rng = np.random.default_rng(seed)
    r = mesh.radius * scale
    # expected sphere-overlap partners = n * (4/3 pi (2r)^3) / L^3
    L = (n * (4.0 / 3.0) * np.pi * (2 * r) ** 3 / neighbours) ** (1.0 / 3.0)
    t = (rng.random((n, 3)) - 0.5) * L
    q = random_quaternions(rng, n)
    mats = trs_matrices(t, q, np.float64(scale))
    return Scene([mesh], np.zeros(n, np.uint32), mats, np.ones(n, np.uint8), np.arange(1, n + 1, dtype=np.uint32),
                 name=f"{n}x{mesh.name}", meta=dict(cube_side=float(L), seed=seed, neighbours=neighbours))


def atrium_static(seed: int = 7, detail: int = 1):
    """A Sponza-like static set: 163 mesh nodes (floor, walls, 2 storeys of columns and arches, curtains,
    small props) with roughly 262k triangles at detail=1, each node carrying Sponza's non-uniform scale
    (0.0399999991, 0.0400000028, 0.0400000028) and its (0.7071,0,0,0.7071) rotation."""
    rng = np.random.default_rng(seed)
    meshes, t_list, extra_rot = [], [], []
    S = 25.0  # model units per world unit (1/0.04)

    def add(m, pos):
        meshes.append(m); t_list.append(pos)

    d = max(1, int(detail))
    # big structural parts (long, thin, overlapping many bodies)
    add(grid_sheet(96 * d, 48 * d, 55 * S, 25 * S, bump=0.02 * S), [0, 0, 0])                # floor 9.2k
    add(grid_sheet(96 * d, 48 * d, 55 * S, 25 * S, bump=0.05 * S), [0, 0, 27 * S])           # upper floor
    add(grid_sheet(64 * d, 32 * d, 55 * S, 14 * S, bump=0.3 * S), [0, 0, 54 * S])            # roof
    for sx in (-1, 1):
        w = grid_sheet(80 * d, 40 * d, 27 * S, 25 * S, bump=0.1 * S)
        w = Mesh(w.positions.reshape(-1, 3)[:, [1, 0, 2]].reshape(-1, 9).copy(), w.normals.reshape(-1, 3)[:, [1, 0, 2]].reshape(-1, 9).copy(), w.vertex_ids, "wallx")
        add(w, [sx * 55 * S, 0, 27 * S])
    for sz in (-1, 1):
        w = grid_sheet(96 * d, 40 * d, 55 * S, 27 * S, bump=0.1 * S)
        w = Mesh(w.positions.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1, 9).copy(), w.normals.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1, 9).copy(), w.vertex_ids, "wallz")
        add(w, [0, sz * 25 * S, 27 * S])
    # columns (2 storeys x 2 rows x 14) = 56
    for storey in range(2):
        for row in (-1, 1):
            for k in range(14):
                add(cylinder(24 * d, 12 * d, 1.1 * S, 5.5 * S), [(-45 + 7 * k) * S, row * 11 * S, (6 + 27 * storey) * S])
    # arches (tori halves approximated by full tori) = 52
    for storey in range(2):
        for row in (-1, 1):
            for k in range(13):
                add(torus(40 * d, 16 * d, 3.2 * S, 0.5 * S), [(-41.5 + 7 * k) * S, row * 11 * S, (12 + 27 * storey) * S])
    # curtains (high-poly sheets) = 12
    for k in range(12):
        c = grid_sheet(70 * d, 70 * d, 3 * S, 5 * S, bump=0.6 * S, seed=k)
        c = Mesh(c.positions.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1, 9).copy(), c.normals.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1, 9).copy(), c.vertex_ids, "curtain")
        add(c, [(-38 + 7 * k) * S, (-1) ** k * 9.5 * S, 20 * S])
    # props: vases / spheres / boxes to reach 163 nodes
    while len(meshes) < 163:
        kind = len(meshes) % 3
        pos = [(rng.random() * 100 - 50) * S, (rng.random() * 40 - 20) * S, (1 + rng.random() * 3) * S]
        if kind == 0:
            add(uv_sphere(24 * d, 17 * d, 1.2 * S), pos)
        elif kind == 1:
            add(box_mesh(1.0 * S, 1.5 * S, 0.8 * S, sub=4 * d), pos)
        else:
            add(torus(30 * d, 14 * d, 1.0 * S, 0.35 * S), pos)
    n = len(meshes)
    # node transform: rotation (0.7071,0,0,0.7071) [x y z w], Sponza's non-uniform scale, translation 0:
    # geometry above is authored in the pre-rotation frame (z up) so the world ends up y-up like Sponza.
    q = np.tile(np.array([[-0.70710678, 0.0, 0.0, 0.70710678]]), (n, 1))
    s = np.tile(np.array([[0.0399999991, 0.0400000028, 0.0400000028]]), (n, 1))
    # translations are baked into the mesh vertices (Sponza's nodes have none)
    out = []
    for m, t in zip(meshes, t_list):
        p = m.positions.reshape(-1, 3).astype(np.float64) + np.asarray(t, np.float64)
        out.append(Mesh(np.ascontiguousarray(p.astype(np.float32).reshape(-1, 9)), m.normals, m.vertex_ids, m.name))
    mats = trs_matrices(np.zeros((n, 3)), q, s)
    return out, mats


def scene_static_vs_bodies(body: Mesh, n_bodies: int, seed: int = 2026, body_scale=(0.2, 0.5), detail: int = 1,
                           static=None) -> Scene:
    """Config 3: a 163-node static scene (should_callback=1) against n dynamic bodies (should_callback=0),
    so only static-dynamic pairs survive the broad phase (SweepAndPrune.cpp:60)."""
    rng = np.random.default_rng(seed)
    st_meshes, st_mats = static if static is not None else atrium_static(detail=detail)
    ns = len(st_meshes)
    lo = np.array([-54.0, 0.3, -24.0]); hi = np.array([54.0, 50.0, 24.0])
    t = lo + rng.random((n_bodies, 3)) * (hi - lo)
    q = random_quaternions(rng, n_bodies)
    sc = body_scale[0] + rng.random(n_bodies) * (body_scale[1] - body_scale[0])
    dyn = trs_matrices(t, q, sc)
    mats = np.concatenate([st_mats, dyn]).astype(np.float32)
    mesh_index = np.concatenate([np.arange(ns, dtype=np.uint32), np.full(n_bodies, ns, np.uint32)])
    cb = np.concatenate([np.ones(ns, np.uint8), np.zeros(n_bodies, np.uint8)])
    ents = np.arange(1, ns + n_bodies + 1, dtype=np.uint32)
    return Scene(st_meshes + [body], mesh_index, np.ascontiguousarray(mats), cb, ents,
                 name=f"static{ns}+{n_bodies}x{body.name}", meta=dict(seed=seed))


# ---- BASELINE config 5: a skinned / morphed character (a ring of joints along a torus) --------------------------------------------------
@dataclass
class Character:
    mesh: Mesh                  # the bind pose as a triangle mesh (vertex_ids index `vertices`)
    vertices: np.ndarray        # (n_vertices, T + 1, 4) f32: base vertex (w = 1), then its morph targets (deltas, w = 0): the engine's layout
    joints: np.ndarray          # (n_vertices, G, 4) u16
    weights: np.ndarray         # (n_vertices, G, 4) f32, rows sum to 1
    inverse_bind: np.ndarray    # (J, 16) f32 column-major
    joint_centres: np.ndarray   # (J, 3) f64

    @property
    def n_joints(self) -> int:
        return int(self.inverse_bind.shape[0])

    def pose(self, phase: float, amplitude: float = 0.35):
        """Joint matrices (J, 16) and morph weights (T,) of one frame of a seeded animation: every joint turns about the ring's tangent by
        an angle that travels round the ring (modelMatrices[j].positionMatrix of the shader: bind translation x rotation)."""
        J = self.n_joints
        ang = amplitude * np.sin(phase + 4.0 * np.pi * np.arange(J) / J)
        t = 2 * np.pi * np.arange(J) / J
        tangent = np.stack([-np.sin(t), np.cos(t), np.zeros(J)], 1)
        q = np.concatenate([tangent * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], 1)
        mats = trs_matrices(self.joint_centres, q, np.ones((J, 3)))
        T = self.vertices.shape[1] - 1
        mw = (0.5 + 0.5 * np.sin(phase * 1.7 + np.arange(T))).astype(np.float32)
        return mats, mw


def character(nu: int = 142, nv: int = 71, n_joints: int = 64, n_targets: int = 2, R: float = 1.0, r: float = 0.35) -> Character:
    """A nu x nv torus (142 x 71 = 20,164 triangles, 10,082 vertices) skinned to a ring of joints: 4 joints per vertex (the two nearest on
    either side, smooth weights), `n_targets` morph targets (radial bulges)."""
    mesh = torus(nu, nv, R, r)
    u = np.arange(nu) * (2 * np.pi / nu); v = np.arange(nv) * (2 * np.pi / nv)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    verts = np.stack([(R + r * np.cos(vv)) * np.cos(uu), (R + r * np.cos(vv)) * np.sin(uu), r * np.sin(vv)], -1).reshape(-1, 3)
    n = verts.shape[0]
    V = np.zeros((n, n_targets + 1, 4), np.float32)
    V[:, 0, :3] = verts; V[:, 0, 3] = 1.0
    rad = np.stack([np.cos(vv) * np.cos(uu), np.cos(vv) * np.sin(uu), np.sin(vv)], -1).reshape(-1, 3)
    for k in range(n_targets):
        V[:, k + 1, :3] = (0.08 * np.cos((k + 2) * uu + k).reshape(-1, 1) * rad)
    pos = uu.reshape(-1) / (2 * np.pi) * n_joints                     # position along the ring in joint units
    j0 = np.floor(pos).astype(np.int64)
    f = pos - j0
    idx = np.stack([(j0 - 1) % n_joints, j0 % n_joints, (j0 + 1) % n_joints, (j0 + 2) % n_joints], 1)
    w = np.stack([(1 - f) ** 3, 3 * f ** 3 - 6 * f ** 2 + 4, -3 * f ** 3 + 3 * f ** 2 + 3 * f + 1, f ** 3], 1) / 6.0      # cubic B-spline: smooth, sums to 1
    t = 2 * np.pi * np.arange(n_joints) / n_joints
    centres = np.stack([R * np.cos(t), R * np.sin(t), np.zeros(n_joints)], 1)
    inv_bind = trs_matrices(-centres, np.tile([0.0, 0.0, 0.0, 1.0], (n_joints, 1)), np.ones((n_joints, 3)))
    return Character(mesh, V, idx.astype(np.uint16).reshape(n, 1, 4), w.astype(np.float32).reshape(n, 1, 4), inv_bind, centres)
