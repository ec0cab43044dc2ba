"""oracle/bind.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

* ``RefOracle``  -> oracle/_ref/libimr_ref.so : the UNMODIFIED reference sources compiled
  in place from /root/reference (oracle/Makefile ``ref``) behind oracle/ref_shim.cpp.
* ``PortOracle`` -> oracle/liboracle_port.so  : the plain-C restatement (oracle/imr_oracle.c).

Both expose the same Python surface so tests can run either against the CUDA path.
Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libimr_ref.so")
REF_GLTF_SO = os.path.join(HERE, "_ref", "libimr_ref_gltf.so")
PORT_SO = os.path.join(HERE, "liboracle_port.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(which: str = "all") -> None:
    """Compile the checkers (``port``, ``ref`` or ``all``).  ``ref`` is a no-op when
    /root/reference is absent (the GPU box): the prebuilt _ref/ library travels with the repo."""
    subprocess.run(["make", "-s", "-C", HERE, which], check=True)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class FlatTree:
    """Flat interchange form of an OBB tree (see imr_oracle.h)."""
    boxes: np.ndarray      # (nv,12) f32
    left: np.ndarray       # (nv,) i32
    right: np.ndarray      # (nv,) i32
    tri_off: np.ndarray    # (nv,) u32
    tri_cnt: np.ndarray    # (nv,) u32
    tri_pos: np.ndarray    # (n,9) f32, leaf order
    tri_nrm: np.ndarray    # (n,9) f32
    tri_vid: np.ndarray    # (n,3) u32
    tri_orig: np.ndarray   # (n,) u32

    @property
    def nv(self):
        return self.boxes.shape[0]

    @property
    def n_tri(self):
        return self.tri_pos.shape[0]


@dataclass
class PairResult:
    hit_ids: np.ndarray    # (h,2) u32 original triangle indices (first tree, second tree)
    hit_seg: np.ndarray    # (h,7) f32 source(3) target(3) weight
    n_combos: int
    n_tri_tests: int
    n_hits: int
    n_coplanar: int
    rays_first: int
    rays_second: int
    colliding: bool
    avg: np.ndarray        # (6,) f32
    seconds: tuple = (0.0, 0.0)


class _Base:
    prefix = ""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self._sig()

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- common predicate surface -------------------------------------
    def sat(self, a12, b12, m16=None) -> int:
        a12 = _c(a12, np.float32); b12 = _c(b12, np.float32)
        m = None if m16 is None else _c(m16, np.float32)
        f = self._fn("sat"); f.restype = C.c_int
        f.argtypes = [C.c_void_p] * 3
        return f(a12.ctypes.data, b12.ctypes.data, _opt(m))

    def surface(self, a12, m16=None) -> np.float32:
        a12 = _c(a12, np.float32)
        m = None if m16 is None else _c(m16, np.float32)
        f = self._fn("surface"); f.restype = C.c_float
        f.argtypes = [C.c_void_p] * 2
        return np.float32(f(a12.ctypes.data, _opt(m)))

    def box_transform(self, a12, m16):
        a12 = _c(a12, np.float32); m16 = _c(m16, np.float32)
        out = np.empty(12, np.float32)
        f = self._fn("box_transform"); f.restype = None
        f.argtypes = [C.c_void_p] * 3
        f(a12.ctypes.data, m16.ctypes.data, out.ctypes.data)
        return out

    def tri_tri(self, a, b, m16=None):
        a = _c(a, np.float32).reshape(-1, 9); b = _c(b, np.float32).reshape(-1, 9)
        n = a.shape[0]
        m = None if m16 is None else _c(m16, np.float32)
        flags = np.zeros(n, np.uint8)
        seg = np.zeros((n, 6), np.float32)
        f = self._fn("tri_tri"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        f(a.ctypes.data, b.ctypes.data, _opt(m), n, flags.ctypes.data, seg.ctypes.data)
        return flags, seg

    def pair_matrix(self, a16, b16):
        a16 = _c(a16, np.float32); b16 = _c(b16, np.float32)
        out = np.empty(16, np.float32)
        f = self._fn("pair_matrix"); f.restype = None
        f.argtypes = [C.c_void_p] * 3
        f(a16.ctypes.data, b16.ctypes.data, out.ctypes.data)
        return out

    def obb_from_points(self, pts):
        pts = _c(pts, np.float32).reshape(-1, 3)
        out = np.empty(12, np.float32)
        f = self._fn("obb_from_points"); f.restype = None
        f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        f(pts.ctypes.data, pts.shape[0], out.ctypes.data)
        return out

    def sweep_axes(self):
        out = np.empty(9, np.float32)
        f = self._fn("sweep_axes"); f.restype = None
        f.argtypes = [C.c_void_p]
        f(out.ctypes.data)
        return out.reshape(3, 3)

    # ---- response rays (Ray.cpp, ShootUncollideRays.cpp, CollisionDetection.cpp:80-103) ----
    def ray_tree(self, tree, m16, origin, direction):
        """Ray::IntersectOBBtree: (doIntersect, itBackfaces, distance, bary(2), leaf-order triangle index)."""
        m16 = _c(m16, np.float32); o = _c(origin, np.float32); d = _c(direction, np.float32)
        out = np.zeros(3, np.float32); tri = C.c_uint32(0); back = C.c_int(0)
        f = self._fn("ray_tree"); f.restype = C.c_int
        f.argtypes = [C.c_void_p] * 7
        hit = f(tree.h, m16.ctypes.data, o.ctypes.data, d.ctypes.data, out.ctypes.data, C.byref(tri), C.byref(back))
        return bool(hit), bool(back.value), out[0], out[1:3].copy(), int(tri.value)

    def pair_delta(self, ta, ma, pa, tb, mb, pb):
        """One ordered pair through CollisionDetection.cpp:44-103: (colliding, deltaVector first, deltaVector second)."""
        ma = _c(ma, np.float32); mb = _c(mb, np.float32); pa = _c(pa, np.float32); pb = _c(pb, np.float32)
        d = np.zeros(6, np.float32); extra = np.zeros(2, np.uint64)
        f = self._fn("pair_delta"); f.restype = C.c_int
        f.argtypes = [C.c_void_p] * 8
        col = f(ta.h, ma.ctypes.data, pa.ctypes.data, tb.h, mb.ctypes.data, pb.ctypes.data, d.ctypes.data, extra.ctypes.data)
        return bool(col), d[:3].copy(), d[3:].copy()

    def triangle_list(self, points, normals, indices, mode):
        """Triangle::CreateTriangleList on one primitive: (positions (n,9), normals (n,9), vertex ids (n,3))."""
        pts = _c(points, np.float32).reshape(-1, 3); idx = _c(indices, np.uint32).reshape(-1)
        nrm = None if normals is None else _c(normals, np.float32).reshape(-1, 3)
        f = self._fn("triangle_list"); f.restype = C.c_uint64
        if self.kind == "reference":
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
            call = lambda p9, n9, v3: f(pts.ctypes.data, len(pts), _opt(nrm), idx.ctypes.data, len(idx), int(mode), p9, n9, v3)
        else:
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
            call = lambda p9, n9, v3: f(pts.ctypes.data, _opt(nrm), idx.ctypes.data, len(idx), int(mode), p9, n9, v3)
        n = int(call(None, None, None))
        pos = np.zeros((n, 9), np.float32); nr = np.zeros((n, 9), np.float32); vid = np.zeros((n, 3), np.uint32)
        if n:
            call(pos.ctypes.data, nr.ctypes.data, vid.ctypes.data)
        return pos, nr, vid

    # ---- trees ---------------------------------------------------------
    def tree_build(self, pos, nrm=None, vid=None):
        pos = _c(pos, np.float32).reshape(-1, 9)
        nrm = None if nrm is None else _c(nrm, np.float32).reshape(-1, 9)
        vid = None if vid is None else _c(vid, np.uint32).reshape(-1, 3)
        f = self._fn("tree_create" if self.prefix == "imr_ref_" else "tree_build")
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        h = f(pos.ctypes.data, _opt(nrm), _opt(vid), pos.shape[0])
        return Tree(self, h)

    def _tree_free(self, h):
        f = self._fn("tree_destroy" if self.prefix == "imr_ref_" else "tree_free")
        f.restype = None; f.argtypes = [C.c_void_p]
        f(h)

    def _tree_counts(self, h):
        fv = self._fn("tree_vertex_count"); fv.restype = C.c_uint64; fv.argtypes = [C.c_void_p]
        ft = self._fn("tree_tri_count"); ft.restype = C.c_uint64; ft.argtypes = [C.c_void_p]
        return int(fv(h)), int(ft(h))

    def _tree_flatten(self, h) -> FlatTree:
        nv, n = self._tree_counts(h)
        ft = FlatTree(np.empty((nv, 12), np.float32), np.empty(nv, np.int32), np.empty(nv, np.int32),
                      np.empty(nv, np.uint32), np.empty(nv, np.uint32), np.empty((n, 9), np.float32),
                      np.empty((n, 9), np.float32), np.empty((n, 3), np.uint32), np.empty(n, np.uint32))
        f = self._fn("tree_flatten"); f.restype = None
        f.argtypes = [C.c_void_p] * 10
        f(h, ft.boxes.ctypes.data, ft.left.ctypes.data, ft.right.ctypes.data, ft.tri_off.ctypes.data,
          ft.tri_cnt.ctypes.data, ft.tri_pos.ctypes.data, ft.tri_nrm.ctypes.data, ft.tri_vid.ctypes.data,
          ft.tri_orig.ctypes.data)
        return ft


class Tree:
    def __init__(self, owner: _Base, handle):
        self.owner = owner
        self.h = handle
        self._flat = None

    def __del__(self):
        try:
            if self.h:
                self.owner._tree_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def flat(self) -> FlatTree:
        if self._flat is None:
            self._flat = self.owner._tree_flatten(self.h)
        return self._flat

    @property
    def root_box(self):
        return self.flat.boxes[0]


class RefOracle(_Base):
    """The unmodified reference (kind = "reference")."""
    prefix = "imr_ref_"
    kind = "reference"

    def __init__(self, path=REF_SO):
        super().__init__(path)

    def _sig(self):
        pass

    def broad(self, mats, trees, should_cb):
        """SweepAndPrune::ExecuteSweepAndPrune on <= 65534 entries.  Returns (pairs (k,2) u32, seconds)."""
        mats = _c(mats, np.float32).reshape(-1, 16)
        n = mats.shape[0]
        cb = _c(should_cb, np.uint8)
        handles = (C.c_void_p * n)(*[t.h for t in trees])
        secs = C.c_double(0)
        cap = max(1024, 64 * n)
        f = self.lib.imr_ref_broad; f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        while True:
            pairs = np.empty((cap, 2), np.uint32)
            k = f(mats.ctypes.data, handles, cb.ctypes.data, n, pairs.ctypes.data, cap, C.byref(secs))
            if k == 2 ** 64 - 1:
                raise ValueError("reference broad phase is limited to 65534 entries (Entity is uint16_t)")
            if k <= cap:
                return pairs[:k].copy(), secs.value
            cap = int(k)

    def mid(self, ta: Tree, ma, tb: Tree, mb):
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        f = self.lib.imr_ref_mid; f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        cap = 1 << 16
        secs = C.c_double(0)
        while True:
            combos = np.empty((cap, 4), np.uint32)
            k = f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, combos.ctypes.data, cap, C.byref(secs))
            if k <= cap:
                return combos[:k].copy(), secs.value
            cap = int(k)

    def mid_stats(self, ta, ma, tb, mb):
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        out = np.zeros(5, np.uint64)
        f = self.lib.imr_ref_mid_stats; f.restype = None
        f.argtypes = [C.c_void_p] * 5
        f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, out.ctypes.data)
        return dict(visits=int(out[0]), passes=int(out[1]), combos=int(out[2]), tri_tests=int(out[3]), max_depth=int(out[4]))

    def pair(self, ta, ma, tb, mb) -> PairResult:
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        f = self.lib.imr_ref_pair; f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                      C.c_void_p, C.c_void_p, C.c_void_p]
        cap = 1 << 12
        while True:
            ids = np.empty((cap, 2), np.uint32); seg = np.empty((cap, 7), np.float32)
            summ = np.zeros(7, np.uint64); avg = np.zeros(6, np.float32); secs = np.zeros(2, np.float64)
            f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, ids.ctypes.data, seg.ctypes.data, cap,
              summ.ctypes.data, avg.ctypes.data, secs.ctypes.data)
            h = int(summ[2])
            if h <= cap:
                return PairResult(ids[:h].copy(), seg[:h].copy(), int(summ[0]), int(summ[1]), h, int(summ[3]),
                                  int(summ[4]), int(summ[5]), bool(summ[6]), avg, (float(secs[0]), float(secs[1])))
            cap = h

    def shoot(self, ta, ma, tb, mb, rays_first, rays_second):
        """ShootUncollideRays::ExecuteShootUncollideRays on explicit ray lists (n x 6, first's model space): world-space delta."""
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        r1 = _c(rays_first, np.float32).reshape(-1, 6); r2 = _c(rays_second, np.float32).reshape(-1, 6)
        d = np.zeros(3, np.float32)
        f = self.lib.imr_ref_shoot; f.restype = None
        f.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, r1.ctypes.data, len(r1), r2.ctypes.data, len(r2), d.ctypes.data)
        return d, None

    def pair_rays(self, ta, ma, tb, mb):
        """The rays of one ordered pair in the reference's own order (n x 6 each)."""
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        f = self.lib.imr_ref_pair_rays; f.restype = None
        f.argtypes = [C.c_void_p] * 6 + [C.c_uint64, C.c_void_p]
        n2 = np.zeros(2, np.uint64)
        f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, None, None, 0, n2.ctypes.data)
        cap = int(max(n2.max(), 1))
        r1 = np.zeros((cap, 6), np.float32); r2 = np.zeros((cap, 6), np.float32)
        f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, r1.ctypes.data, r2.ctypes.data, cap, n2.ctypes.data)
        return r1[:int(n2[0])].copy(), r2[:int(n2[1])].copy()


class PortOracle(_Base):
    """The plain-C restatement (kind = "port")."""
    prefix = "imro_"
    kind = "port"

    def __init__(self, path=PORT_SO):
        super().__init__(path)

    def _sig(self):
        pass

    def repose(self, vertices, n_targets, joints, weights, morph_weights, joint_matrices, inverse_bind):
        """dynamicMeshShader_glsl.comp:99-145, position stream: vertices (n, T + 1, 4) f32, joints (n, G, 4) u16 / weights (n, G, 4) f32 or
        None, morph_weights (T,), matrices (J, 16) each -> (n, 4) f32."""
        v = _c(vertices, np.float32).reshape(-1, n_targets + 1, 4); n = v.shape[0]
        g = 0 if joints is None else _c(joints, np.uint16).reshape(n, -1, 4).shape[1]
        jn = None if joints is None else _c(joints, np.uint16); w = None if weights is None else _c(weights, np.float32)
        mw = _c(morph_weights if n_targets else np.zeros(1), np.float32)
        jm = None if joint_matrices is None else _c(joint_matrices, np.float32).reshape(-1, 16)
        ib = None if inverse_bind is None else _c(inverse_bind, np.float32).reshape(-1, 16)
        out = np.zeros((n, 4), np.float32)
        f = self.lib.imro_repose; f.restype = None
        f.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        f(n, n_targets, v.ctypes.data, g, _opt(jn), _opt(w), mw.ctypes.data, 0 if jm is None else jm.shape[0], _opt(jm), _opt(ib), out.ctypes.data)
        return out

    def eig3(self, A):
        A = _c(A, np.float64).reshape(9)
        V = np.empty(9, np.float64); d = np.empty(3, np.float64)
        f = self.lib.imro_eig3; f.restype = None; f.argtypes = [C.c_void_p] * 3
        f(A.ctypes.data, V.ctypes.data, d.ctypes.data)
        return V.reshape(3, 3), d

    def tree_import(self, ft: FlatTree) -> Tree:
        f = self.lib.imro_tree_import; f.restype = C.c_void_p
        f.argtypes = [C.c_uint64] + [C.c_void_p] * 5 + [C.c_uint64] + [C.c_void_p] * 4
        keep = [_c(ft.boxes, np.float32), _c(ft.left, np.int32), _c(ft.right, np.int32), _c(ft.tri_off, np.uint32),
                _c(ft.tri_cnt, np.uint32), _c(ft.tri_pos, np.float32), _c(ft.tri_nrm, np.float32),
                _c(ft.tri_vid, np.uint32), _c(ft.tri_orig, np.uint32)]
        h = f(ft.nv, keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data,
              keep[4].ctypes.data, ft.n_tri, keep[5].ctypes.data, keep[6].ctypes.data, keep[7].ctypes.data,
              keep[8].ctypes.data)
        return Tree(self, h)

    def extents(self, mats, root_boxes):
        mats = _c(mats, np.float32).reshape(-1, 16); rb = _c(root_boxes, np.float32).reshape(-1, 12)
        out = np.empty((mats.shape[0], 6), np.float32)
        f = self.lib.imro_extents; f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        f(mats.ctypes.data, rb.ctypes.data, mats.shape[0], out.ctypes.data)
        return out

    def broad(self, mats, trees, should_cb):
        """32-bit restatement of the sweep; `trees` may be Tree objects or an (n,12) array of root boxes."""
        import time
        mats = _c(mats, np.float32).reshape(-1, 16)
        n = mats.shape[0]
        if isinstance(trees, np.ndarray):
            rb = _c(trees, np.float32).reshape(-1, 12)
        else:
            rb = np.stack([t.root_box for t in trees]).astype(np.float32) if n else np.zeros((0, 12), np.float32)
        cb = _c(should_cb, np.uint8)
        f = self.lib.imro_broad; f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        cap = max(1024, 64 * n)
        while True:
            pairs = np.empty((cap, 2), np.uint32)
            t0 = time.perf_counter()
            k = f(mats.ctypes.data, rb.ctypes.data, cb.ctypes.data, n, pairs.ctypes.data, cap)
            dt = time.perf_counter() - t0
            if k <= cap:
                return pairs[:k].copy(), dt
            cap = int(k)

    def mid(self, ta, ma, tb, mb):
        import time
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        f = self.lib.imro_mid; f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        cap = 1 << 16
        while True:
            combos = np.empty((cap, 4), np.uint32)
            t0 = time.perf_counter()
            k = f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, combos.ctypes.data, cap, None)
            dt = time.perf_counter() - t0
            if k <= cap:
                return combos[:k].copy(), dt
            cap = int(k)

    def mid_stats(self, ta, ma, tb, mb):
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        out = np.zeros(5, np.uint64)
        f = self.lib.imro_mid; f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, None, 0, out.ctypes.data)
        return dict(visits=int(out[0]), passes=int(out[1]), combos=int(out[2]), tri_tests=int(out[3]), max_depth=int(out[4]))

    def pair(self, ta, ma, tb, mb, want_rays=False):
        import time
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        f = self.lib.imro_pair; f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        cap = 1 << 12
        ray_cap = 1 << 12
        while True:
            ids = np.empty((cap, 2), np.uint32); seg = np.empty((cap, 7), np.float32)
            summ = np.zeros(7, np.uint64); avg = np.zeros(6, np.float32)
            r1 = np.zeros((ray_cap, 6), np.float32) if want_rays else None
            r2 = np.zeros((ray_cap, 6), np.float32) if want_rays else None
            t0 = time.perf_counter()
            f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, ids.ctypes.data, seg.ctypes.data, cap,
              summ.ctypes.data, avg.ctypes.data, _opt(r1), _opt(r2), ray_cap)
            dt = time.perf_counter() - t0
            h = int(summ[2])
            if h <= cap and (not want_rays or max(int(summ[4]), int(summ[5])) <= ray_cap):
                res = PairResult(ids[:h].copy(), seg[:h].copy(), int(summ[0]), int(summ[1]), h, int(summ[3]),
                                 int(summ[4]), int(summ[5]), bool(summ[6]), avg, (0.0, dt))
                if want_rays:
                    res.rays = (r1[:int(summ[4])].copy(), r2[:int(summ[5])].copy())
                return res
            cap = max(cap, h)
            ray_cap = max(ray_cap, int(summ[4]), int(summ[5]))

    def shoot(self, ta, ma, tb, mb, rays_first, rays_second):
        """ShootUncollideRays::ExecuteShootUncollideRays on explicit ray lists (n x 6, first's model space): world-space delta, responses."""
        ma = _c(ma, np.float32); mb = _c(mb, np.float32)
        r1 = _c(rays_first, np.float32).reshape(-1, 6); r2 = _c(rays_second, np.float32).reshape(-1, 6)
        d = np.zeros(3, np.float32); nr = C.c_uint64(0)
        f = self.lib.imro_shoot; f.restype = None
        f.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        f(ta.h, ma.ctypes.data, tb.h, mb.ctypes.data, r1.ctypes.data, len(r1), r2.ctypes.data, len(r2), d.ctypes.data, C.byref(nr))
        return d, int(nr.value)


def frame_pairs(orc, mats, trees, pairs, threads: int = 1):
    """Mid + narrow over a pair list with the checker `orc` (the loop of CollisionDetection.cpp:44-69), split over
    `threads` host threads (ctypes releases the GIL; both libraries' batch entry points are re-entrant).
    Returns dict(combos, tri_tests, colliding, with_combos, wall_s, mid_s, narrow_s)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    mats = _c(mats, np.float32).reshape(-1, 16)
    pairs = _c(pairs, np.uint32).reshape(-1, 2)
    n = mats.shape[0]
    handles = (C.c_void_p * n)(*[t.h for t in trees])
    ref = orc.kind == "reference"
    f = orc.lib.imr_ref_frame_pairs if ref else orc.lib.imro_frame_pairs
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p] + ([C.c_void_p] if ref else [])

    def run(chunk):
        tot = np.zeros(4, np.uint64); secs = np.zeros(2, np.float64)
        if len(chunk):
            if ref:
                f(mats.ctypes.data, handles, chunk.ctypes.data, len(chunk), tot.ctypes.data, secs.ctypes.data)
            else:
                f(mats.ctypes.data, handles, chunk.ctypes.data, len(chunk), tot.ctypes.data)
        return tot, secs

    threads = max(1, int(threads))
    # interleaved slices balance the load (neighbouring pairs have similar cost)
    chunks = [np.ascontiguousarray(pairs[i::threads]) for i in range(threads)]
    t0 = time.perf_counter()
    if threads == 1:
        res = [run(chunks[0])]
    else:
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(run, chunks))
    wall = time.perf_counter() - t0
    tot = sum(r[0] for r in res); secs = sum(r[1] for r in res)
    return dict(combos=int(tot[0]), tri_tests=int(tot[1]), colliding=int(tot[2]), with_combos=int(tot[3]), wall_s=wall,
                mid_s=float(secs[0]), narrow_s=float(secs[1]))


def hit_fingerprint(tri_first, tri_second) -> int:
    """The order-free fingerprint of a hit set that imr_ref_frame_pairs_detail computes (sum of a 64-bit mix of the two ORIGINAL triangle
    indices, modulo 2^64), for hit lists that come from somewhere else (the device)."""
    x = (np.asarray(tri_first, np.uint64) << np.uint64(32)) | np.asarray(tri_second, np.uint64)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd); x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53); x ^= x >> np.uint64(33)
    return x


def frame_pairs_detail(orc, mats, trees, pairs, threads: int = 1) -> np.ndarray:
    """Per pair of `pairs`: (non-coplanar hits, coplanar hits, colliding, fingerprint of the hit set) from the unmodified reference
    (imr_ref_frame_pairs_detail), the pair loop split over `threads` host threads.  Returns (k, 3) u32 and (k,) u64."""
    from concurrent.futures import ThreadPoolExecutor
    assert orc.kind == "reference"
    mats = _c(mats, np.float32).reshape(-1, 16)
    pairs = _c(pairs, np.uint32).reshape(-1, 2)
    handles = (C.c_void_p * mats.shape[0])(*[t.h for t in trees])
    f = orc.lib.imr_ref_frame_pairs_detail
    f.restype = None; f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    out = np.zeros((len(pairs), 5), np.uint32)
    threads = max(1, int(threads))

    def run(t):
        chunk = np.ascontiguousarray(pairs[t::threads]); o = np.zeros((len(chunk), 5), np.uint32)
        if len(chunk):
            f(mats.ctypes.data, handles, chunk.ctypes.data, len(chunk), o.ctypes.data)
        out[t::threads] = o

    if threads == 1:
        run(0)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, range(threads)))
    return out[:, :3].copy(), out[:, 3].astype(np.uint64) | (out[:, 4].astype(np.uint64) << np.uint64(32))


class _RefGltfView(C.Structure):
    _fields_ = [("points", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_uint32)),
                ("n_points", C.c_uint64), ("n_indices", C.c_uint64), ("mode", C.c_uint32), ("skipped", C.c_uint32),
                ("source_index", C.c_uint32), ("has_indices", C.c_uint32)]


class RefGltf:
    """A .gltf / .glb read by the reference's own reader (tinygltf) and laid out the way the engine hands it to
    Triangle::CreateTriangleList (ref_gltf_shim.cpp): primitives(mesh) = [(points, normals | None, indices | None, mode, source index)]
    in the engine's recording order, skinned / morphed primitives left out."""

    def __init__(self, path):
        self.lib = C.CDLL(REF_GLTF_SO)
        self.lib.imr_refgltf_open.restype = C.c_void_p; self.lib.imr_refgltf_open.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        self.lib.imr_refgltf_close.argtypes = [C.c_void_p]
        self.lib.imr_refgltf_mesh_count.argtypes = [C.c_void_p]; self.lib.imr_refgltf_mesh_count.restype = C.c_uint32
        self.lib.imr_refgltf_primitive_count.argtypes = [C.c_void_p, C.c_uint32]; self.lib.imr_refgltf_primitive_count.restype = C.c_uint32
        self.lib.imr_refgltf_primitive.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]; self.lib.imr_refgltf_primitive.restype = None
        err = C.create_string_buffer(512)
        self.h = self.lib.imr_refgltf_open(os.fsencode(path), err, len(err))
        if not self.h:
            raise ValueError(err.value.decode(errors="replace"))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.imr_refgltf_close(self.h); self.h = None

    @property
    def n_meshes(self):
        return int(self.lib.imr_refgltf_mesh_count(self.h))

    def primitives(self, mesh):
        out = []
        for k in range(self.lib.imr_refgltf_primitive_count(self.h, mesh)):
            v = _RefGltfView()
            self.lib.imr_refgltf_primitive(self.h, mesh, k, C.byref(v))
            if v.skipped:
                continue
            npts, nidx = int(v.n_points), int(v.n_indices)
            pts = np.ctypeslib.as_array(v.points, (npts, 3)).copy() if npts else np.zeros((0, 3), np.float32)
            nrm = np.ctypeslib.as_array(v.normals, (npts, 3)).copy() if v.normals else None
            idx = (np.ctypeslib.as_array(v.indices, (nidx,)).copy() if nidx else np.zeros(0, np.uint32)) if v.has_indices else None
            out.append((pts, nrm, idx, int(v.mode), int(v.source_index)))
        return out


def load(prefer: str = "reference"):
    """Return the strongest available checker: the real reference if its .so exists, else the port."""
    if prefer == "reference" and os.path.exists(REF_SO):
        return RefOracle()
    if not os.path.exists(PORT_SO):
        build("port")
    return PortOracle()
