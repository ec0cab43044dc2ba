"""Host-side mirror of the reference's collision interface over the C ABI (include/imrcd.h).

Names and call order follow the reference ("IMR/" = inMyRoom_vulkan/):
  CollisionDetection.Reset / AddCollisionDetectionEntry / ExecuteCollisionDetection
      IMR/include/CollisionDetection/CollisionDetection.h:11-35, src/CollisionDetection/CollisionDetection.cpp:28-129
  CollisionDetectionEntry / CollisionCallbackData     IMR/include/ECS/ECStypes.h:149-163
  OBBtree(triangles)                                  IMR/include/Geometry/OBBtree.h:105
The C++ adapter a maintainer would drop into the engine is csrc/host/ (see INTEGRATION.md); this Python
layer exists so the parity tests and bench.py can drive the same C entry points.  All compute happens in
libimrcd.so on the GPU; nothing here computes geometry.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import IMRCD_BUILD_MORTON, IMRCD_BUILD_REFERENCE, EntityPair, FrameStats, TriHit  # noqa: F401


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ImrcdError(RuntimeError):
    pass


class Context:
    """One GPU context (imrcd_ctx).  `stream` is a raw cudaStream_t (int) or None."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.imrcd_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise ImrcdError(f"imrcd_create(device={device}) failed: {_lib.ERRORS.get(rc, rc)}")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.imrcd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            msg = self.lib.imrcd_last_error(self.h)
            raise ImrcdError(f"{_lib.ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")

    # ---- multi-GPU: the end-of-frame merge lives in the library (imrcd_comm_*) ----------
    def comm_unique_id(self) -> bytes:
        _lib.preload_nccl()
        buf = (C.c_uint8 * 128)()
        self.check(self.lib.imrcd_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, n_ranks: int):
        """Collective over the ranks that share `uid` (one context per GPU); sets the frame shard to (rank, n_ranks)."""
        _lib.preload_nccl()
        raw = (C.c_uint8 * 128).from_buffer_copy(uid)
        self.check(self.lib.imrcd_comm_init(self.h, raw, rank, n_ranks))

    def comm_destroy(self):
        self.check(self.lib.imrcd_comm_destroy(self.h))

    def comm_transport(self) -> int:
        """0 no communicator, 1 ncclAllGather, 2 peer memory over NVLink (imrcd_comm_transport)."""
        return int(self.lib.imrcd_comm_transport(self.h))

    # ---- unit-level hooks ---------------------------------------------------------------
    def test_sat(self, boxes_a, boxes_b, mats=None):
        a = _c(boxes_a, np.float32).reshape(-1, 12); b = _c(boxes_b, np.float32).reshape(-1, 12)
        m = None if mats is None else _c(mats, np.float32).reshape(-1, 16)
        n = a.shape[0]
        v = np.zeros(n, np.uint8); sa = np.zeros(n, np.float32); sb = np.zeros(n, np.float32)
        self.check(self.lib.imrcd_test_sat(self.h, n, _ptr(a), _ptr(b), _ptr(m), _ptr(v), _ptr(sa), _ptr(sb)))
        return v, sa, sb

    def test_tri_tri(self, tris_a, tris_b, mat16=None):
        a = _c(tris_a, np.float32).reshape(-1, 9); b = _c(tris_b, np.float32).reshape(-1, 9)
        m = None if mat16 is None else _c(mat16, np.float32).reshape(16)
        n = a.shape[0]
        flags = np.zeros(n, np.uint8); seg = np.zeros((n, 6), np.float32)
        self.check(self.lib.imrcd_test_tri_tri(self.h, n, _ptr(a), _ptr(b), _ptr(m), _ptr(flags), _ptr(seg)))
        return flags, seg

    def test_pair_matrix(self, a, b):
        a = _c(a, np.float32).reshape(-1, 16); b = _c(b, np.float32).reshape(-1, 16)
        out = np.zeros_like(a)
        self.check(self.lib.imrcd_test_pair_matrix(self.h, a.shape[0], _ptr(a), _ptr(b), _ptr(out)))
        return out

    def test_obb_fit(self, points):
        p = _c(points, np.float32).reshape(-1, 3)
        out = np.zeros(12, np.float32)
        self.check(self.lib.imrcd_test_obb_fit(self.h, p.shape[0], _ptr(p), _ptr(out)))
        return out

    def test_ray_tree(self, tree, mats, origins, directions):
        """Ray::IntersectOBBtree on n rays: (hit, backface, distance, bary (n,2), leaf-order triangle index)."""
        m = _c(mats, np.float32).reshape(-1, 16); o = _c(origins, np.float32).reshape(-1, 3); d = _c(directions, np.float32).reshape(-1, 3)
        n = m.shape[0]
        flags = np.zeros(n, np.uint8); out = np.zeros((n, 3), np.float32); tri = np.zeros(n, np.uint32)
        self.check(self.lib.imrcd_test_ray_tree(self.h, tree.mesh_id, n, _ptr(m), _ptr(o), _ptr(d), _ptr(flags), _ptr(out), _ptr(tri)))
        return (flags & 1).astype(bool), (flags & 2).astype(bool), out[:, 0].copy(), out[:, 1:3].copy(), tri


class Skin:
    """What the engine's dynamic-mesh pass reads per vertex (imrcd_skin_create): vertices (n, T + 1, 4) = base + morph targets,
    joints (n, G, 4) u16 and weights (n, G, 4) f32 (None: morph only)."""

    def __init__(self, ctx: "Context", vertices, joints=None, weights=None):
        v = _c(vertices, np.float32)
        assert v.ndim == 3 and v.shape[2] == 4
        self.ctx = ctx; self.n_vertices = v.shape[0]; self.n_targets = v.shape[1] - 1
        jn = None if joints is None else _c(joints, np.uint16).reshape(self.n_vertices, -1, 4)
        w = None if weights is None else _c(weights, np.float32).reshape(self.n_vertices, -1, 4)
        self.n_groups = 0 if jn is None else jn.shape[1]
        sid = C.c_uint32()
        ctx.check(ctx.lib.imrcd_skin_create(ctx.h, self.n_vertices, self.n_targets, _ptr(v), self.n_groups, _ptr(jn), _ptr(w), C.byref(sid)))
        self.skin_id = sid.value


def repose_meshes(ctx: "Context", trees, morph_weights=None, joint_matrices=None, inverse_bind=None) -> None:
    """imrcd_meshes_repose: one batched pass over `trees` (each bound to a Skin).  morph_weights: (len(trees), T) or None;
    joint_matrices / inverse_bind: (len(trees), J, 16) or None.  Follow with refit_meshes(ctx)."""
    ids = np.array([t.mesh_id for t in trees], np.uint32)
    mw = None if morph_weights is None else _c(morph_weights, np.float32)
    jm = None if joint_matrices is None else _c(joint_matrices, np.float32).reshape(len(trees), -1, 16)
    ib = None if inverse_bind is None else _c(inverse_bind, np.float32).reshape(len(trees), -1, 16)
    nj = None if jm is None else np.full(len(trees), jm.shape[1], np.uint32)
    ctx.check(ctx.lib.imrcd_meshes_repose(ctx.h, len(ids), _ptr(ids), _ptr(mw), _ptr(jm), _ptr(ib), _ptr(nj)))


def last_repose_ms(ctx: "Context") -> float:
    ms = C.c_float()
    ctx.check(ctx.lib.imrcd_mesh_last_repose_ms(ctx.h, C.byref(ms)))
    return ms.value


def reposed_vertices(ctx: "Context", n_vertices: int) -> np.ndarray:
    out = np.zeros((n_vertices, 4), np.float32)
    ctx.check(ctx.lib.imrcd_test_reposed_vertices(ctx.h, _ptr(out), n_vertices))
    return out


def refit_meshes(ctx: "Context", trees=None) -> float:
    """Batched refit of `trees` (None = every mesh updated since its last refit); returns the device time in ms."""
    if trees is None:
        ctx.check(ctx.lib.imrcd_mesh_refit(ctx.h, None, 0))
    else:
        ids = np.array([t.mesh_id for t in trees], np.uint32)
        ctx.check(ctx.lib.imrcd_mesh_refit(ctx.h, _ptr(ids), len(ids)))
    ms = C.c_float()
    ctx.check(ctx.lib.imrcd_mesh_last_refit_ms(ctx.h, C.byref(ms)))
    return ms.value


class OBBtree:
    """Device-resident OBB tree of one mesh: the replacement for OBBtree::OBBtree(std::vector<Triangle>&&)."""

    def __init__(self, ctx: Context, positions, normals=None, vertex_ids=None, build_mode: int = IMRCD_BUILD_MORTON):
        self.ctx = ctx
        pos = _c(positions, np.float32).reshape(-1, 9)
        nrm = None if normals is None else _c(normals, np.float32).reshape(-1, 9)
        vid = None if vertex_ids is None else _c(vertex_ids, np.uint32).reshape(-1, 3)
        mid = C.c_uint32()
        ctx.check(ctx.lib.imrcd_mesh_create(ctx.h, _ptr(pos), _ptr(nrm), _ptr(vid), pos.shape[0], build_mode, C.byref(mid)))
        self.mesh_id = mid.value

    @classmethod
    def from_flat(cls, ctx: Context, flat) -> "OBBtree":
        """Test-only: upload a tree built elsewhere (an oracle FlatTree or anything with the same fields)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        keep = [_c(flat.boxes, np.float32), _c(flat.left, np.int32), _c(flat.right, np.int32), _c(flat.tri_off, np.uint32),
                _c(flat.tri_cnt, np.uint32), _c(flat.tri_pos, np.float32), _c(flat.tri_nrm, np.float32),
                _c(flat.tri_vid, np.uint32), _c(flat.tri_orig, np.uint32)]
        mid = C.c_uint32()
        ctx.check(ctx.lib.imrcd_mesh_import_tree(ctx.h, keep[0].shape[0], _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]),
                                                 _ptr(keep[4]), keep[5].shape[0], _ptr(keep[5]), _ptr(keep[6]), _ptr(keep[7]),
                                                 _ptr(keep[8]), C.byref(mid)))
        self.mesh_id = mid.value
        return self

    def info(self):
        nt = C.c_uint64(); nv = C.c_uint64()
        self.ctx.check(self.ctx.lib.imrcd_mesh_info(self.ctx.h, self.mesh_id, C.byref(nt), C.byref(nv)))
        return nt.value, nv.value

    def build_ms(self) -> float:
        ms = C.c_float()
        self.ctx.check(self.ctx.lib.imrcd_mesh_last_build_ms(self.ctx.h, C.byref(ms)))
        return ms.value

    @classmethod
    def from_mesh_id(cls, ctx: Context, mesh_id: int):
        """Wrap a mesh the library already holds (imrcd_gltf_load, imrcd_mesh_end)."""
        self = cls.__new__(cls)
        self.ctx = ctx; self.mesh_id = int(mesh_id)
        return self

    @classmethod
    def from_primitives(cls, ctx: Context, primitives, build_mode: int = IMRCD_BUILD_MORTON):
        """The engine's way in (PrimitivesOfMeshes::StartRecordOBBtree / GetOBBtreeAndReset): `primitives` is a sequence of
        (points (n, 3 or 4), normals or None, indices or None, glTF draw mode); Triangle::CreateTriangleList runs on the device."""
        ctx.check(ctx.lib.imrcd_mesh_begin(ctx.h))
        for points, normals, indices, mode in primitives:
            pts = _c(points, np.float32)
            stride = pts.shape[1]
            nrm = None if normals is None else _c(normals, np.float32)
            assert nrm is None or nrm.shape == pts.shape
            idx = None if indices is None else _c(indices, np.uint32).reshape(-1)
            ctx.check(ctx.lib.imrcd_mesh_add_primitive(ctx.h, _ptr(pts), pts.shape[0], stride, _ptr(nrm), _ptr(idx), 0 if idx is None else len(idx), int(mode)))
        mid = C.c_uint32()
        ctx.check(ctx.lib.imrcd_mesh_end(ctx.h, int(build_mode), C.byref(mid)))
        self = cls.__new__(cls)
        self.ctx = ctx; self.mesh_id = mid.value
        return self

    def bind_skin(self, skin: "Skin"):
        self.ctx.check(self.ctx.lib.imrcd_mesh_bind_skin(self.ctx.h, self.mesh_id, skin.skin_id))

    def update_positions(self, positions, normals=None):
        """New triangle positions (original input order); call refit() / refit_meshes() before the next frame."""
        pos = _c(positions, np.float32).reshape(-1, 9)
        nrm = None if normals is None else _c(normals, np.float32).reshape(-1, 9)
        assert pos.shape[0] == self.info()[0]
        self.ctx.check(self.ctx.lib.imrcd_mesh_update_positions(self.ctx.h, self.mesh_id, _ptr(pos), _ptr(nrm)))

    def refit(self):
        ids = np.array([self.mesh_id], np.uint32)
        self.ctx.check(self.ctx.lib.imrcd_mesh_refit(self.ctx.h, _ptr(ids), 1))

    def export(self):
        """Read the tree back in the flat pre-order form (returns a SimpleNamespace with FlatTree's fields)."""
        from types import SimpleNamespace
        n, nv = self.info()
        o = SimpleNamespace(boxes=np.zeros((nv, 12), np.float32), left=np.zeros(nv, np.int32), right=np.zeros(nv, np.int32),
                            tri_off=np.zeros(nv, np.uint32), tri_cnt=np.zeros(nv, np.uint32), tri_pos=np.zeros((n, 9), np.float32),
                            tri_nrm=np.zeros((n, 9), np.float32), tri_vid=np.zeros((n, 3), np.uint32), tri_orig=np.zeros(n, np.uint32))
        self.ctx.check(self.ctx.lib.imrcd_mesh_export_tree(self.ctx.h, self.mesh_id, _ptr(o.boxes), _ptr(o.left), _ptr(o.right),
                                                           _ptr(o.tri_off), _ptr(o.tri_cnt), _ptr(o.tri_pos), _ptr(o.tri_nrm),
                                                           _ptr(o.tri_vid), _ptr(o.tri_orig)))
        o.nv = nv; o.n_tri = n
        return o


@dataclass
class CollisionDetectionEntry:          # IMR/include/ECS/ECStypes.h:149-156
    currentGlobalMatrix: np.ndarray     # 16 floats, column-major
    previousGlobalMatrix: np.ndarray
    OBBtree_ptr: OBBtree
    shouldCallback: bool
    entity: int


@dataclass
class CollisionCallbackData:            # IMR/include/ECS/ECStypes.h:158-163
    familyEntity: int
    collideWithEntity: int
    deltaVector: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))


class CollisionDetection:
    """Drop-in for the reference's `class CollisionDetection` (same three calls, same callback fan-out).

    `ecs` is any object with `GetEntityAncestors(entity) -> list[int]` (EntitiesHandler.cpp:186-204, root first)
    and `components` (iterable of objects with `CollisionCallback(list[(entity, list[CollisionCallbackData])])`,
    ComponentBaseClass.h:25); it may be None when only the result arrays are wanted.
    """

    def __init__(self, ecs=None, ctx: Optional[Context] = None, device: int = 0, stream: Optional[int] = None):
        self.ctx = ctx or Context(device, stream)
        self.lib = self.ctx.lib
        self.ecs = ecs
        self._n = 0

    # -- the reference's three calls ------------------------------------------------------
    def Reset(self):                                        # CollisionDetection.cpp:28
        self.ctx.check(self.lib.imrcd_frame_reset(self.ctx.h))
        self._n = 0

    def AddCollisionDetectionEntry(self, e: CollisionDetectionEntry):   # CollisionDetection.cpp:33
        cur = _c(e.currentGlobalMatrix, np.float32).reshape(16); prev = _c(e.previousGlobalMatrix, np.float32).reshape(16)
        self.ctx.check(self.lib.imrcd_frame_add_entry(self.ctx.h, _ptr(cur), _ptr(prev), e.OBBtree_ptr.mesh_id,
                                                      1 if e.shouldCallback else 0, int(e.entity)))
        self._n += 1

    def ExecuteCollisionDetection(self):                    # CollisionDetection.cpp:38-129
        if self._n < 2:                                     # :40
            return
        self.ctx.check(self.lib.imrcd_frame_execute(self.ctx.h))
        if self.ecs is not None:
            self._make_callbacks()

    # -- bulk path (numpy arrays instead of one call per entry) ----------------------------
    def add_entries(self, matrices, mesh_ids, should_callback=None, entities=None, previous=None):
        m = _c(matrices, np.float32).reshape(-1, 16)
        n = m.shape[0]
        p = None if previous is None else _c(previous, np.float32).reshape(-1, 16)
        mid = _c(mesh_ids, np.uint32).reshape(n)
        cb = None if should_callback is None else _c(should_callback, np.uint8).reshape(n)
        ent = None if entities is None else _c(entities, np.uint32).reshape(n)
        self.ctx.check(self.lib.imrcd_frame_add_entries(self.ctx.h, n, _ptr(m), _ptr(p), _ptr(mid), _ptr(cb), _ptr(ent)))
        self._n += n

    def map_entries(self, n: int):
        """Zero-copy submission: numpy views of the context's pinned staging for `n` more entries
        (current (n,16) f32, previous (n,16) f32, mesh_ids (n,) u32, should_callback (n,) u8, entities (n,) u32).
        Fill them, then call commit_entries(n)."""
        from types import SimpleNamespace
        ptrs = [C.c_void_p() for _ in range(5)]
        self.ctx.check(self.lib.imrcd_frame_map_entries(self.ctx.h, n, *[C.byref(p) for p in ptrs]))
        def view(p, ctype, shape):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=shape)
        return SimpleNamespace(current=view(ptrs[0], C.c_float, (n, 16)), previous=view(ptrs[1], C.c_float, (n, 16)),
                               mesh_ids=view(ptrs[2], C.c_uint32, (n,)), should_callback=view(ptrs[3], C.c_uint8, (n,)),
                               entities=view(ptrs[4], C.c_uint32, (n,)))

    def commit_entries(self, n: int, previous_valid: bool = False):
        self.ctx.check(self.lib.imrcd_frame_commit_entries(self.ctx.h, n, 1 if previous_valid else 0))
        self._n += n

    def set_shard(self, rank: int, n_ranks: int):
        self.ctx.check(self.lib.imrcd_frame_set_shard(self.ctx.h, rank, n_ranks))

    def upload(self):
        self.ctx.check(self.lib.imrcd_frame_upload(self.ctx.h))

    def run(self):
        self.ctx.check(self.lib.imrcd_frame_run(self.ctx.h))

    def run_async(self):
        """Enqueue every kernel of the frame and return at once (imrcd_frame_run_async); finish() must follow."""
        self.ctx.check(self.lib.imrcd_frame_run_async(self.ctx.h))

    def finish(self) -> bool:
        """Wait for a frame started with run_async(); True when a buffer overflowed and the frame was run again."""
        rc = self.lib.imrcd_frame_finish(self.ctx.h)
        if rc < 0:
            self.ctx.check(rc)
        return rc == 1

    def fetch(self):
        self.ctx.check(self.lib.imrcd_frame_fetch(self.ctx.h))

    # -- results ------------------------------------------------------------------------------
    def stats(self) -> dict:
        st = FrameStats()
        self.ctx.check(self.lib.imrcd_frame_get_stats(self.ctx.h, C.byref(st)))
        return st.as_dict()

    def results(self, want_hits: bool = True):
        """(entity_pairs, hits) as numpy structured arrays (copies)."""
        pp = C.POINTER(EntityPair)(); np_ = C.c_uint64(); hp = C.POINTER(TriHit)(); nh = C.c_uint64()
        self.ctx.check(self.lib.imrcd_frame_results(self.ctx.h, C.byref(pp), C.byref(np_), C.byref(hp) if want_hits else None, C.byref(nh)))
        pairs = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_uint8)), shape=(np_.value * C.sizeof(EntityPair),)).copy() if np_.value else np.zeros(0, np.uint8)
        pairs = pairs.view(PAIR_DTYPE)
        if want_hits and nh.value:
            hits = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nh.value * C.sizeof(TriHit),)).copy().view(HIT_DTYPE)
        else:
            hits = np.zeros(0, HIT_DTYPE)
        return pairs, hits

    def results_local(self) -> np.ndarray:
        """This rank's own colliding pairs (results() gives the merged records of all ranks when the context has a communicator)."""
        pp = C.POINTER(EntityPair)(); np_ = C.c_uint64()
        self.ctx.check(self.lib.imrcd_frame_results_local(self.ctx.h, C.byref(pp), C.byref(np_)))
        if not np_.value:
            return np.zeros(0, PAIR_DTYPE)
        return np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_uint8)), shape=(np_.value * C.sizeof(EntityPair),)).copy().view(PAIR_DTYPE)

    def broad_pairs(self) -> np.ndarray:
        pp = C.POINTER(C.c_uint32)(); n = C.c_uint64()
        self.ctx.check(self.lib.imrcd_frame_pairs(self.ctx.h, C.byref(pp), C.byref(n)))
        if not n.value:
            return np.zeros((0, 2), np.uint32)
        return np.ctypeslib.as_array(pp, shape=(n.value, 2)).copy()

    def combos(self) -> np.ndarray:
        """(k,5) u32: pair, offA, cntA, offB, cntB"""
        pp = C.POINTER(C.c_uint32)(); n = C.c_uint64()
        self.ctx.check(self.lib.imrcd_frame_combos(self.ctx.h, C.byref(pp), C.byref(n)))
        if not n.value:
            return np.zeros((0, 5), np.uint32)
        raw = np.ctypeslib.as_array(pp, shape=(n.value, 4)).copy()
        return np.stack([raw[:, 0], raw[:, 1], raw[:, 3] & 0xffff, raw[:, 2], raw[:, 3] >> 16], 1)

    # -- callback fan-out, CollisionDetection.cpp:70-141 --------------------------------------
    def _make_callbacks(self):
        pairs, _ = self.results(want_hits=False)
        callbacks: dict[int, list[CollisionCallbackData]] = {}
        for p in pairs:
            first = CollisionCallbackData(int(p["entity_first"]), int(p["entity_second"]), np.array(p["delta_first"], np.float32))
            second = CollisionCallbackData(int(p["entity_second"]), int(p["entity_first"]), np.array(p["delta_second"], np.float32))
            fa = self.ecs.GetEntityAncestors(first.familyEntity)
            sa = self.ecs.GetEntityAncestors(second.familyEntity)
            for i, a in enumerate(fa):                                   # :109-116
                if i >= len(sa) or fa[i] != sa[i]:
                    callbacks.setdefault(a, []).append(first)
            for i, a in enumerate(sa):                                   # :118-125
                if i >= len(fa) or fa[i] != sa[i]:
                    callbacks.setdefault(a, []).append(second)
        vec = list(callbacks.items())                                    # MakeCallbacks :131-141
        for comp in self.ecs.components:
            if comp is not None:
                comp.CollisionCallback(vec)


PAIR_DTYPE = np.dtype([("entry_first", "<u4"), ("entry_second", "<u4"), ("entity_first", "<u4"), ("entity_second", "<u4"),
                       ("n_hits", "<u4"), ("n_rays_first", "<u4"), ("n_rays_second", "<u4"), ("flags", "<u4"),
                       ("avg_first", "<f4", 3), ("avg_second", "<f4", 3), ("delta_first", "<f4", 3), ("delta_second", "<f4", 3)])
HIT_DTYPE = np.dtype([("pair", "<u4"), ("tri_first", "<u4"), ("tri_second", "<u4"), ("source", "<f4", 3), ("target", "<f4", 3),
                      ("weight", "<f4")])
assert PAIR_DTYPE.itemsize == 80 and HIT_DTYPE.itemsize == 40


class Group:
    """One process, several GPUs (imrcd_group_*): meshes replicated, every frame sharded over the devices and merged by the library's
    own NCCL all-gather; the calls mirror CollisionDetection's."""

    def __init__(self, device_ids: Sequence[int]):
        self.lib = _lib.load()
        _lib.preload_nccl()
        ids = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
        h = C.c_void_p()
        rc = self.lib.imrcd_group_create(ids, len(device_ids), C.byref(h))
        if rc != 0:
            raise ImrcdError(f"imrcd_group_create({list(device_ids)}) failed: {_lib.ERRORS.get(rc, rc)}")
        self.h = h
        self.n = len(device_ids)

    def close(self):
        if getattr(self, "h", None):
            self.lib.imrcd_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            msg = self.lib.imrcd_group_last_error(self.h)
            raise ImrcdError(f"{_lib.ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")

    def comm_transport(self) -> int:
        """How the group's end-of-frame merge travels (imrcd_comm_transport of its first context): 1 ncclAllGather, 2 peer memory."""
        self.lib.imrcd_group_ctx.restype = C.c_void_p
        return int(self.lib.imrcd_comm_transport(C.c_void_p(self.lib.imrcd_group_ctx(self.h, 0))))

    def mesh_create(self, positions, normals=None, vertex_ids=None, build_mode: int = IMRCD_BUILD_MORTON) -> int:
        pos = _c(positions, np.float32).reshape(-1, 9)
        nrm = None if normals is None else _c(normals, np.float32).reshape(-1, 9)
        vid = None if vertex_ids is None else _c(vertex_ids, np.uint32).reshape(-1, 3)
        mid = C.c_uint32()
        self.check(self.lib.imrcd_group_mesh_create(self.h, _ptr(pos), _ptr(nrm), _ptr(vid), pos.shape[0], build_mode, C.byref(mid)))
        return mid.value

    def Reset(self):
        self.check(self.lib.imrcd_group_frame_reset(self.h))

    def add_entries(self, matrices, mesh_ids, should_callback=None, entities=None, previous=None):
        m = _c(matrices, np.float32).reshape(-1, 16); n = m.shape[0]
        p = None if previous is None else _c(previous, np.float32).reshape(-1, 16)
        mid = _c(mesh_ids, np.uint32).reshape(n)
        cb = None if should_callback is None else _c(should_callback, np.uint8).reshape(n)
        ent = None if entities is None else _c(entities, np.uint32).reshape(n)
        self.check(self.lib.imrcd_group_frame_add_entries(self.h, n, _ptr(m), _ptr(p), _ptr(mid), _ptr(cb), _ptr(ent)))

    def ExecuteCollisionDetection(self):
        self.check(self.lib.imrcd_group_frame_execute(self.h))

    def results(self) -> np.ndarray:
        pp = C.POINTER(EntityPair)(); np_ = C.c_uint64()
        self.check(self.lib.imrcd_group_frame_results(self.h, C.byref(pp), C.byref(np_)))
        if not np_.value:
            return np.zeros(0, PAIR_DTYPE)
        return np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_uint8)), shape=(np_.value * C.sizeof(EntityPair),)).copy().view(PAIR_DTYPE)
