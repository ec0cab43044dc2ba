"""glTF files -> collision trees, the Python face of imrcd_gltf_* (include/imrcd.h).

Mirrors the engine's load path for collision geometry -- tinygltf + PrimitivesOfMeshes::AddPrimitive per primitive, one
OBBtree per glTF mesh (IMR/src/Graphics/Meshes/MeshesOfNodes.cpp:34-53, PrimitivesOfMeshes.cpp:44-175,835-863).  The file
is read on the host by the library; triangles (Triangle::CreateTriangleList) and trees are made on the device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .collision import IMRCD_BUILD_MORTON, Context, OBBtree


class _PrimitiveView(C.Structure):
    _fields_ = [("points", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_uint32)),
                ("n_points", C.c_uint64), ("n_indices", C.c_uint64), ("mode", C.c_uint32), ("skipped", C.c_uint32),
                ("source_index", C.c_uint32), ("has_indices", C.c_uint32)]


class GltfFile:
    """An opened .gltf / .glb: host only, no device needed.  primitives(mesh) lists a mesh's primitives exactly as they are
    handed to imrcd_mesh_add_primitive, in the order the reference would record them."""

    def __init__(self, path):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.imrcd_gltf_open(os.fsencode(path), C.byref(self.h), err, len(err))
        if rc:
            self.h = None
            raise ValueError(err.value.decode(errors="replace") or f"imrcd_gltf_open failed ({rc})")

    def close(self):
        if self.h:
            self.lib.imrcd_gltf_close(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_meshes(self) -> int:
        n = C.c_uint32()
        assert self.lib.imrcd_gltf_mesh_count(self.h, C.byref(n)) == 0
        return n.value

    def primitives(self, mesh: int):
        """[(points (n, 3) f32, normals or None, indices u32 or None, draw mode, source index)] of the primitives that go into
        the tree (skinned / morphed ones are left out, as in the reference)."""
        n = C.c_uint32()
        if self.lib.imrcd_gltf_primitive_count(self.h, mesh, C.byref(n)):
            raise IndexError(mesh)
        out = []
        for k in range(n.value):
            v = _PrimitiveView()
            assert self.lib.imrcd_gltf_primitive(self.h, mesh, k, C.byref(v)) == 0
            if v.skipped:
                continue
            npts, nidx = int(v.n_points), int(v.n_indices)
            pts = np.ctypeslib.as_array(v.points, (npts, 3)).copy() if npts else np.zeros((0, 3), np.float32)
            nrm = np.ctypeslib.as_array(v.normals, (npts, 3)).copy() if v.normals else None
            if v.has_indices:
                idx = np.ctypeslib.as_array(v.indices, (nidx,)).copy() if nidx else np.zeros(0, np.uint32)
            else:
                idx = None
            out.append((pts, nrm, idx, int(v.mode), int(v.source_index)))
        return out

    def build_mesh(self, ctx: Context, mesh: int, build_mode: int = IMRCD_BUILD_MORTON) -> OBBtree:
        mid = C.c_uint32()
        ctx.check(self.lib.imrcd_gltf_build_mesh(ctx.h, self.h, mesh, int(build_mode), C.byref(mid)))
        return OBBtree.from_mesh_id(ctx, mid.value)


def load_gltf(ctx: Context, path, build_mode: int = IMRCD_BUILD_MORTON):
    """One OBBtree per mesh of the file (imrcd_gltf_load), index i = meshes[i]."""
    n = C.c_uint32()
    ctx.check(ctx.lib.imrcd_gltf_load(ctx.h, os.fsencode(path), int(build_mode), None, 0, C.byref(n)))
    ids = (C.c_uint32 * max(n.value, 1))()
    ctx.check(ctx.lib.imrcd_gltf_load(ctx.h, os.fsencode(path), int(build_mode), ids, n.value, C.byref(n)))
    return [OBBtree.from_mesh_id(ctx, ids[i]) for i in range(n.value)]
