"""A small glTF 2.0 writer for the F4 tests: the same meshes in the three containers the engine loads (.gltf + .bin, .gltf with a
base64 data URI, .glb).  Test infrastructure only."""
import base64
import json
import os
import struct

import numpy as np

from inmyroom_vulkan_b200 import scenes

CT = {np.dtype(np.uint8): 5121, np.dtype(np.uint16): 5123, np.dtype(np.uint32): 5125, np.dtype(np.float32): 5126}
TYPE = {1: "SCALAR", 2: "VEC2", 3: "VEC3", 4: "VEC4"}


class _Builder:
    def __init__(self):
        self.blob = bytearray(); self.views = []; self.accessors = []

    def view(self, data: bytes, stride=None):
        while len(self.blob) % 4:
            self.blob += b"\0"
        v = {"buffer": 0, "byteOffset": len(self.blob), "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        self.blob += data
        self.views.append(v)
        return len(self.views) - 1

    def accessor(self, a: np.ndarray, with_bounds=False):
        a = np.ascontiguousarray(a)
        comps = 1 if a.ndim == 1 else a.shape[1]
        acc = {"bufferView": self.view(a.tobytes()), "componentType": CT[a.dtype], "count": int(a.shape[0]), "type": TYPE[comps]}
        if with_bounds and a.shape[0]:
            acc["min"] = [float(x) for x in a.min(0)]; acc["max"] = [float(x) for x in a.max(0)]
        self.accessors.append(acc)
        return len(self.accessors) - 1

    def interleaved(self, pts, nrm):
        """POSITION and NORMAL sharing one strided bufferView (the reference reads tightly packed only: product-only case)."""
        both = np.ascontiguousarray(np.concatenate([pts, nrm], 1).astype(np.float32))
        v = self.view(both.tobytes(), stride=24)
        ids = []
        for off in (0, 12):
            self.accessors.append({"bufferView": v, "byteOffset": off, "componentType": 5126, "count": int(len(pts)), "type": "VEC3"})
            ids.append(len(self.accessors) - 1)
        return ids


def write(path, meshes, container="bin"):
    """meshes: [[primitive dict]] with keys points (n,3) f32 [glTF axes], normals | None, indices | None (u8 / u16 / u32), mode | None,
    skinned (bool), morph (bool), interleave (bool).  container: 'bin' (.gltf + .bin), 'uri' (.gltf, base64), 'glb'."""
    b = _Builder()
    jm = []
    for mi, prims in enumerate(meshes):
        jp = []
        for p in prims:
            pts = np.asarray(p["points"], np.float32); nrm = p.get("normals")
            attrs = {}
            if p.get("interleave") and nrm is not None:
                attrs["POSITION"], attrs["NORMAL"] = b.interleaved(pts, np.asarray(nrm, np.float32))
            else:
                attrs["POSITION"] = b.accessor(pts, with_bounds=True)
                if nrm is not None:
                    attrs["NORMAL"] = b.accessor(np.asarray(nrm, np.float32))
            d = {"attributes": attrs}
            if p.get("indices") is not None:
                d["indices"] = b.accessor(np.asarray(p["indices"]))
            if p.get("mode") is not None:
                d["mode"] = int(p["mode"])
            if p.get("skinned"):
                attrs["JOINTS_0"] = b.accessor(np.zeros((len(pts), 4), np.uint16))
                attrs["WEIGHTS_0"] = b.accessor(np.tile(np.array([1, 0, 0, 0], np.float32), (len(pts), 1)))
            if p.get("morph"):
                d["targets"] = [{"POSITION": b.accessor(np.full_like(pts, 0.25), with_bounds=True)}]
            jp.append(d)
        jm.append({"name": f"mesh é{mi} \"q\"", "primitives": jp})
    doc = {"asset": {"version": "2.0", "generator": "tests/gltf_writer.py"}, "scene": 0, "scenes": [{"nodes": list(range(len(jm)))}],
           "nodes": [{"mesh": i, "translation": [float(i), 0.0, -1.5e-3]} for i in range(len(jm))],
           "meshes": jm, "accessors": b.accessors, "bufferViews": b.views, "buffers": [{"byteLength": len(b.blob)}]}
    blob = bytes(b.blob)
    if container == "glb":
        js = json.dumps(doc).encode()
        js += b" " * (-len(js) % 4)
        bn = blob + b"\0" * (-len(blob) % 4)
        with open(path, "wb") as f:
            f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(bn)))
            f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
            f.write(struct.pack("<II", len(bn), 0x004E4942) + bn)
        return path
    if container == "uri":
        doc["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
    else:
        name = os.path.splitext(os.path.basename(path))[0] + " data.bin"         # a space: the relative URI is percent-encoded
        with open(os.path.join(os.path.dirname(path), name), "wb") as f:
            f.write(blob)
        doc["buffers"][0]["uri"] = name.replace(" ", "%20")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return path


def _indexed(mesh):
    vid = mesh.vertex_ids.reshape(-1)
    nv = int(vid.max()) + 1
    pts = np.zeros((nv, 3), np.float32); nrm = np.zeros((nv, 3), np.float32)
    pts[vid] = mesh.positions.reshape(-1, 3); nrm[vid] = mesh.normals.reshape(-1, 3)
    return pts, nrm, vid.astype(np.uint32)


def sample_meshes():
    """Every situation the engine's loader meets: several triangle-list primitives in one mesh (the 'triangles first' order), mixed draw
    modes with a line loop, a primitive without normals, one without indices, u16 and u32 indices, an omitted mode, a skinned and a
    morphed primitive (left out of the tree), a mesh made only of such primitives."""
    tp, tn, ti = _indexed(scenes.torus(20, 10))
    sp, sn, si = _indexed(scenes.uv_sphere(12, 9))
    bp, bn, bi = _indexed(scenes.box_mesh(1.0, 2.0, 3.0, sub=3))
    strip = (np.arange(30, dtype=np.uint32) * 7) % len(sp)
    return [
        [dict(points=tp, normals=tn, indices=ti.astype(np.uint16), mode=4), dict(points=sp + 3.0, normals=sn, indices=si),
         dict(points=bp - 2.0, normals=None, indices=bi.astype(np.uint16)), dict(points=sp * 0.5, normals=sn, indices=si[:90].astype(np.uint16), mode=4)],
        [dict(points=sp, normals=sn, indices=strip, mode=5), dict(points=tp, normals=tn, indices=ti[:300], mode=4),
         dict(points=sp, normals=None, indices=strip.astype(np.uint16), mode=2), dict(points=bp, normals=bn, indices=bi, mode=None),
         dict(points=sp, normals=sn, indices=strip[:20], mode=6), dict(points=tp[:40], normals=tn[:40], indices=np.arange(40, dtype=np.uint16), mode=1),
         dict(points=tp[:9], normals=None, indices=np.arange(9, dtype=np.uint32), mode=0), dict(points=sp, normals=sn, indices=strip, mode=3)],
        [dict(points=bp, normals=bn, indices=bi, mode=4, skinned=True), dict(points=tp, normals=tn, indices=ti, mode=4),
         dict(points=sp, normals=sn, indices=si, mode=4, morph=True)],
        [dict(points=bp, normals=bn, indices=bi, mode=4, skinned=True)],
        [dict(points=tp[ti][:120], normals=tn[ti][:120], indices=None, mode=4)],
    ]
