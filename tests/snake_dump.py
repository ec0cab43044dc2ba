"""Parser of the frame dumps written by tests/cpp/snake_harness.cpp (SNAKE_DUMP=<file>, shadow build): the meshes of the scene, every
frame's collision entries as ModelCollisionComp::Update made them, and the reference engine's verdict (colliding (entity, other) rows with
their deltaVectors).  Floats are C hex floats (%a): bit-exact.  Test infrastructure."""
import gzip

import numpy as np

from inmyroom_vulkan_b200 import scenes


def _f(tok):
    return float.fromhex(tok)


def load(path):
    """-> (meshes: [(points (n,3), normals (n,3), indices u32)], frames: [dict(entity, mesh, callback, cur (n,16), prev (n,16), callbacks [(family, other, delta(3))])])"""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as f:
        lines = f.read().split("\n")
    meshes, frames, i = [], [], 0
    while i < len(lines):
        w = lines[i].split()
        if not w:
            i += 1; continue
        if w[0] == "mesh":
            npts, nidx = int(w[1]), int(w[2])
            vals = np.array([[_f(t) for t in lines[i + 1 + k].split()] for k in range(npts)], np.float64).astype(np.float32)
            i += 1 + npts
            idx = []
            while len(idx) < nidx:
                idx += [int(t) for t in lines[i].split()]; i += 1
            meshes.append((vals[:, :3].copy(), vals[:, 3:].copy(), np.array(idx, np.uint32)))
        elif w[0] == "frame":
            frames.append(dict(entity=[], mesh=[], callback=[], cur=[], prev=[], callbacks=[])); i += 1
        elif w[0] == "entry":
            fr = frames[-1]
            fr["entity"].append(int(w[1])); fr["mesh"].append(int(w[2])); fr["callback"].append(int(w[3]))
            m = [_f(t) for t in w[4:36]]
            fr["cur"].append(m[:16]); fr["prev"].append(m[16:]); i += 1
        elif w[0] == "callback":
            frames[-1]["callbacks"].append((int(w[1]), int(w[2]), np.array([_f(t) for t in w[3:6]], np.float32))); i += 1
        else:
            i += 1
    for fr in frames:
        fr["entity"] = np.array(fr["entity"], np.uint32); fr["mesh"] = np.array(fr["mesh"], np.uint32); fr["callback"] = np.array(fr["callback"], np.uint8)
        fr["cur"] = np.array(fr["cur"], np.float32).reshape(-1, 16); fr["prev"] = np.array(fr["prev"], np.float32).reshape(-1, 16)
    return meshes, frames


def scene_of(meshes_flat, fr):
    return scenes.Scene(meshes=meshes_flat, mesh_index=fr["mesh"], matrices=fr["cur"], should_callback=fr["callback"], entities=fr["entity"], previous=fr["prev"])
