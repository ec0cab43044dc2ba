"""GPU: the end-of-frame merge inside the library (csrc/imrcd_comm.cu) and the entity-sharded frame.

  * a communicator of ONE rank (any GPU box): the gather path (ncclAllGather on the frame's stream, compaction, speculative D2H,
    the collective retry decision) gives the records the plain context gives, over changing frames;
  * one process, two GPUs (imrcd_group_*) and two processes, one GPU each (imrcd_comm_init, rendezvous over torch.distributed):
    merged records == the single-GPU frame, with entries that change from frame to frame, a frame that outgrows the gather blocks
    (capacity raised collectively) and a frame with fewer than two entries.  Skipped on a box with one GPU.  Both transports of the
    merge are run: peer memory over NVLink (k_p2p_push / k_p2p_wait_compact, the default where the GPUs reach each other) and the
    ncclAllGather it replaces (IMRCD_P2P=0).
"""
import os
import socket

import numpy as np
import pytest

from helpers import same_entity_pairs
from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import CollisionDetection, Context, Group, OBBtree

pytestmark = pytest.mark.gpu


def _scene(n_bodies, seed):
    static = scenes.atrium_static(detail=1)
    keep = list(range(0, 8)) + list(range(60, 70))
    static = ([static[0][i] for i in keep], static[1][keep])
    return scenes.scene_static_vs_bodies(scenes.uv_sphere(24, 17), n_bodies, seed=seed, body_scale=(0.5, 1.5), static=static)


def _frames():
    """(scene, previous or None): different entries every frame; frame 2 has > 1024 colliding pairs per rank at N=2 (the gather blocks
    start at 1024 rows), frame 3 moves (response stage), frame 4 is smaller again."""
    out = []
    for k, (n, seed) in enumerate(((1500, 5), (2500, 6), (60000, 7), (1200, 8), (900, 9))):
        sc = _scene(n, seed)
        prev = None
        if k == 3:
            prev = sc.matrices.copy()
            prev[18:, 12:15] += (np.random.default_rng(1).normal(size=(sc.n_entries - 18, 3)) * 0.02).astype(np.float32)
        out.append((sc, prev))
    return out


def _key(e):
    return np.lexsort((e["entry_second"], e["entry_first"]))


def _single_gpu_reference(ctx, frames):
    cd = CollisionDetection(ctx=ctx)
    trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in frames[0][0].meshes]
    res = []
    for sc, prev in frames:
        ids = np.array([trees[m].mesh_id for m in sc.mesh_index], np.uint32)
        cd.Reset(); cd.add_entries(sc.matrices, ids, sc.should_callback, sc.entities, prev); cd.ExecuteCollisionDetection()
        ep, _ = cd.results(want_hits=False)
        res.append(ep[_key(ep)])
    return res


def _same(a, b):
    same_entity_pairs(a, b)
    for f in ("delta_first", "delta_second"):
        x = np.asarray(a[f], np.float64); y = np.asarray(b[f], np.float64)
        ok = np.linalg.norm(x - y, axis=1) <= 1e-4 * np.maximum(np.linalg.norm(y, axis=1), 1e-30) + 1e-9
        ok |= np.isnan(x).any(1) & np.isnan(y).any(1)
        assert ok.all(), f


def test_comm_of_one_rank_gives_the_plain_result(gpu_ctx):
    frames = _frames()
    want = _single_gpu_reference(gpu_ctx, frames)
    assert len(want[2]) > 1024 and len(want[0]) > 20
    ctx = Context(0)
    ctx.comm_init(ctx.comm_unique_id(), 0, 1)
    cd = CollisionDetection(ctx=ctx)
    trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in frames[0][0].meshes]
    for (sc, prev), w in zip(frames, want):
        ids = np.array([trees[m].mesh_id for m in sc.mesh_index], np.uint32)
        cd.Reset(); cd.add_entries(sc.matrices, ids, sc.should_callback, sc.entities, prev)
        cd.upload(); cd.run_async(); cd.finish(); cd.fetch()                    # the asynchronous route: the collective sits behind the kernels
        ep, _ = cd.results(want_hits=False)
        _same(w, ep[_key(ep)])
        st = cd.stats()
        assert st["n_merged"] == len(w) == st["n_colliding"]
        loc = cd.results_local()
        _same(w, loc[_key(loc)])
    # fewer than two entries: nothing runs, nothing is reported (CollisionDetection.cpp:40)
    sc = frames[0][0]
    cd.Reset(); cd.add_entries(sc.matrices[:1], np.array([trees[sc.mesh_index[0]].mesh_id], np.uint32), sc.should_callback[:1], sc.entities[:1])
    cd.upload(); cd.run(); cd.fetch()
    assert len(cd.results(want_hits=False)[0]) == 0
    ctx.comm_destroy()
    ctx.close()


def _need_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def _peer_access():
    import torch
    return torch.cuda.can_device_access_peer(0, 1) and torch.cuda.can_device_access_peer(1, 0)


@pytest.mark.parametrize("p2p", [1, 0])
def test_group_of_two_gpus_in_one_process(gpu_ctx, monkeypatch, p2p):
    _need_two_gpus()
    monkeypatch.setenv("IMRCD_P2P", str(p2p))
    frames = _frames()
    want = _single_gpu_reference(gpu_ctx, frames)
    g = Group([0, 1])
    ids_of = [g.mesh_create(m.positions, m.normals, m.vertex_ids) for m in frames[0][0].meshes]
    for (sc, prev), w in zip(frames, want):
        ids = np.array([ids_of[m] for m in sc.mesh_index], np.uint32)
        g.Reset(); g.add_entries(sc.matrices, ids, sc.should_callback, sc.entities, prev); g.ExecuteCollisionDetection()
        ep = g.results()
        _same(w, ep[_key(ep)])
    assert g.comm_transport() == (2 if (p2p and _peer_access()) else 1)
    sc = frames[0][0]                                                              # one entry: no frame
    g.Reset(); g.add_entries(sc.matrices[:1], np.array([ids_of[sc.mesh_index[0]]], np.uint32), sc.should_callback[:1], sc.entities[:1]); g.ExecuteCollisionDetection()
    assert len(g.results()) == 0
    g.close()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from inmyroom_vulkan_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = Context(rank)
        parallel.init_comm(ctx, rank, world)
        cd = CollisionDetection(ctx=ctx)
        frames = _frames()
        trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in frames[0][0].meshes]
        for k, (sc, prev) in enumerate(frames):
            ids = np.array([trees[m].mesh_id for m in sc.mesh_index], np.uint32)
            cd.Reset(); cd.add_entries(sc.matrices, ids, sc.should_callback, sc.entities, prev)
            cd.upload(); cd.run_async(); cd.finish(); cd.fetch()
            ep, _ = cd.results(want_hits=False)
            st = cd.stats()
            assert st["n_entries_local"] < st["n_entries"]                         # this rank kept its share of the entries only
            np.save(os.path.join(out_dir, f"r{rank}_f{k}.npy"), ep)
            np.save(os.path.join(out_dir, f"r{rank}_f{k}_local.npy"), cd.results_local())
        with open(os.path.join(out_dir, f"r{rank}_transport.txt"), "w") as f:
            f.write(str(ctx.comm_transport()))
        dist.barrier()
        ctx.comm_destroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [1, 0])
def test_two_processes_one_gpu_each(gpu_ctx, tmp_path, monkeypatch, p2p):
    _need_two_gpus()
    monkeypatch.setenv("IMRCD_P2P", str(p2p))                                      # the workers are spawned with this environment
    import torch.multiprocessing as mp
    frames = _frames()
    want = _single_gpu_reference(gpu_ctx, frames)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert int((tmp_path / f"r{r}_transport.txt").read_text()) == (2 if (p2p and _peer_access()) else 1)
    for k, w in enumerate(want):
        merged = [np.load(tmp_path / f"r{r}_f{k}.npy") for r in range(2)]
        assert np.array_equal(merged[0].view(np.uint8), merged[1].view(np.uint8)), "the ranks disagree on the merged records"
        _same(w, merged[0][_key(merged[0])])
        local = [np.load(tmp_path / f"r{r}_f{k}_local.npy") for r in range(2)]
        assert len(local[0]) + len(local[1]) == len(w) and min(len(local[0]), len(local[1])) > 0.3 * len(w)     # disjoint and roughly even
