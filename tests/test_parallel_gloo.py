"""CPU, world_size 2 over gloo: the N>1 host logic -- the end-of-frame variable-length gather reassembles every rank's
records in rank order (that the device-side shards partition the pair list is a GPU test: test_gpu_frame.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inmyroom_vulkan_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = [7, 0][rank] if world == 2 else int(rng.integers(0, 9))
        rec = np.zeros(n, parallel.PAIR_DTYPE)
        rec["entry_first"] = rank * 1000 + np.arange(n); rec["entity_second"] = 77 + rank; rec["avg_first"] = rng.random((n, 3))
        local = torch.from_numpy(rec.view(np.uint8).reshape(n, parallel.RECORD_BYTES).copy())
        merged = parallel.all_gather_varlen(local)
        np.save(os.path.join(out_dir, f"local{rank}.npy"), rec)
        np.save(os.path.join(out_dir, f"merged{rank}.npy"), merged.numpy())
        empty = parallel.all_gather_varlen(torch.zeros((0, parallel.RECORD_BYTES), dtype=torch.uint8))
        assert empty.shape == (0, parallel.RECORD_BYTES)
        # the frame gather: one fixed-capacity collective, capacity 4 so that rank 0's 7 records force the grow-and-retry path
        fg = parallel.FrameGather(None, world, rank, capacity=4)
        blocks, counts = fg.exchange(local)
        got = torch.cat([blocks[r, 1:1 + c] for r, c in enumerate(counts)], 0)
        np.save(os.path.join(out_dir, f"fg{rank}.npy"), got.numpy())
        assert fg.cap >= 7
    finally:
        dist.destroy_process_group()


def test_all_gather_varlen_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    locals_ = [np.load(tmp_path / f"local{r}.npy") for r in range(world)]
    expect = np.concatenate(locals_).view(np.uint8).reshape(-1, parallel.RECORD_BYTES)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"merged{r}.npy"), expect)
        assert np.array_equal(np.load(tmp_path / f"fg{r}.npy"), expect)
