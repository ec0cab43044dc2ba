"""Multi-GPU plumbing of the collision frame: one process per GPU, pairs sharded by entity, ONE exchange per frame.

The path shards naturally: trees and the entry table are replicated, every rank runs the (cheap) sort
of the broad phase and keeps only the pairs whose owner entry -- the larger entry index of the pair -- satisfies
owner % world == rank (imrcd_frame_set_shard), so each colliding entity pair and all of its triangle hits are produced
by exactly one rank and the per-pair contact reduction needs no cross-GPU step.  The only collective is the
end-of-frame merge of the colliding-pair records (80 B each): an all-gather of per-rank counts followed by an
all-gather of max-padded record blocks, over NCCL on NVLink (gloo on CPU in the tests).

The reference has no counterpart (it is a single-threaded host loop, CollisionDetection.cpp:44-129); what is kept is
its contract: after ExecuteCollisionDetection every consumer sees the complete colliding set of the frame.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from .collision import PAIR_DTYPE

RECORD_BYTES = PAIR_DTYPE.itemsize     # 80


def owner_of_pairs(pairs: np.ndarray) -> np.ndarray:
    """Shard key of broad-phase pairs: the larger entry index (k_sweep, csrc/imrcd_frame.cu)."""
    pairs = np.asarray(pairs).reshape(-1, 2)
    return np.maximum(pairs[:, 0], pairs[:, 1])


def shard_mask(pairs: np.ndarray, rank: int, world: int) -> np.ndarray:
    return (owner_of_pairs(pairs) % np.uint32(world)) == np.uint32(rank)


def all_gather_varlen(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather rows of a (n_i, W) uint8 tensor whose n_i differs per rank; returns the (sum n_i, W) concatenation in
    rank order on every rank.  Works for CUDA tensors over NCCL and CPU tensors over gloo."""
    world = dist.get_world_size(group)
    assert local.dtype == torch.uint8 and local.dim() == 2
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts_h = counts.cpu().tolist()
    mx = max(counts_h)
    if mx == 0:
        return local.new_zeros((0, local.shape[1]))
    padded = local.new_zeros((mx, local.shape[1]))
    padded[: local.shape[0]] = local
    out = local.new_empty((world * mx, local.shape[1]))
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view(world, mx, local.shape[1])
    return torch.cat([out[r, : counts_h[r]] for r in range(world)], 0)


class _DevMem:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class FrameGather:
    """End-of-frame merge of the colliding entity pairs of all ranks."""

    def __init__(self, cd, world: int, rank: int, group=None):
        self.cd = cd; self.world = world; self.rank = rank; self.group = group
        self.last = None

    def _local_device_records(self) -> torch.Tensor:
        ctx = self.cd.ctx
        dp = C.c_void_p(); n = C.c_uint64(); dh = C.c_void_p(); nh = C.c_uint64()
        ctx.check(ctx.lib.imrcd_frame_results_device(ctx.h, C.byref(dp), C.byref(n), C.byref(dh), C.byref(nh)))
        if n.value == 0 or not dp.value:
            return torch.zeros((0, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        t = torch.as_tensor(_DevMem(dp.value, n.value * RECORD_BYTES), device="cuda")
        return t.view(n.value, RECORD_BYTES)

    def gather_device(self) -> torch.Tensor:
        """Call after cd.run(): every rank ends up with all ranks' records in HBM."""
        self.last = all_gather_varlen(self._local_device_records(), self.group)
        return self.last

    def gather_host(self) -> np.ndarray:
        """Call after cd.ExecuteCollisionDetection(): merged records as a numpy structured array."""
        g = self.gather_device()
        return g.cpu().numpy().reshape(-1).view(PAIR_DTYPE) if g.shape[0] else np.zeros(0, PAIR_DTYPE)
