"""Multi-GPU plumbing of the collision frame: one process per GPU, the frame sharded by entity, ONE exchange per frame.

The path shards naturally: trees are replicated, every rank is handed the whole entry list and keeps its share
(imrcd_frame_set_shard: all entries with shouldCallback + its blocks of the others), so each candidate pair, all of its triangle
hits, its contact reduction and its response belong to exactly one rank and nothing crosses GPUs before the end of the frame.
The only collective is the end-of-frame merge of the colliding-pair records (80 B each): ONE all-gather of fixed-capacity blocks
whose header row carries the rank's record count and overflow bits.  It lives INSIDE libimrcd.so (csrc/imrcd_comm.cu: stores into the
peers' buffers over NVLink + arrival flags where the GPUs reach each other's memory, else ncclAllGather, on the frame's own stream); what is left here is the rendezvous (init_comm: the NCCL unique id travels over torch.distributed) and
a CPU model of the same block protocol over gloo for the world-size-2 host-logic tests.

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



def init_comm(ctx, rank: int, world: int, group=None):
    """Attach an NCCL communicator to a library context (imrcd_comm_init), one process per GPU: rank 0 makes the unique id, it is
    broadcast over the already initialised torch.distributed group, every rank joins.  From then on imrcd_frame_run / run_async + finish /
    execute of this context end with the in-library all-gather and imrcd_frame_results returns the merged records of all ranks."""
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).clone()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    uid = uid.to(dev)
    dist.broadcast(uid, src=0, group=group)
    ctx.comm_init(uid.cpu().numpy().tobytes(), rank, world)


class FrameGather:
    """CPU model of the library's end-of-frame merge (csrc/imrcd_comm.cu) for the gloo tests: ONE all-gather of fixed-capacity blocks
    whose first row carries the rank's record count; the capacity doubles (and the gather is repeated) when some rank outgrows it, a
    decision every rank takes from the same gathered headers."""

    def __init__(self, cd, world: int, rank: int, group=None, capacity: int = 256):
        self.cd = cd; self.world = world; self.rank = rank; self.group = group
        self.cap = capacity
        self._send = None; self._recv = None

    def _buffers(self, device):
        if self._send is None or self._send.shape[0] != self.cap + 1 or self._send.device != device:
            self._send = torch.zeros((self.cap + 1, RECORD_BYTES), dtype=torch.uint8, device=device)
            self._recv = torch.empty((self.world * (self.cap + 1), RECORD_BYTES), dtype=torch.uint8, device=device)
        return self._send, self._recv

    def exchange(self, local: torch.Tensor):
        """local: (n, 80) uint8 records of this rank.  Returns (blocks, counts) where blocks is (world, cap + 1, 80) and counts the
        per-rank record counts on the host."""
        n = int(local.shape[0])
        while True:          # every rank uses the same capacity in every collective: it only changes from the gathered counts
            send, recv = self._buffers(local.device)
            send[0, :8] = torch.tensor([n], dtype=torch.int64).view(torch.uint8).to(local.device, non_blocking=True)
            k = min(n, self.cap)
            if k:
                send[1:k + 1] = local[:k]
            dist.all_gather_into_tensor(recv, send, group=self.group)
            blocks = recv.view(self.world, self.cap + 1, RECORD_BYTES)
            counts = blocks[:, 0, :8].contiguous().view(torch.int64).reshape(-1).cpu().tolist()
            if max(counts) <= self.cap:
                return blocks, counts
            self.cap = 1 << (max(counts) - 1).bit_length()      # somebody outgrew the blocks: everyone retries
