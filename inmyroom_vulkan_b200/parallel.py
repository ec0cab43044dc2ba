"""Multi-GPU plumbing of the collision frame: one process per GPU, the broad-phase sweep sharded by chunk, ONE exchange per frame.

The path shards naturally: trees and the entry table are replicated, every rank runs the (cheap) sort
of the broad phase and then sweeps only every world-th chunk of it (imrcd_frame_set_shard; a chunk is one entity x 512
consecutive candidates of its window), so each candidate pair, all of its triangle hits and its contact reduction belong
to exactly one rank and nothing crosses GPUs before the end of the frame.  The only collective is the
end-of-frame merge of the colliding-pair records (80 B each): ONE all-gather of fixed-capacity blocks whose header row carries
the rank's record count, taken straight from the library's result block in HBM, over NCCL on NVLink (gloo on CPU in the tests).

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


class _DevMem:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class FrameGather:
    """End-of-frame merge of the colliding entity pairs of all ranks: ONE all-gather of fixed-capacity blocks whose first
    row carries the rank's record count (no separate count exchange, no host round trip before the payload moves).
    The capacity doubles (and the gather is repeated) on the rare frame where some rank outgrows it."""

    def __init__(self, cd, world: int, rank: int, group=None, capacity: int = 256):
        self.cd = cd; self.world = world; self.rank = rank; self.group = group
        self.cap = capacity
        self.last = None
        self._send = None; self._recv = None
        self._hdr_event = None; self._cap_alloc = capacity

    def _local_device_records(self) -> torch.Tensor:
        ctx = self.cd.ctx
        dp = C.c_void_p(); n = C.c_uint64(); dh = C.c_void_p(); nh = C.c_uint64()
        ctx.check(ctx.lib.imrcd_frame_results_device(ctx.h, C.byref(dp), C.byref(n), C.byref(dh), C.byref(nh)))
        if n.value == 0 or not dp.value:
            return torch.zeros((0, RECORD_BYTES), dtype=torch.uint8, device="cuda")
        t = torch.as_tensor(_DevMem(dp.value, n.value * RECORD_BYTES), device="cuda")
        return t.view(n.value, RECORD_BYTES)

    def _buffers(self, device):
        if self._send is None or self._send.shape[0] != self.cap + 1 or self._send.device != device:
            self._send = torch.zeros((self.cap + 1, RECORD_BYTES), dtype=torch.uint8, device=device)
            self._recv = torch.empty((self.world * (self.cap + 1), RECORD_BYTES), dtype=torch.uint8, device=device)
        return self._send, self._recv

    def exchange(self, local: torch.Tensor):
        """local: (n, 80) uint8 records of this rank (CUDA over NCCL, CPU over gloo).  Returns (blocks, counts) where
        blocks is (world, cap + 1, 80) on the device and counts the per-rank record counts on the host."""
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
            self.cap = 1 << (max(counts) - 1).bit_length()      # somebody else outgrew the blocks: everyone retries

    # ---- device path: the library keeps the records behind a header row (imrcd_frame_results_block), so the frame's block goes into the
    #      collective as it lies in HBM: no count exchange, no staging copy, no host round trip before or after the payload moves ----
    def _block_view(self):
        ctx = self.cd.ctx
        dp = C.c_void_p(); n = C.c_uint64(); cap = C.c_uint64()
        ctx.check(ctx.lib.imrcd_frame_results_block(ctx.h, C.byref(dp), C.byref(n), C.byref(cap)))
        if not dp.value:
            raise RuntimeError("imrcd_frame_results_block: no result block (run a frame first)")
        self.cap = min(self.cap, int(cap.value))
        key = (dp.value, self.cap)
        if getattr(self, "_view_key", None) != key:
            self._view = torch.as_tensor(_DevMem(dp.value, (self.cap + 1) * RECORD_BYTES), device="cuda").view(self.cap + 1, RECORD_BYTES)
            self._view_key = key
            self._recv = torch.empty((self.world * (self.cap + 1), RECORD_BYTES), dtype=torch.uint8, device="cuda")
            self._hdr = torch.empty((self.world, 8), dtype=torch.uint8).pin_memory()
        return self._view, int(n.value), int(cap.value)

    def gather_device(self) -> torch.Tensor:
        """Call after cd.run(): every rank ends up with all ranks' blocks in HBM (ONE collective, nothing else on the stream).
        The per-rank counts travel in the blocks' header rows; counts() reads them (and tells when the capacity was too small)."""
        send, n, cap_alloc = self._block_view()
        dist.all_gather_into_tensor(self._recv, send, group=self.group)
        blocks = self._recv.view(self.world, self.cap + 1, RECORD_BYTES)
        self._hdr.copy_(blocks[:, 0, :8], non_blocking=True)
        self._hdr_event = torch.cuda.Event(); self._hdr_event.record()
        self.last = (blocks, None)
        self._cap_alloc = cap_alloc
        return blocks

    def run_and_gather_device(self) -> torch.Tensor:
        """One frame and its merge with a single host wait: the frame's kernels are enqueued (imrcd_frame_run_async), the collective goes
        onto the stream right behind them - the block's header row is written on the device, so nothing about the frame has to be
        known on the host yet - and only then does the host wait (imrcd_frame_finish).  On the rare frame where a buffer overflowed and
        the library ran the frame again, the collective is repeated on the new block."""
        self.cd.run_async()
        blocks = self.gather_device()
        if self.cd.finish():
            blocks = self.gather_device()
        return blocks

    def counts(self):
        """Per-rank record counts of the last gather_device(); None when some rank had more records than the capacity (the capacity
        is raised for the next gather, every rank sees the same headers and decides alike)."""
        if self._hdr_event is None:
            raise RuntimeError("FrameGather.counts() before gather_device()")
        self._hdr_event.synchronize()
        counts = self._hdr.view(torch.int64).reshape(-1).tolist()
        if max(counts) > self.cap:
            self.cap = min(1 << (max(counts) - 1).bit_length(), self._cap_alloc)
            return None
        self.last = (self.last[0], counts)
        return counts

    def execute_host(self) -> np.ndarray:
        """ExecuteCollisionDetection for N ranks, entries already added: upload, the frame, the merge and ONE host wait; returns the
        merged records of all ranks as a numpy structured array (what gather_host() gives after cd.ExecuteCollisionDetection())."""
        self.cd.upload()
        blocks = self.run_and_gather_device()
        while True:
            counts = self.counts()
            if counts is not None:
                break
            blocks = self.gather_device()
        h = blocks.cpu().numpy()
        parts = [h[r, 1:1 + c].reshape(-1) for r, c in enumerate(counts) if c]
        return np.concatenate(parts).view(PAIR_DTYPE) if parts else np.zeros(0, PAIR_DTYPE)

    def gather_host(self) -> np.ndarray:
        """Call after cd.ExecuteCollisionDetection(): merged records as a numpy structured array."""
        while True:
            blocks = self.gather_device()
            counts = self.counts()
            if counts is not None:
                break
        h = blocks.cpu().numpy()
        parts = [h[r, 1:1 + c].reshape(-1) for r, c in enumerate(counts) if c]
        return np.concatenate(parts).view(PAIR_DTYPE) if parts else np.zeros(0, PAIR_DTYPE)
