"""Batch sharding across the GPUs of one box and the gather of detections to rank 0.

The reference has no multi-GPU predict: `select_device("0,1")` just returns cuda:0.  Images are
independent through conv, decode and NMS, so each rank runs the whole graph on a contiguous slice of the batch with no
data-path collective; the one exchange step is the gather of the padded detections
(B_local, max_det, 6) fp32 + counts (B_local,) int32 — 7.2 KB per image — over NCCL (NVLink 5 / NVSwitch).
Works with any torch.distributed backend (gloo on CPU is used by the tests).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split; the first `n_items % world_size` ranks get one extra item."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    s, e = shard_bounds(batch.shape[0], world_size, rank)
    return batch[s:e]


class DetectionGather:
    """Pre-allocated all-gather of (out, counts); every rank calls `gather`, rank `dst` reads `out_all/counts_all`.

    The per-rank batch must be equal on all ranks (pad the last shard).  ONE collective per step: the counts travel as an
    extra row behind each image's max_det rows (exact in fp32 below 2^24), so the exchange is a single
    `all_gather_into_tensor` of (B_local, max_det + 1, 6) - latency-bound (tens of microseconds), enqueued on the compute
    stream."""

    def __init__(self, b_local: int, max_det: int, device, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.key = (b_local, max_det)
        self.b_local, self.max_det = b_local, max_det
        self.packed = torch.zeros((b_local, max_det + 1, 6), device=device, dtype=torch.float32)
        self.packed_all = torch.zeros((self.world * b_local, max_det + 1, 6), device=device, dtype=torch.float32)
        self.out_all = self.packed_all[:, :max_det]                       # views: (world * B_local, max_det, 6)
        self.counts_all = torch.zeros((self.world * b_local,), device=device, dtype=torch.int32)

    def gather(self, out: torch.Tensor, counts: torch.Tensor):
        self.packed[:, : self.max_det].copy_(out)
        self.packed[:, self.max_det, 0].copy_(counts)                     # int32 -> fp32, exact
        if self.world == 1:
            self.packed_all.copy_(self.packed)
        elif dist.get_backend(self.group) == "gloo":                      # gloo lacks all_gather_into_tensor on some builds
            dist.all_gather(list(self.packed_all.chunk(self.world)), self.packed, group=self.group)
        else:
            dist.all_gather_into_tensor(self.packed_all, self.packed, group=self.group)
        self.counts_all.copy_(self.packed_all[:, self.max_det, 0])        # fp32 -> int32
        return self.out_all, self.counts_all


def split_detections(out_all: torch.Tensor, counts_all: torch.Tensor, n_items: Optional[int] = None) -> list[torch.Tensor]:
    """Padded (B, max_det, 6) + counts -> list of (k_i, 6) tensors (the reference's NMS output layout)."""
    counts = counts_all.tolist()
    n = len(counts) if n_items is None else n_items
    return [out_all[i, : counts[i]] for i in range(n)]


def shard_of(i: int, n_items: int, world_size: int) -> tuple[int, int]:
    """(rank, position inside that rank's shard) of item i under `shard_bounds`."""
    for r in range(world_size):
        s, e = shard_bounds(n_items, world_size, r)
        if s <= i < e:
            return r, i - s
    raise ValueError(f"item {i} outside [0, {n_items})")
