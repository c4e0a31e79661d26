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

    The per-rank batch must be equal on all ranks (pad the last shard), which makes the exchange one
    `all_gather_into_tensor` per tensor — latency-bound (tens of microseconds), enqueued on the compute stream."""

    def __init__(self, b_local: int, max_det: int, device, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.out_all = torch.zeros((self.world * b_local, max_det, 6), device=device, dtype=torch.float32)
        self.counts_all = torch.zeros((self.world * b_local,), device=device, dtype=torch.int32)
        self.b_local = b_local

    def gather(self, out: torch.Tensor, counts: torch.Tensor):
        if self.world == 1:
            self.out_all.copy_(out)
            self.counts_all.copy_(counts)
        elif dist.get_backend(self.group) == "gloo":      # gloo lacks all_gather_into_tensor on some builds
            outs = list(self.out_all.chunk(self.world))
            cnts = list(self.counts_all.chunk(self.world))
            dist.all_gather(outs, out.contiguous(), group=self.group)
            dist.all_gather(cnts, counts.contiguous(), group=self.group)
        else:
            dist.all_gather_into_tensor(self.out_all, out.contiguous(), group=self.group)
            dist.all_gather_into_tensor(self.counts_all, counts.contiguous(), group=self.group)
        return self.out_all, self.counts_all


def split_detections(out_all: torch.Tensor, counts_all: torch.Tensor, n_items: Optional[int] = None) -> list[torch.Tensor]:
    """Padded (B, max_det, 6) + counts -> list of (k_i, 6) tensors (the reference's NMS output layout)."""
    counts = counts_all.tolist()
    n = len(counts) if n_items is None else n_items
    return [out_all[i, : counts[i]] for i in range(n)]
