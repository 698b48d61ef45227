"""Batch-sharded data parallelism for the drop-in model.

The graph (N <= 8192 nodes) is too small to shard and the recurrence forbids time sharding,
so every rank holds a full replica, takes ``B / world`` samples of each global batch, and
the ONLY collective per step is one all-reduce (sum) of one flat fp32 gradient bucket over
the intra-box NVLink/NVSwitch communicator.

``FlatGradBucket`` makes every ``param.grad`` a view into one contiguous buffer, so autograd
accumulates straight into the bucket and the all-reduce needs no gather/scatter copies.
Parameters that receive no gradient in a configuration (``node_vec1/2`` under bidirection,
unused ``weights_g``) keep zero-filled slots, identically on all ranks.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        """Replaces optimizer.zero_grad(): one memset, and re-pins the views in case a caller
        set ``p.grad = None``."""
        self.flat.zero_()
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce_mean(self, group: Optional[dist.ProcessGroup] = None):
        """Sum over ranks, then divide by the world size (each rank's loss is a mean over its
        shard, so the mean of shard gradients is the global-batch gradient when the loss mask
        density is equal across shards - SURVEY.md a16)."""
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(group)
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.mul_(1.0 / world)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Global L2 clip over the bucket (same set of parameters the reference's
        ``clip_grad_norm_(model.parameters(), 5)`` covers; executor:420-421)."""
        norm = torch.linalg.vector_norm(self.flat)
        scale = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        self.flat.mul_(scale)
        return norm


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Make every rank start from rank ``src``'s weights."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Rank r takes B/world consecutive samples of the global batch."""
    out = {}
    for k, v in batch.items():
        b = v.shape[0]
        if b % world != 0:
            raise ValueError("global batch %d is not divisible by world size %d" % (b, world))
        per = b // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def train_step(model, batch, optimizer, bucket: FlatGradBucket, max_grad_norm: Optional[float] = 5.0):
    """One step of ``TrafficStateExecutor._train_epoch`` (executor:413-422) with the
    gradient all-reduce inserted before clipping: zero_grad -> calculate_loss -> backward ->
    all-reduce -> clip_grad_norm_ -> optimizer.step.  Returns the (device) loss tensor."""
    bucket.zero()
    loss = model.calculate_loss(batch)
    loss.backward()
    bucket.all_reduce_mean()
    if max_grad_norm is not None:
        bucket.clip_grad_norm_(max_grad_norm)
    optimizer.step()
    return loss.detach()
