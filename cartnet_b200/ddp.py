"""Data-parallel plumbing for the one place the hot path shards: whole crystals per
rank, balanced by EDGE count, and ONE all-reduce of the flat fp32 gradient per optimiser step
(NCCL over NVLink on the GPU box, gloo in the CPU tests). The reference has no distributed code
(it runs independent seeds per GPU, /root/reference/scripts/train_cartnet_adp.sh:3-14); BatchNorm
statistics stay per-rank, exactly what per-GPU batches in the reference would see.
"""
from __future__ import annotations

from typing import List, Sequence

import weakref

import torch
import torch.distributed as dist


def shard_by_edges(edge_counts: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time greedy partition of crystals over ranks by edge count (the layer cost is
    linear in edges). Deterministic: ties broken by crystal index. Returns per-rank sorted index lists."""
    order = sorted(range(len(edge_counts)), key=lambda i: (-int(edge_counts[i]), i))
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (loads[q], q))
        parts[r].append(i)
        loads[r] += int(edge_counts[i])
    return [sorted(p) for p in parts]


class FlatGradAllReduce:
    """Keeps every parameter's .grad as a view into one flat buffer so that the gradient exchange is a
    single collective (2,498,438 floats = 9.99 MB for the default CartNet)."""

    def __init__(self, params, group=None, direct: bool = False):
        """direct=True additionally lets the hand-written backward kernels store a parameter's gradient straight
        into its slice of the flat buffer (no AccumulateGrad `+=` launch per parameter). The caller promises what
        a plain training loop does anyway: zero() before every backward. A second backward without zero() in between
        (gradient accumulation over micro-batches) is detected per parameter and falls back to autograd's accumulation,
        and the shortcut ends with this object's lifetime."""
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.direct = bool(direct)
        self.epoch = 0                     # bumped by zero(): a parameter is written directly at most once per epoch
        for p in self.params:
            p._cn_direct_owner = weakref.ref(self) if self.direct else None
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.epoch += 1
        self.flat.zero_()
        off = 0
        for p in self.params:   # re-attach in case an optimiser / zero_grad(set_to_none=True) dropped the views
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + 1].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def allreduce_mean(self):
        if not dist.is_available() or not dist.is_initialized():
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None):
    """Parameters and buffers (BatchNorm running statistics included) from rank `src` to every rank."""
    if not dist.is_available() or not dist.is_initialized():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
