"""Data-parallel plumbing (one process per GPU, ``torch.distributed``).

The reference has no distributed code at all.  The path shards in exactly one way — utterances are
independent (src/models/adapter_router.py:610-622 is an independent loop) — so:

* inference: the batch is split by utterance index into contiguous shards, base weights / adapter stacks / LID head
  replicated; NO collective on the data path (``shard_range``).
* LoRA training (reference trainer, src/training/trainer.py:251-277): each rank computes its local dA / dB (K3) into
  ONE flat fp32 bucket; a single in-place NCCL all-reduce over NVLink sums it, then it is divided by the world size
  and the gradient-norm clip runs once over the flat buffer (``FlatGradBucket``) instead of a per-parameter loop.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [start, end) of ``n_items`` utterances for ``rank`` (sizes differ by at most 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, world_size: Optional[int] = None, rank: Optional[int] = None) -> torch.Tensor:
    """This rank's utterances of a batch-major tensor."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    s, e = shard_range(t.shape[0], world_size, rank)
    return t[s:e]


class FlatGradBucket:
    """All trainable (LoRA) parameters' gradients as views into one flat fp32 buffer.

    ``param.grad`` is pre-set to a view of the bucket, so autograd's AccumulateGrad adds the K3 results in place and
    the all-reduce needs no gather/scatter copies.  Order = reverse registration order, so the bucket fills front to
    back as backward proceeds (decoder last layer first, encoder first layer last).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], reverse: bool = True):
        ps = [p for p in params if p.requires_grad]
        if not ps:
            raise ValueError("no trainable parameters")
        if reverse:
            ps = ps[::-1]
        self.params: List[torch.nn.Parameter] = ps
        dev = ps[0].device
        total = sum(p.numel() for p in ps)
        self.buffer = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in ps:
            if p.dtype != torch.float32:
                raise TypeError("FlatGradBucket expects fp32 LoRA parameters (PEFT keeps adapters in fp32)")
            v = self.buffer[off: off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            off += p.numel()

    def attach(self) -> int:
        """Re-point every ``param.grad`` at its bucket view; returns how many had been detached.  An optimizer's
        ``zero_grad(set_to_none=True)`` (the torch default) drops the views: autograd would then allocate fresh
        ``.grad`` tensors and the collective would reduce a bucket of zeros."""
        n = 0
        for p, v in zip(self.params, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
                n += 1
        return n

    def zero_(self) -> None:
        self.buffer.zero_()
        self.attach()

    def check_attached(self) -> None:
        """Raise if any gradient lives outside the bucket (``zero_grad(set_to_none=True)`` ran after ``zero_()``)."""
        bad = [i for i, (p, v) in enumerate(zip(self.params, self.views))
               if p.grad is None or p.grad.data_ptr() != v.data_ptr()]
        if bad:
            raise RuntimeError(f"{len(bad)} parameter gradients are not views of the flat bucket (first: #{bad[0]}); "
                               "call bucket.zero_() / bucket.attach() AFTER optimizer.zero_grad(), before backward")

    def all_reduce_mean(self, group=None, async_op: bool = False):
        """Sum the bucket over ranks (one collective), then scale by 1/world.  Returns a work handle if async."""
        self.check_attached()
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        if async_op:
            work = dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group, async_op=True)

            class _Handle:
                def wait(_self):
                    work.wait()
                    self.buffer.div_(world)
            return _Handle()
        dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)
        self.buffer.div_(world)
        return None

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """One fused norm over the flat buffer (replaces clip_grad_norm_ over all parameters,
        src/training/trainer.py:264-267).  Call after the all-reduce."""
        total = torch.linalg.vector_norm(self.buffer)
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.buffer.mul_(coef)
        return total
