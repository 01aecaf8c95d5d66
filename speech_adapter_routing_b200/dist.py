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
        self._offsets: List[int] = []
        off = 0
        for p in ps:
            if p.dtype != torch.float32:
                raise TypeError("FlatGradBucket expects fp32 LoRA parameters (PEFT keeps adapters in fp32)")
            v = self.buffer[off: off + p.numel()].view_as(p)
            p.grad = v
            p._sar_direct_grad = True      # lora_linear._QVLoRAFn: K3 may accumulate into this slice in place
            self.views.append(v)
            self._offsets.append(off)
            off += p.numel()
        self._offsets.append(off)
        self._index = {id(p): i for i, p in enumerate(ps)}
        self._overlap = None

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
        if self._overlap is not None:
            self._overlap.reset()

    # ---- all-reduce overlapped with backward -----------------------------------------------------------------------
    def enable_overlap(self, n_chunks: int = 4, group=None) -> None:
        """Split the bucket into ``n_chunks`` contiguous chunks (at parameter boundaries, ~equal bytes).  As soon as every
        gradient of a chunk has landed during backward — the bucket is in backward order, so chunks complete front to
        back — the chunk's all-reduce is launched on a side stream behind an event, while backward continues on the
        compute stream (src/training/trainer.py:251-277 semantics are kept: call ``finish_overlap()`` before the clip).
        Call ``zero_()`` at the start of every step."""
        self._overlap = _OverlappedAllReduce(self, n_chunks, group)

    def set_overlap_enabled(self, enabled: bool) -> None:
        """Switch the chunk all-reduces off / on (e.g. for a local reference step) without removing the hooks."""
        if self._overlap is not None:
            self._overlap.enabled = bool(enabled)

    def finish_overlap(self) -> None:
        """Wait for the chunk all-reduces (launching any chunk that never completed) and scale by 1/world."""
        if self._overlap is None:
            raise RuntimeError("enable_overlap() was not called")
        self.check_attached()
        self._overlap.finish()

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


class _OverlappedAllReduce:
    def __init__(self, bucket: FlatGradBucket, n_chunks: int, group):
        self.b, self.group = bucket, group
        total = bucket.buffer.numel()
        n_chunks = max(1, min(n_chunks, len(bucket.params)))
        bounds, target = [0], total / n_chunks
        for i in range(len(bucket.params)):
            end = bucket._offsets[i + 1]
            if end >= target * len(bounds) and len(bounds) < n_chunks:
                bounds.append(i + 1)
        if bounds[-1] != len(bucket.params):
            bounds.append(len(bucket.params))
        self.param_bounds = bounds                      # chunk c = params [bounds[c], bounds[c+1])
        self.chunk_of = [0] * len(bucket.params)
        for c in range(len(bounds) - 1):
            for i in range(bounds[c], bounds[c + 1]):
                self.chunk_of[i] = c
        self.side = torch.cuda.Stream(device=bucket.buffer.device) if bucket.buffer.is_cuda else None
        self.works: List = []
        self.launched: List[bool] = []
        self.pending: List[int] = []
        self.enabled = True
        self.reset()
        from . import lora_linear

        import weakref

        lora_linear.GRAD_READY_LISTENERS.append(weakref.WeakMethod(self._ready))   # gradients K3 wrote in place
        me = weakref.ref(self)

        def hook(q):
            o = me()
            if o is not None:
                o._ready((q,))

        for p in bucket.params:                                       # gradients autograd accumulated
            p.register_post_accumulate_grad_hook(hook)

    def reset(self) -> None:
        n = len(self.param_bounds) - 1
        self.pending = [self.param_bounds[c + 1] - self.param_bounds[c] for c in range(n)]
        self.launched = [False] * n
        self.seen = set()
        self.works = []

    def _active(self) -> bool:
        return self.enabled and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _ready(self, params) -> None:
        if not self.enabled:            # switched off (local reference step, or an early micro-step of an accumulation)
            return
        for p in params:
            i = self.b._index.get(id(p))
            if i is None:
                continue
            if i in self.seen:
                if self._active() and self.launched[self.chunk_of[i]]:
                    # a second backward since zero_() (gradient accumulation): this chunk has already been summed over
                    # the ranks, a later local gradient on top of it would never be reduced
                    raise RuntimeError("FlatGradBucket: a gradient arrived for a chunk that was already all-reduced in this "
                                       "step. With gradient accumulation call set_overlap_enabled(False) for all but the "
                                       "last micro-step (and zero_() once per optimizer step).")
                continue
            self.seen.add(i)
            c = self.chunk_of[i]
            self.pending[c] -= 1
            if self.pending[c] == 0:
                self._launch(c)

    def _launch(self, c: int) -> None:
        if self.launched[c]:
            return
        self.launched[c] = True
        if not self._active():
            return
        s, e = self.b._offsets[self.param_bounds[c]], self.b._offsets[self.param_bounds[c + 1]]
        view = self.b.buffer[s:e]
        if self.side is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(view.device))     # everything queued so far: this chunk's K3 launches
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        for c in range(len(self.launched)):
            if not self.launched[c]:
                self._launch(c)
        if not self._active():
            return
        if self.side is not None:
            torch.cuda.current_stream(self.b.buffer.device).wait_stream(self.side)
        for w in self.works:
            w.wait()
        self.b.buffer.div_(dist.get_world_size(self.group))
