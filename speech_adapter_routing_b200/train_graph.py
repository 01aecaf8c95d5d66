"""One CUDA graph for the forward + backward of a LoRA training step.

With 16 clips per GPU (the reference's batch size, scripts/train_lora.py:77) a training step is ~2500 kernel launches of
30-80 us each: measured with torch.profiler on a B200, 64.6 ms of GPU work took 109 ms of host time — the step is bound by
Python / launch overhead, not by any kernel.  ``GraphedTrainStep`` captures

    bucket.zero_()  ->  model(input_features, labels).loss  ->  loss.backward()  [->  overlapped NCCL all-reduce of the bucket]

once (after a few eager warm-up steps on a side stream) and replays it per batch: the libsar launches, ATen's LayerNorm /
GELU kernels, cuDNN's attention forward / backward and NCCL's all-reduce are all capturable, and K3 writes the gradients
into the static flat bucket.  The optimizer and the gradient clip stay outside the graph (a handful of launches).
Opt-in: the reference trainer's loop (src/training/trainer.py:233-304) keeps working without it.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import whisper_train
from .dist import FlatGradBucket
from .operand_refresh import OperandRefresh, cached_tensors


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, bucket: FlatGradBucket, input_features: torch.Tensor, labels: torch.Tensor,
                 warmup: int = 3, capture_allreduce: bool = False, check: bool = True, check_tol: float = 2e-2):
        """``input_features`` / ``labels``: example batch (shapes and dtypes are frozen into the graph).  ``check``: replay
        the captured step once on the example batch and compare its gradients with an eager step's (raises on mismatch:
        a library op that is not capture-safe must not silently train on garbage).

        Data parallel (torch.distributed initialised, world > 1): by default the graph holds the LOCAL step and
        ``__call__`` all-reduces the flat bucket right after the replay, as one eager NCCL call on the same stream (7 MB for
        whisper-small r16: 45 us alone).  ``capture_allreduce=True`` (experimental, needs ``bucket.enable_overlap``) instead
        makes the chunked all-reduces on the bucket's side stream nodes of the graph."""
        if not input_features.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors")
        self.model, self.bucket = model, bucket
        self.x = input_features.clone()
        self.labels = labels.clone()
        self.overlap = capture_allreduce and bucket._overlap is not None
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if bucket._overlap is not None and not self.overlap:
            bucket.set_overlap_enabled(False)      # this object owns the collective now: one call after each replay
        self.loss: Optional[torch.Tensor] = None
        dev = input_features.device
        self.refresh: Optional[OperandRefresh] = None
        if check and any(isinstance(m, torch.nn.Dropout) and m.p > 0 and m.training for m in model.modules()):
            check = False       # an eager step and a replay draw different dropout masks: nothing to compare bit-wise
        reference = self._eager_reference() if check else None
        failure: Optional[BaseException] = None
        try:
            self._warmup_and_capture(model, dev, warmup)
        except BaseException as e:          # noqa: BLE001 — re-raised below, on every rank together
            failure = e
        if self.world > 1:
            # ranks must agree before the next collective: one that failed to capture would otherwise leave the others
            # waiting in the check's all-reduce
            ok = torch.tensor([0 if failure is not None else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0 and failure is None:
                failure = RuntimeError("another rank failed to capture the training step")
        if failure is not None:
            raise failure
        # the graph addresses the operand caches (frozen-weight stacks, index vectors, ...) directly: keep them alive
        # when a later eager call rebuilds the caches after a parameter update
        self._keep = cached_tensors(model)
        if reference is not None:
            self._replay()
            torch.cuda.synchronize(dev)
            ref_loss, ref_grad = reference
            err = float((bucket.buffer - ref_grad).abs().max() / ref_grad.abs().max().clamp_min(1e-30))
            bad = not (err <= check_tol and abs(float(self.loss) - ref_loss) <= check_tol * max(abs(ref_loss), 1.0))
            if self.world > 1:                  # every rank raises, or none
                flag = torch.tensor([int(bad)], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                bad = bool(flag.item())
            if bad:
                raise RuntimeError(f"captured training step disagrees with the eager step (gradient rel. diff {err:.3e}, "
                                   f"loss {float(self.loss):.5f} vs {ref_loss:.5f}); fused-layer refusals: "
                                   f"{whisper_train.REFUSED}")

    def _warmup_and_capture(self, model, dev, warmup: int) -> None:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):       # first launches set kernel attributes / build operand packs
                    self._body()
                    # drop the autograd graph of the warm-up step: its AccumulateGrad nodes carry the stream they were made
                    # on, and the engine would join the capture stream with that (uncaptured) stream at the end of backward
                    self.loss = None
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            # every cached bf16 LoRA operand the warm-up built is re-derived from the live parameters by the first node of
            # the graph, so optimizer steps between replays are seen (operand_refresh.py)
            self.refresh = OperandRefresh(model)
            self.graph = torch.cuda.CUDAGraph()
            # NCCL's watchdog thread polls events while the capture runs: only this thread's calls may be policed then
            with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local" if self.overlap else "global"):
                self._body()
                self.loss = self.loss.detach()

    def _eager_reference(self):
        """Loss and flat (averaged) gradient of one eager step on the example batch, for the post-capture check."""
        was = self.bucket._overlap is not None and self.bucket._overlap.enabled
        if self.bucket._overlap is not None:
            self.bucket.set_overlap_enabled(False)
        try:
            self.bucket.zero_()
            out = self.model(input_features=self.x, labels=self.labels)
            loss = out.loss if hasattr(out, "loss") else out["loss"]
            loss.backward()
            if self.world > 1:
                self.bucket.all_reduce_mean()          # a step ends with the averaged gradient
            ref = (float(loss.detach()), self.bucket.buffer.clone())
        finally:
            if self.bucket._overlap is not None:
                self.bucket.set_overlap_enabled(was and self.overlap)
        return ref

    def _body(self) -> None:
        if self.refresh is not None:
            self.refresh.run()
        self.bucket.zero_()
        whisper_train.ASSUME_CAUSAL_MASK = True       # no padding mask is passed: a square mask is the causal one
        try:
            out = self.model(input_features=self.x, labels=self.labels)
            self.loss = out.loss if hasattr(out, "loss") else out["loss"]
            self.loss.backward()
        finally:
            whisper_train.ASSUME_CAUSAL_MASK = False
        if self.overlap:
            self.bucket.finish_overlap()

    def __call__(self, input_features: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """Runs one step on this batch; returns the (static) loss tensor.  Gradients are in ``bucket.buffer``."""
        self.x.copy_(input_features, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self._replay()
        return self.loss

    def _replay(self) -> None:
        self.graph.replay()
        if self.world > 1 and not self.overlap:
            self.bucket.all_reduce_mean()
