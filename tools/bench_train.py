"""BASELINE config 5: whisper-small LoRA r16 training step (forward + LoRA-only backward, bf16 base model, fp32 adapters),
batch 16 per GPU, adapter gradients all-reduced over NCCL in one flat bucket, fused clip, AdamW on the LoRA tensors.

    python tools/bench_train.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py

Mirrors the reference trainer's step (src/training/trainer.py:251-277, 314-339) with use_gradient_checkpointing as in
WhisperLoRA's default.  Prints one JSON line on rank 0 (a secondary measurement: bench.py carries the headline).
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("SAR_RANDOM_INIT", "1")
import speech_adapter_routing_b200 as sar  # noqa: E402
from speech_adapter_routing_b200 import ops  # noqa: E402
from speech_adapter_routing_b200.dist import FlatGradBucket  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-checkpointing", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = sar.WhisperLoRA("whisper-small", lora_r=16, lora_alpha=32, lora_dropout=0.0, device=str(dev),
                        use_gradient_checkpointing=not args.no_checkpointing)
    w.train()
    cfg = w.model.config
    params = [p for p in w.model.parameters() if p.requires_grad]
    bucket = FlatGradBucket(params)
    opt = torch.optim.AdamW(params, lr=1e-4)
    g = torch.Generator().manual_seed(100 + rank)
    B, T_dec = args.batch, 128
    xs = [torch.randn(B, cfg.num_mel_bins, 3000, generator=g).to(dev).to(torch.bfloat16) for _ in range(2)]
    labels = torch.randint(5, cfg.vocab_size, (B, T_dec), generator=g).to(dev)

    def step(i):
        bucket.zero_()
        loss = w(input_features=xs[i % 2], labels=labels).loss
        loss.backward()
        bucket.all_reduce_mean()
        bucket.clip_grad_norm_(1.0)
        opt.step()
        return loss

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ops.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    if rank == 0:
        print(json.dumps({"metric": "whisper_small_lora_r16_train_step_clips_per_sec", "value": world * B / (ms * 1e-3),
                          "unit": "clips/s", "n_gpus": world, "ms_per_step": ms, "batch_per_gpu": B, "t_dec": T_dec,
                          "gradient_checkpointing": not args.no_checkpointing, "loss": float(loss),
                          "trainable_params": sum(p.numel() for p in params), "allreduce_bytes": bucket.buffer.numel() * 4,
                          "libsar_launches": dict(ops.LAUNCHES)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
