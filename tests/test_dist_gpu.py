"""Data-parallel LoRA training on 2 / 4 / 8 B200s (one process per GPU, NCCL over NVLink): the all-reduced adapter
gradients of N ranks, each on its shard of the batch, must equal the single-GPU gradients on the concatenated batch
(the one collective of the path; reference step: src/training/trainer.py:251-277).  Needs >= 2 GPUs
(`gpurun --gpus N`); skipped on a single-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SAR_RANDOM_INIT="1", HF_HUB_OFFLINE="1")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import speech_adapter_routing_b200 as sar
        from speech_adapter_routing_b200.dist import FlatGradBucket, shard_batch

        w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device=str(dev),
                            use_gradient_checkpointing=True)           # same seed on every rank: identical replicas
        w.train()
        with torch.no_grad():
            g = torch.Generator().manual_seed(5)
            for m in sar.lora_modules(w.model).values():               # PEFT's zero lora_B would zero every dA
                m.lora_B["default"].weight.copy_((torch.randn(m.out_features, 16, generator=g) * 0.02).to(dev))
        cfg = w.model.config
        params = [p for p in w.model.parameters() if p.requires_grad]
        bucket = FlatGradBucket(params)
        bucket.enable_overlap(n_chunks=4)
        per_rank = 2
        g = torch.Generator().manual_seed(11)
        x = torch.randn(world * per_rank, cfg.num_mel_bins, 3000, generator=g).to(torch.bfloat16)
        labels = torch.randint(5, cfg.vocab_size, (world * per_rank, 12), generator=g)
        bucket.zero_()
        w(input_features=shard_batch(x).to(dev), labels=shard_batch(labels).to(dev)).loss.backward()
        launched_during_backward = all(bucket._overlap.launched)
        bucket.finish_overlap()
        reduced = bucket.buffer.clone()
        # the same replica on the WHOLE batch, no collective
        bucket.set_overlap_enabled(False)
        dist.barrier()
        bucket.zero_()
        w(input_features=x.to(dev), labels=labels.to(dev)).loss.backward()
        single = bucket.buffer.clone()
        err = ((reduced - single).abs().max() / single.abs().max().clamp_min(1e-12)).item()
        same_across_ranks = [torch.zeros_like(reduced) for _ in range(world)]
        dist.all_gather(same_across_ranks, reduced)
        identical = all(torch.equal(t, reduced) for t in same_across_ranks)
        # the same local step as ONE captured CUDA graph per rank, the flat bucket all-reduced by one NCCL call after each
        # replay: its averaged gradients equal the eager step's
        xs, ls = shard_batch(x).to(dev), shard_batch(labels).to(dev)
        step = sar.GraphedTrainStep(w, bucket, xs, ls, warmup=2)        # check=True compares with an eager reduced step
        step(xs, ls)
        torch.cuda.synchronize()
        gerr = ((bucket.buffer - reduced).abs().max() / reduced.abs().max().clamp_min(1e-12)).item()
        q.put((rank, err, bool(launched_during_backward), bool(identical), float(single.abs().max()), gerr))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_allreduced_lora_grads_equal_single_gpu_grads(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results = [q.get(timeout=300) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():                 # a rank stuck in a collective must not hang the suite
                p.kill()
    for rank, err, launched, identical, gmax, gerr in sorted(results):
        assert gmax > 0
        assert gerr <= 2e-2, (rank, gerr)                              # graph replay vs the eager reduced step
        assert launched, "a chunk's all-reduce was not launched during backward"
        assert identical, "ranks disagree on the reduced bucket"
        # bf16 activations, fp32 reductions in a different order (per-rank partial sums, then NCCL's tree)
        assert err <= 2e-2, (rank, err)
