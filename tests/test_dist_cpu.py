"""Data-parallel host logic on CPU: utterance sharding and the flat LoRA-gradient bucket all-reduce, exercised with
world_size = 2 over gloo (the GPU path uses the same code with NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_adapter_routing_b200.dist import FlatGradBucket, shard_batch, shard_range


def test_shard_range_is_contiguous_balanced_and_complete():
    for n in (0, 1, 7, 64, 65, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_batch_without_process_group_is_identity():
    t = torch.arange(10).reshape(5, 2)
    assert torch.equal(shard_batch(t), t)
    assert torch.equal(shard_batch(t, 2, 1), t[3:])


def test_flat_bucket_views_alias_param_grads():
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    b = FlatGradBucket(ps)
    assert b.buffer.numel() == 17
    (ps[0].sum() * 2 + ps[1].sum() * 3).backward()     # AccumulateGrad adds in place into the views
    assert torch.equal(b.buffer, torch.cat([torch.full((5,), 3.0), torch.full((12,), 2.0)]))   # reverse order
    assert ps[0].grad.data_ptr() == b.views[1].data_ptr()
    n = b.clip_grad_norm_(1.0)
    assert torch.allclose(n, torch.tensor((5 * 9 + 12 * 4) ** 0.5))
    assert torch.allclose(torch.linalg.vector_norm(b.buffer), torch.tensor(1.0), atol=1e-5)
    b.zero_()
    assert float(b.buffer.abs().sum()) == 0.0 and ps[1].grad is b.views[0]


def test_flat_bucket_refuses_to_reduce_detached_grads():
    """optimizer.zero_grad(set_to_none=True) after bucket.zero_() drops the views; the bucket must notice
    instead of all-reducing zeros."""
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    b = FlatGradBucket(ps)
    opt = torch.optim.SGD(ps, lr=0.1)
    b.zero_()
    opt.zero_grad()                                  # set_to_none=True is the default
    (ps[0].sum() + ps[1].sum()).backward()           # grads land in fresh tensors, the bucket stays zero
    with pytest.raises(RuntimeError, match="not views of the flat bucket"):
        b.all_reduce_mean()
    assert b.attach() == 2
    b.zero_()
    (ps[0].sum() + ps[1].sum()).backward()
    b.all_reduce_mean()                              # world size 1: checks only
    assert torch.equal(b.buffer, torch.ones_like(b.buffer))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                      # identical replicas
        A = torch.nn.Parameter(torch.randn(4, 6))
        Bm = torch.nn.Parameter(torch.randn(6, 4))
        bucket = FlatGradBucket([A, Bm])
        # global batch of 5 "utterances", sharded 3 / 2; loss = mean over the GLOBAL batch
        g = torch.Generator().manual_seed(1)
        x = torch.randn(5, 6, generator=g)
        xs = shard_batch(x)
        assert xs.shape[0] == (3 if rank == 0 else 2)
        bucket.zero_()
        loss = ((xs @ A.t()) @ Bm.t()).pow(2).sum() / 5 * world   # local sum scaled so that the rank-mean is the global mean
        loss.backward()
        h = bucket.all_reduce_mean(async_op=True)
        h.wait()
        # single-process gradient on the concatenated batch
        A1 = A.detach().clone().requires_grad_(True)
        B1 = Bm.detach().clone().requires_grad_(True)
        (((x @ A1.t()) @ B1.t()).pow(2).sum() / 5).backward()
        ok = torch.allclose(A.grad, A1.grad, atol=1e-5) and torch.allclose(Bm.grad, B1.grad, atol=1e-5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduced_grads_equal_single_process_grads():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def _overlap_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(3, 4)) for _ in range(6)]          # a 6-"layer" chain, bucket in backward order
        bucket = FlatGradBucket(ps)
        bucket.enable_overlap(n_chunks=3)
        assert bucket._overlap.param_bounds == [0, 2, 4, 6]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(4, 4, generator=g)
        xs = shard_batch(x)
        ok = True
        for step in range(2):                                                    # second step: reset() re-arms the chunks
            bucket.zero_()
            h = xs
            for p_ in ps:
                h = torch.tanh(h @ p_.t() @ p_)
            (h.pow(2).sum() / 4 * world).backward()
            assert all(bucket._overlap.launched)                                 # every chunk went out DURING backward
            bucket.finish_overlap()
            refs = [p_.detach().clone().requires_grad_(True) for p_ in ps]
            h = x
            for r_ in refs:
                h = torch.tanh(h @ r_.t() @ r_)
            (h.pow(2).sum() / 4).backward()
            ok = ok and all(torch.allclose(p_.grad, r_.grad, atol=1e-5) for p_, r_ in zip(ps, refs))
        # gradient accumulation: a second backward without zero_() must not silently stack local gradients on chunks that
        # were already reduced; with the overlap switched off for the first micro-step it works
        h = xs
        for p_ in ps:
            h = torch.tanh(h @ p_.t() @ p_)
        try:
            h.sum().backward()
            ok = False
        except RuntimeError as e:
            ok = ok and "gradient accumulation" in str(e)
        bucket.zero_()
        bucket.set_overlap_enabled(False)
        for micro in range(2):
            if micro == 1:
                bucket.set_overlap_enabled(True)
            h = xs
            for p_ in ps:
                h = torch.tanh(h @ p_.t() @ p_)
            (h.pow(2).sum() / 4 * world).backward()
        bucket.finish_overlap()
        refs = [p_.detach().clone().requires_grad_(True) for p_ in ps]
        h = x
        for r_ in refs:
            h = torch.tanh(h @ r_.t() @ r_)
        (2 * h.pow(2).sum() / 4).backward()
        ok = ok and all(torch.allclose(p_.grad, r_.grad, atol=1e-5) for p_, r_ in zip(ps, refs))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_overlapped_chunked_allreduce_matches_single_process_grads():
    """FlatGradBucket.enable_overlap: chunks are all-reduced as soon as their gradients land (hooks fire during
    backward), results equal the single-process gradients of the global batch."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p_ in procs:
        p_.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
