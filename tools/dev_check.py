"""Developer bring-up script (GPU): runs each libsar kernel on a ladder of shapes against the CPU/torch oracle and
prints error statistics + CUDA-event timings.  Not part of the product; the judged tests live in tests/.

    python tools/dev_check.py k2 k1 rows k3 perf
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from oracle import fixtures, lora as olora, router as orouter  # noqa: E402
from speech_adapter_routing_b200 import ops  # noqa: E402

DEV = "cuda:0"


def log(*a):
    print(*a, flush=True)


def err_stats(y, ref):
    y = y.float().cpu()
    ref = ref.float().cpu()
    diff = (y - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    return diff.max().item(), diff.max().item() / denom, diff


def block_map(diff, tol, rb=32, cb=64):
    """Which (row-block, col-block) tiles contain errors — helps to spot swizzle / descriptor mistakes."""
    M, N = diff.shape
    bad = []
    for r0 in range(0, M, rb):
        for c0 in range(0, N, cb):
            if (diff[r0:r0 + rb, c0:c0 + cb] > tol).any():
                bad.append((r0, c0))
    return bad


def run_k1_case(name, B, T, d_in, d_out, r, n_adapters, mix="uniform", with_bias=True, base_only_every=0,
                block_n=0, grid=0, lora=True, seed=1234, kernel=0, swap=False):
    case = fixtures.make_lora_case(B, T, d_in, d_out, max(r, 8), max(n_adapters, 1), seed=seed, mix=mix,
                                   with_bias=with_bias, base_only_every=base_only_every)
    idx = case.utt_adapter if lora else torch.full((B,), -1, dtype=torch.int32)
    ref = olora.lora_linear_routed_k1_rounding(case.x, case.W, case.bias, case.A_stack, case.B_stack, case.scaling, idx)
    x = case.x.to(DEV); W = case.W.to(DEV)
    bias = None if case.bias is None else case.bias.to(DEV)
    if lora:
        A = case.A_stack.to(DEV); Bp = ops.pack_lora_b(case.B_stack.to(DEV)); ia = idx.to(DEV)
    else:
        A = Bp = ia = None
    y, _ = ops.qv_lora_fwd(x, W, bias, A, Bp, ia, case.scaling, block_n=block_n, grid=grid, kernel=kernel,
                           swap_halves=swap)
    torch.cuda.synchronize()
    mx, rel, diff = err_stats(y, ref)
    ok = rel < 2e-2
    log(f"[k1] {name:38s} B={B} T={T} d={d_in}->{d_out} r={r} n={n_adapters} bn={block_n or 'auto'}: "
        f"max_abs={mx:.4g} rel={rel:.3g} {'OK' if ok else 'FAIL'}")
    if not ok:
        d2 = diff.reshape(B * T, d_out)
        bad = block_map(d2, 0.02 * ref.float().abs().max().item())
        log(f"      bad 32x64 blocks: {len(bad)} of {((B*T+31)//32)*((d_out+63)//64)}; first: {bad[:12]}")
        yy = y.float().cpu().reshape(B * T, d_out); rr = ref.float().reshape(B * T, d_out)
        log("      y[0,:8]  =", [round(v, 4) for v in yy[0, :8].tolist()])
        log("      ref[0,:8]=", [round(v, 4) for v in rr[0, :8].tolist()])
        log("      y[1,:8]  =", [round(v, 4) for v in yy[1, :8].tolist()])
        log("      ref[1,:8]=", [round(v, 4) for v in rr[1, :8].tolist()])
        nanc = torch.isnan(yy).sum().item()
        log(f"      nan count={nanc}  zero count={(yy == 0).sum().item()}")
    return ok


def suite_k1():
    ok = True
    ok &= run_k1_case("base 1 tile 64x64x64", 1, 128, 64, 64, 16, 0, lora=False, with_bias=False)
    ok &= run_k1_case("base 1 tile + bias", 1, 128, 64, 64, 16, 0, lora=False)
    ok &= run_k1_case("base K=768 N=64", 1, 128, 768, 64, 16, 0, lora=False)
    ok &= run_k1_case("base 768x768 bn128", 1, 128, 768, 768, 16, 0, lora=False, block_n=128)
    ok &= run_k1_case("base 768x768 bn192", 1, 128, 768, 768, 16, 0, lora=False, block_n=192)
    ok &= run_k1_case("base tail T=92", 1, 92, 768, 768, 16, 0, lora=False)
    ok &= run_k1_case("base T=1500 B=2", 2, 1500, 768, 768, 16, 0, lora=False)
    ok &= run_k1_case("lora r16 single tile", 1, 128, 64, 64, 16, 1, mix="single")
    ok &= run_k1_case("lora r16 768 single adapter", 1, 128, 768, 768, 16, 1, mix="single")
    ok &= run_k1_case("lora r16 4 adapters", 8, 300, 768, 768, 16, 4)
    ok &= run_k1_case("lora r32 4 adapters 1024", 6, 200, 1024, 1024, 32, 4)
    ok &= run_k1_case("lora r64 8 adapters 1280", 9, 130, 1280, 1280, 64, 8)
    ok &= run_k1_case("lora r48", 3, 130, 768, 768, 48, 2)
    ok &= run_k1_case("lora + base-only utts", 8, 200, 768, 768, 16, 4, base_only_every=3)
    ok &= run_k1_case("lora 1 CTA serial (grid=1)", 4, 260, 768, 768, 16, 4, grid=1)
    ok &= run_k1_case("lora T=1 decode-shaped", 16, 1, 768, 768, 16, 4)
    ok &= run_k1_case("lora B=64 T=1500", 64, 1500, 768, 768, 16, 4)
    log("[k1] suite", "PASSED" if ok else "FAILED")
    return ok


def suite_k1v2():
    ok = True
    K = dict(kernel=2)
    a = run_k1_case("v2 base 1 unit (swap=0)", 1, 256, 64, 128, 16, 0, lora=False, with_bias=False, **K)
    b = run_k1_case("v2 base 1 unit (swap=1)", 1, 256, 64, 128, 16, 0, lora=False, with_bias=False, swap=True, **K)
    log(f"[k1v2] half assignment: swap=0 {'OK' if a else 'FAIL'}, swap=1 {'OK' if b else 'FAIL'}")
    ok &= a
    ok &= run_k1_case("v2 base K=768 N=768 bn192", 1, 256, 768, 768, 16, 0, lora=False, **K)
    ok &= run_k1_case("v2 base bn128", 1, 256, 768, 768, 16, 0, lora=False, block_n=128, **K)
    ok &= run_k1_case("v2 base tail T=300", 1, 300, 768, 768, 16, 0, lora=False, **K)
    ok &= run_k1_case("v2 base T=1500 B=2", 2, 1500, 768, 768, 16, 0, lora=False, **K)
    ok &= run_k1_case("v2 lora r16 1 unit", 1, 256, 768, 768, 16, 1, mix="single", **K)
    ok &= run_k1_case("v2 lora r16 4 adapters", 8, 300, 768, 768, 16, 4, **K)
    ok &= run_k1_case("v2 lora mid-unit ranges (5 pairs)", 3, 256, 768, 768, 16, 4, grid=10, **K)
    ok &= run_k1_case("v2 lora 1 pair serial", 4, 260, 768, 768, 16, 4, grid=2, **K)
    ok &= run_k1_case("v2 lora r32 1024 bn128", 6, 200, 1024, 1024, 32, 4, **K)
    ok &= run_k1_case("v2 lora r64 1280 bn128", 9, 130, 1280, 1280, 64, 8, **K)
    ok &= run_k1_case("v2 lora r48", 3, 130, 768, 768, 48, 2, **K)
    ok &= run_k1_case("v2 lora + base-only utts", 8, 200, 768, 768, 16, 4, base_only_every=3, **K)
    ok &= run_k1_case("v2 lora T=100 (half pair empty)", 5, 100, 768, 768, 16, 4, **K)
    ok &= run_k1_case("v2 lora B=64 T=1500", 64, 1500, 768, 768, 16, 4, **K)
    log("[k1v2] suite", "PASSED" if ok else "FAILED")
    return ok


def suite_k2():
    ok = True
    for (B, T, d, C, dt) in [(4, 50, 128, 4, torch.float32), (7, 333, 768, 4, torch.bfloat16),
                             (64, 1500, 768, 4, torch.bfloat16), (5, 1500, 1280, 8, torch.bfloat16),
                             (3, 100, 1024, 4, torch.float32)]:
        sd = fixtures.make_router_state_dict(d, C)
        h, langs = fixtures.make_encoder_states(B, T, d, C, dtype=dt)
        ref = orouter.classifier_forward(h, sd)
        ridx = ref["probs"].argmax(-1)
        rperm, rseg = orouter.segments(ridx, C)
        p = ops.RouterParams.from_state_dict(sd, DEV)
        out = ops.router_fwd(h.to(DEV), p)
        torch.cuda.synchronize()
        dl = (out.logits.cpu() - ref["logits"]).abs().max().item()
        dp = (out.probs.cpu() - ref["probs"]).abs().max().item()
        same = bool((out.idx.cpu().long() == ridx).all())
        perm_ok = bool((out.perm.cpu() == rperm).all()) and bool((out.seg_starts.cpu() == rseg).all())
        margin = orouter.top2_margin(ref["logits"]).min().item()
        good = same and perm_ok and dl < 1e-3
        ok &= good
        log(f"[k2] B={B} T={T} d={d} C={C} {str(dt)[6:]}: dlogit={dl:.3g} dprob={dp:.3g} idx_equal={same} "
            f"perm/seg_ok={perm_ok} min_margin={margin:.3g} {'OK' if good else 'FAIL'}")
    log("[k2] suite", "PASSED" if ok else "FAILED")
    return ok


def suite_rows():
    ok = True
    for (M, d, r, n) in [(16, 768, 16, 4), (64, 768, 16, 4), (130, 1280, 64, 8)]:
        case = fixtures.make_lora_case(M, 1, d, d, r, n, seed=77, base_only_every=5)
        ref = olora.lora_linear_routed_k1_rounding(case.x, case.W, case.bias, case.A_stack, case.B_stack,
                                                   case.scaling, case.utt_adapter).reshape(M, d)
        y = ops.qv_lora_fwd_rows(case.x.reshape(M, d).to(DEV), case.W.to(DEV), case.bias.to(DEV),
                                 case.A_stack.to(DEV), ops.pack_lora_b(case.B_stack.to(DEV)),
                                 case.utt_adapter.to(DEV), case.scaling)
        torch.cuda.synchronize()
        mx, rel, _ = err_stats(y, ref)
        good = rel < 2e-2
        ok &= good
        log(f"[rows] M={M} d={d} r={r} n={n}: max_abs={mx:.4g} rel={rel:.3g} {'OK' if good else 'FAIL'}")
    log("[rows] suite", "PASSED" if ok else "FAILED")
    return ok


def suite_k3():
    ok = True
    for (B, T, d, r, n) in [(2, 128, 128, 16, 1), (4, 300, 768, 16, 3), (3, 1500, 768, 16, 1), (3, 200, 1024, 32, 2),
                            (2, 130, 1280, 64, 2)]:
        case = fixtures.make_lora_case(B, T, d, d, r, n, seed=99, base_only_every=(3 if n > 1 else 0))
        g = torch.Generator().manual_seed(5)
        dy = (torch.randn(B, T, d, generator=g) * 0.1).to(torch.bfloat16)
        dx_ref, dA_ref, dB_ref = olora.lora_linear_backward(dy.float(), case.x.float(), case.W.float(),
                                                            case.A_stack.float(), case.B_stack.float(), case.scaling,
                                                            case.utt_adapter)
        x = case.x.to(DEV); W = case.W.to(DEV); A = case.A_stack.to(DEV); Bm = case.B_stack.to(DEV)
        ia = case.utt_adapter.to(DEV)
        y, u = ops.qv_lora_fwd(x, W, None, A, ops.pack_lora_b(Bm), ia, case.scaling, save_u=True)
        Wt = W.t().contiguous()
        At = ops.pack_lora_b(A.transpose(1, 2).contiguous())   # [n, d_in, 64]
        Bt = Bm.transpose(1, 2).contiguous()                    # [n, r, d_out]
        dA = torch.zeros(n, r, d, dtype=torch.float32, device=DEV)
        dB = torch.zeros(n, d, r, dtype=torch.float32, device=DEV)
        dx = ops.qv_lora_bwd(dy.to(DEV), x, u, Wt, At, Bt, ia, dA, dB, case.scaling)
        torch.cuda.synchronize()
        _, rx, _ = err_stats(dx, dx_ref)
        _, ra, _ = err_stats(dA, dA_ref)
        _, rb, _ = err_stats(dB, dB_ref)
        good = rx < 2e-2 and ra < 2e-2 and rb < 2e-2
        ok &= good
        log(f"[k3] B={B} T={T} d={d} r={r} n={n}: rel dx={rx:.3g} dA={ra:.3g} dB={rb:.3g} {'OK' if good else 'FAIL'}")
    log("[k3] suite", "PASSED" if ok else "FAILED")
    return ok


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def suite_perf():
    for (B, T, d, r, n) in [(64, 1500, 768, 16, 4), (64, 1500, 1024, 32, 4), (64, 1500, 1280, 64, 8)]:
        case = fixtures.make_lora_case(4, 8, d, d, r, n)
        g = torch.Generator().manual_seed(1)
        nbuf = 4
        xs = [torch.randn(B, T, d, device=DEV, dtype=torch.bfloat16) for _ in range(nbuf)]
        ys = [torch.empty(B, T, d, device=DEV, dtype=torch.bfloat16) for _ in range(nbuf)]
        W = case.W.to(DEV); bias = case.bias.to(DEV); A = case.A_stack.to(DEV); Bp = ops.pack_lora_b(case.B_stack.to(DEV))
        ia = torch.randint(0, n, (B,), generator=g).to(torch.int32).to(DEV)
        flops = 2.0 * B * T * d * d + 2.0 * B * T * r * 2 * d
        for bn in ([128, 192] if d % 192 == 0 else [128]):
          for kern in (1, 2):
            it = [0]

            def f():
                i = it[0] % nbuf
                it[0] += 1
                ops.qv_lora_fwd(xs[i], W, bias, A, Bp, ia, 2.0, block_n=bn, out=ys[i], kernel=kern)
            ms = timeit(f)
            log(f"[perf] k1v{kern} lora  d={d} r={r} bn={bn}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s")

            def fb():
                i = it[0] % nbuf
                it[0] += 1
                ops.qv_lora_fwd(xs[i], W, bias, None, None, None, 2.0, block_n=bn, out=ys[i], kernel=kern)
            ms = timeit(fb)
            log(f"[perf] k1v{kern} base  d={d} bn={bn}: {ms*1e3:.1f} us  {2.0*B*T*d*d/ms/1e9:.1f} TFLOP/s")

        def fc():
            i = it[0] % nbuf
            it[0] += 1
            torch.nn.functional.linear(xs[i], W, bias)
        ms = timeit(fc)
        log(f"[perf] cuBLAS base d={d}: {ms*1e3:.1f} us  {2.0*B*T*d*d/ms/1e9:.1f} TFLOP/s")
        # router
        sd = fixtures.make_router_state_dict(d, n)
        p = ops.RouterParams.from_state_dict(sd, DEV)
        ms = timeit(lambda: ops.router_fwd(xs[it[0] % nbuf], p))
        log(f"[perf] k2 router d={d}: {ms*1e3:.1f} us  {B*T*d*2/ms/1e6:.1f} GB/s (includes alloc + 2 launches)")


if __name__ == "__main__":
    suites = sys.argv[1:] or ["k2", "k1", "rows", "k3", "perf"]
    t0 = time.time()
    log("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    allok = True
    for s in suites:
        allok &= bool(globals()["suite_" + s]() in (True, None))
        log(f"--- {s} done at {time.time()-t0:.1f}s")
    sys.exit(0 if allok else 1)
