"""GPU parity tests of the libsar kernels through the C ABI (ctypes) against the CPU oracle on identical seeded inputs.

Tolerances (floating point, bf16 storage / fp32 accumulate):
  K1 / rows / K3-dx vs the same-rounding oracle:  max|err| <= 2^-7 * max|ref|   (one bf16 rounding of the result)
  K1 vs the fp32 PEFT-formula oracle:              max|err| <= 2^-6 * max|ref|
  K3 dA / dB (fp32 outputs, bf16 operands) vs fp32 autograd: max|err| <= 2^-6 * max|ref|
  K2 logits vs oracle: 2e-4 absolute; adapter indices, perm and seg_starts: BIT-EXACT.
"""
from pathlib import Path

import pytest
import torch

from oracle import fixtures, lora as olora, router as orouter
from speech_adapter_routing_b200 import ops

pytestmark = pytest.mark.gpu

TIGHT = 2.0 ** -7
LOOSE = 2.0 ** -6


def rel_err(y, ref):
    y, ref = y.float().cpu(), ref.float().cpu()
    return ((y - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def run_k1(case, dev, idx=None, lora=True, **kw):
    idx = case.utt_adapter if idx is None else idx
    bias = None if case.bias is None else case.bias.to(dev)
    if not lora:
        return ops.qv_lora_fwd(case.x.to(dev), case.W.to(dev), bias, None, None, None, case.scaling, **kw)[0]
    return ops.qv_lora_fwd(case.x.to(dev), case.W.to(dev), bias, case.A_stack.to(dev),
                           ops.pack_lora_b(case.B_stack.to(dev)), idx.to(dev), case.scaling, **kw)[0]


# ------------------------------------------------------------------------------------------------ K1
K1_CASES = [
    # B, T, d_in, d_out, r, n, mix, bias, base_only_every
    (1, 1500, 768, 768, 16, 1, "single", True, 0),      # BASELINE config 1 shape (single adapter, one clip)
    (1, 92, 768, 768, 16, 4, "uniform", True, 0),       # tail tile only
    (3, 128, 768, 768, 16, 4, "uniform", True, 0),      # exactly one tile per utterance
    (8, 300, 768, 768, 16, 4, "uniform", True, 3),      # mixed adapters + base-only utterances
    (8, 300, 768, 768, 16, 4, "sorted", False, 0),      # sorted order, no bias
    (8, 300, 768, 768, 16, 4, "skewed", True, 0),
    (6, 200, 1024, 1024, 32, 4, "uniform", True, 0),    # medium geometry r32
    (5, 130, 1280, 1280, 64, 8, "uniform", True, 0),    # large-v3 geometry r64, 8 adapters
    (3, 130, 768, 768, 48, 2, "uniform", True, 0),
    (16, 1, 768, 768, 16, 4, "uniform", True, 0),       # decode-shaped: one row per utterance
    (2, 448, 768, 1024, 16, 2, "uniform", True, 0),     # rectangular, max decoder length
]


@pytest.mark.parametrize("B,T,d_in,d_out,r,n,mix,bias,bo", K1_CASES)
def test_k1_matches_oracle(cuda_dev, B, T, d_in, d_out, r, n, mix, bias, bo):
    c = fixtures.make_lora_case(B, T, d_in, d_out, r, n, mix=mix, with_bias=bias, base_only_every=bo)
    y = run_k1(c, cuda_dev)
    same = olora.lora_linear_routed_k1_rounding(c.x, c.W, c.bias, c.A_stack, c.B_stack, c.scaling, c.utt_adapter)
    assert rel_err(y, same) <= TIGHT
    fp32 = olora.lora_linear_routed(c.x.float(), c.W.float(), None if c.bias is None else c.bias.float(),
                                    c.A_stack.float(), c.B_stack.float(), c.scaling, c.utt_adapter)
    assert rel_err(y, fp32) <= LOOSE
    assert y.shape == (B, T, d_out) and y.dtype == torch.bfloat16


@pytest.mark.parametrize("block_n", [64, 128, 192])
def test_k1_tile_configurations_agree(cuda_dev, block_n):
    c = fixtures.make_lora_case(4, 260, 768, 768, 16, 4)
    ref = olora.lora_linear_routed_k1_rounding(c.x, c.W, c.bias, c.A_stack, c.B_stack, c.scaling, c.utt_adapter)
    assert rel_err(run_k1(c, cuda_dev, block_n=block_n), ref) <= TIGHT
    assert rel_err(run_k1(c, cuda_dev, block_n=block_n, grid=1), ref) <= TIGHT      # one persistent CTA walks all units


def test_k1_full_size_properties(cuda_dev):
    """BASELINE config 2 size (whisper-small, 4 adapters r16, B=64, T=1500): size-independent properties."""
    c = fixtures.make_lora_case(64, 1500, 768, 768, 16, 4)
    y = run_k1(c, cuda_dev)
    # (1) utterances are independent: an utterance computed alone gives bit-identical rows
    for b in (0, 17, 63):
        one = fixtures.LoraCase(c.x[b:b + 1], c.W, c.bias, c.A_stack, c.B_stack, c.utt_adapter[b:b + 1], c.scaling)
        assert torch.equal(run_k1(one, cuda_dev)[0], y[b])
    # (2) adapter id -1 == the base projection; (3) a zero lora_B leaves the base projection unchanged (PEFT init)
    base = run_k1(c, cuda_dev, lora=False)
    assert torch.equal(run_k1(c, cuda_dev, idx=torch.full((64,), -1, dtype=torch.int32)), base)
    z = fixtures.LoraCase(c.x, c.W, c.bias, c.A_stack, torch.zeros_like(c.B_stack), c.utt_adapter, c.scaling)
    assert torch.equal(run_k1(z, cuda_dev), base)
    # (4) checked rows against the oracle
    rows = olora.lora_linear_routed_k1_rounding(c.x[5:7], c.W, c.bias, c.A_stack, c.B_stack, c.scaling,
                                                c.utt_adapter[5:7])
    assert rel_err(y[5:7], rows) <= TIGHT
    # (5) out-of-range adapter ids are treated as "no adapter" rather than read out of bounds
    bad = c.utt_adapter.clone()
    bad[::2] = 99
    yb = run_k1(c, cuda_dev, idx=bad)
    assert torch.equal(yb[0], base[0]) and torch.equal(yb[1], y[1])


def test_k1_rejects_bad_arguments(cuda_dev):
    from speech_adapter_routing_b200._lib import SarError
    x = torch.zeros(1, 4, 100, dtype=torch.bfloat16, device=cuda_dev)
    W = torch.zeros(64, 100, dtype=torch.bfloat16, device=cuda_dev)
    with pytest.raises(SarError, match="multiples of 64"):
        ops.qv_lora_fwd(x, W, None, None, None, None, 1.0)
    with pytest.raises(TypeError):
        ops.qv_lora_fwd(x.float(), W, None, None, None, None, 1.0)


# ------------------------------------------------------------------------------------------------ rows variant
@pytest.mark.parametrize("M,d,r,n", [(1, 768, 16, 4), (16, 768, 16, 4), (64, 768, 16, 4), (130, 1280, 64, 8),
                                     (200, 1024, 32, 4)])
def test_rows_variant_matches_oracle(cuda_dev, M, d, r, n):
    c = fixtures.make_lora_case(M, 1, d, d, r, n, seed=77, base_only_every=5)
    ref = olora.lora_linear_routed_k1_rounding(c.x, c.W, c.bias, c.A_stack, c.B_stack, c.scaling,
                                               c.utt_adapter).reshape(M, d)
    y = ops.qv_lora_fwd_rows(c.x.reshape(M, d).to(cuda_dev), c.W.to(cuda_dev), c.bias.to(cuda_dev),
                             c.A_stack.to(cuda_dev), ops.pack_lora_b(c.B_stack.to(cuda_dev)),
                             c.utt_adapter.to(cuda_dev), c.scaling)
    assert rel_err(y, ref) <= LOOSE      # base is rounded to bf16 before the low-rank term is added (two roundings)


# ------------------------------------------------------------------------------------------------ K2
def run_k2(h, sd, dev):
    return ops.router_fwd(h.to(dev), ops.RouterParams.from_state_dict(sd, dev))


def test_k2_on_reference_golden_vectors(cuda_dev):
    """Inputs and expected outputs come from the REFERENCE's LanguageClassifier (tests/golden/make_golden.py)."""
    g = torch.load(Path(__file__).parent / "golden" / "router_golden.pt")
    for c in g["cases"]:
        out = run_k2(c["h"], c["state_dict"], cuda_dev)
        assert torch.equal(out.idx.cpu().long(), c["labels"])                  # bit-exact routing
        assert (out.logits.cpu() - c["logits"]).abs().max() <= 2e-4
        assert (out.probs.cpu() - c["probs"]).abs().max() <= 1e-4


@pytest.mark.parametrize("B,T,d,C,dt", [(1, 1500, 768, 4, torch.bfloat16), (7, 333, 768, 4, torch.bfloat16),
                                        (5, 1500, 1280, 8, torch.bfloat16), (3, 100, 1024, 4, torch.float32),
                                        (9, 17, 384, 3, torch.float32), (4, 1500, 1024, 4, torch.bfloat16)])
def test_k2_matches_oracle(cuda_dev, B, T, d, C, dt):
    sd = fixtures.make_router_state_dict(d, C)
    h, _ = fixtures.make_encoder_states(B, T, d, C, dtype=dt)
    ref = orouter.classifier_forward(h, sd)
    ridx = ref["probs"].argmax(-1)
    assert orouter.top2_margin(ref["logits"]).min() > 1e-2, "fixture margin too small for an index parity claim"
    out = run_k2(h, sd, cuda_dev)
    assert torch.equal(out.idx.cpu().long(), ridx)
    perm, seg = orouter.segments(ridx, C)
    assert torch.equal(out.perm.cpu(), perm) and torch.equal(out.seg_starts.cpu(), seg)
    assert (out.logits.cpu() - ref["logits"]).abs().max() <= 2e-4
    assert (out.probs.cpu() - ref["probs"]).abs().max() <= 1e-4
    assert out.idx.dtype == torch.int32 and out.logits.dtype == torch.float32


def test_k2_full_size_properties(cuda_dev):
    """B=64 x 1500 frames (config 2) and B=512 (config 4 batch): permutation equivariance, segment invariants."""
    C, d = 4, 768
    sd = fixtures.make_router_state_dict(d, C)
    for B in (64, 512):
        langs = fixtures.language_mix(B, C, "skewed")
        h, _ = fixtures.make_encoder_states(B, 1500 if B == 64 else 64, d, C, langs=langs)
        out = run_k2(h, sd, cuda_dev)
        idx = out.idx.cpu().long()
        g = torch.Generator().manual_seed(3)
        p = torch.randperm(B, generator=g)
        out_p = run_k2(h[p], sd, cuda_dev)
        assert torch.equal(out_p.idx.cpu().long(), idx[p])                      # routing is per utterance
        assert torch.allclose(out_p.logits.cpu(), out.logits.cpu()[p], atol=1e-5)
        perm, seg = out.perm.cpu().long(), out.seg_starts.cpu().long()
        assert sorted(perm.tolist()) == list(range(B))                          # a permutation
        assert torch.equal(idx[perm], idx[perm].sort().values)                  # sorted by adapter
        for k in range(C):                                                      # stable, segment k holds adapter k
            s = perm[seg[k]:seg[k + 1]]
            assert bool((idx[s] == k).all()) and torch.equal(s, s.sort().values)
        assert seg[0] == 0 and seg[-1] == B
        assert torch.allclose(out.probs.cpu().sum(-1), torch.ones(B), atol=1e-5)


def test_k2_ties_pick_first_index(cuda_dev):
    sd = fixtures.make_router_state_dict(128, 4)
    sd["classifier.8.weight"] = torch.zeros_like(sd["classifier.8.weight"])
    sd["classifier.8.bias"] = torch.tensor([0.5, 1.0, 1.0, 0.25])
    h, _ = fixtures.make_encoder_states(3, 20, 128, 4, dtype=torch.float32)
    out = run_k2(h, sd, cuda_dev)
    assert out.idx.cpu().tolist() == [1, 1, 1]                                   # torch.argmax tie rule


@pytest.mark.parametrize("B,T,d,C", [(5, 1500, 768, 4), (3, 333, 1024, 4), (4, 257, 1280, 8), (6, 100, 256, 4)])
def test_k2_with_the_encoder_final_layernorm_folded_in(cuda_dev, B, T, d, C):
    """sar_router_fwd_fused_ln on the encoder's PRE-LayerNorm residual stream == LayerNorm kernel + K2 on its bf16 output
    (SURVEY §8(f)-4): indices / perm / segments bit-exact, logits to fp32 summation noise; and both agree with the
    reference head's restatement applied to bf16(LayerNorm(h))."""
    g = torch.Generator().manual_seed(d + T)
    sd = fixtures.make_router_state_dict(d, C)
    feats, _ = fixtures.make_encoder_states(B, T, d, C, dtype=torch.float32)
    h_pre = (feats * 2.5 + 0.7).to(torch.bfloat16)                    # un-normalised residual stream
    gam = (1.0 + 0.1 * torch.randn(d, generator=g)).to(torch.bfloat16)
    bet = (0.1 * torch.randn(d, generator=g)).to(torch.bfloat16)
    dev = cuda_dev
    params = ops.RouterParams.from_state_dict(sd, dev)
    fused = ops.router_fwd(h_pre.to(dev), params, pre_ln=(gam.to(dev), bet.to(dev), 1e-5))
    x = ops.layernorm_fwd(h_pre.to(dev), gam.to(dev), bet.to(dev), 1e-5)
    two = ops.router_fwd(x, params)
    assert torch.equal(fused.idx, two.idx) and torch.equal(fused.perm, two.perm)
    assert torch.equal(fused.seg_starts, two.seg_starts)
    assert (fused.logits - two.logits).abs().max().item() <= 2e-3     # x may differ from the fused path's by one bf16 ulp
    ref = orouter.classifier_forward(x.cpu(), sd)
    assert orouter.top2_margin(ref["logits"]).min() > 2e-2
    assert torch.equal(fused.idx.cpu().long(), ref["probs"].argmax(-1))
    assert (fused.logits.cpu() - ref["logits"]).abs().max() <= 2e-3


# ------------------------------------------------------------------------------------------------ K3
@pytest.mark.parametrize("B,T,d,r,n,bo", [(2, 128, 128, 16, 1, 0), (4, 300, 768, 16, 3, 3), (3, 1500, 768, 16, 1, 0),
                                          (3, 200, 1024, 32, 2, 0), (2, 130, 1280, 64, 2, 0), (2, 77, 768, 48, 2, 0)])
def test_k3_matches_oracle_autograd(cuda_dev, B, T, d, r, n, bo):
    dev = cuda_dev
    c = fixtures.make_lora_case(B, T, d, d, r, n, seed=99, base_only_every=bo)
    dy = (torch.randn(B, T, d, generator=torch.Generator().manual_seed(5)) * 0.1).to(torch.bfloat16)
    dx_ref, dA_ref, dB_ref = olora.lora_linear_backward(dy.float(), c.x.float(), c.W.float(), c.A_stack.float(),
                                                        c.B_stack.float(), c.scaling, c.utt_adapter)
    x, W, A, Bm, ia = c.x.to(dev), c.W.to(dev), c.A_stack.to(dev), c.B_stack.to(dev), c.utt_adapter.to(dev)
    _, u = ops.qv_lora_fwd(x, W, None, A, ops.pack_lora_b(Bm), ia, c.scaling, save_u=True)
    Wt = W.t().contiguous()
    At = ops.pack_lora_b(A.transpose(1, 2).contiguous())
    Bt = Bm.transpose(1, 2).contiguous()
    dA = torch.zeros(n, r, d, dtype=torch.float32, device=dev)
    dB = torch.zeros(n, d, r, dtype=torch.float32, device=dev)
    dx = ops.qv_lora_bwd(dy.to(dev), x, u, Wt, At, Bt, ia, dA, dB, c.scaling)
    assert rel_err(dx, dx_ref) <= LOOSE
    assert rel_err(dA, dA_ref) <= LOOSE and rel_err(dB, dB_ref) <= LOOSE
    # gradients ACCUMULATE into the bucket (grad accumulation, reference trainer.py:259) and dx may be skipped
    none = ops.qv_lora_bwd(dy.to(dev), x, u, Wt, At, Bt, ia, dA, dB, c.scaling, need_dx=False)
    assert none is None
    assert rel_err(dA, 2 * dA_ref) <= LOOSE and rel_err(dB, 2 * dB_ref) <= LOOSE
    # adapters that no utterance selected keep a zero gradient
    for k in range(n):
        if not bool((c.utt_adapter == k).any()):
            assert float(dA[k].abs().max()) == 0.0 and float(dB[k].abs().max()) == 0.0


def test_k3_is_deterministic(cuda_dev):
    dev = cuda_dev
    c = fixtures.make_lora_case(6, 700, 768, 768, 16, 2, seed=3)
    dy = torch.randn(6, 700, 768, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    x, W, A, Bm, ia = c.x.to(dev), c.W.to(dev), c.A_stack.to(dev), c.B_stack.to(dev), c.utt_adapter.to(dev)
    _, u = ops.qv_lora_fwd(x, W, None, A, ops.pack_lora_b(Bm), ia, c.scaling, save_u=True)
    outs = []
    for _ in range(2):
        dA = torch.zeros(2, 16, 768, dtype=torch.float32, device=dev)
        dB = torch.zeros(2, 768, 16, dtype=torch.float32, device=dev)
        dx = ops.qv_lora_bwd(dy, x, u, W.t().contiguous(), ops.pack_lora_b(A.transpose(1, 2).contiguous()),
                             Bm.transpose(1, 2).contiguous(), ia, dA, dB, c.scaling)
        outs.append((dx.clone(), dA.clone(), dB.clone()))
    assert all(torch.equal(a, b) for a, b in zip(*outs))      # fixed-order reductions, no float atomics


# ------------------------------------------------------------------------------------------------ log-mel front-end
@pytest.mark.parametrize("n_mels,dtype,tol", [(80, torch.float32, 2e-3), (128, torch.float32, 2e-3), (80, torch.bfloat16, 1e-2)])
def test_logmel_matches_oracle(cuda_dev, n_mels, dtype, tol):
    """sar_logmel_fwd (waveform -> input_features on the GPU) vs the float64 oracle that is pinned against
    WhisperFeatureExtractor: a tone in noise (5 s, padded), over-long noise (cut at 30 s), a click in silence, and an
    all-zero clip.  fp32 output within 2e-3 absolute (values span about [-1.5, 1.5]); bf16 output adds one rounding."""
    import numpy as np

    from oracle import logmel as ologmel
    from speech_adapter_routing_b200 import logmel

    rng = np.random.default_rng(3)
    t = np.arange(16000 * 5) / 16000.0
    clips = [
        (0.3 * np.sin(2 * np.pi * 440.0 * t) + 0.05 * rng.standard_normal(t.shape)).astype(np.float32),
        (0.1 * rng.standard_normal(16000 * 31)).astype(np.float32),
        np.concatenate([np.zeros(8000, np.float32), np.ones(1, np.float32), np.zeros(4000, np.float32)]),
        np.zeros(16000, np.float32),
    ]
    out = logmel.log_mel_spectrogram([torch.from_numpy(c).to(cuda_dev) for c in clips], n_mels=n_mels, dtype=dtype)
    assert out.shape == (len(clips), n_mels, 3000) and out.dtype == dtype
    for i, c in enumerate(clips):
        ref = torch.from_numpy(ologmel.log_mel(c, n_mels))
        err = (out[i].double().cpu() - ref).abs().max().item()
        assert err <= tol, f"clip {i}: max abs err {err}"
    # a clip processed alone gives identical values (the per-clip maximum never leaks across clips)
    one = logmel.log_mel_spectrogram([torch.from_numpy(clips[0]).to(cuda_dev)], n_mels=n_mels, dtype=dtype)
    assert torch.equal(one[0], out[0])


# ------------------------------------------------------------------------------------------------ LayerNorm + LoRA-down (fused)
LNU_CASES = [
    # B, T, d, r, n_sets, n_adapters, base_only_every
    (3, 1500, 768, 16, 2, 4, 0),      # whisper-small encoder q|k|v (BASELINE config 2 shape family)
    (4, 128, 768, 16, 2, 4, 3),       # decoder self-attention rows, one base-only utterance
    (5, 13, 768, 16, 1, 4, 0),        # cross-attention q: one set; ragged last 8-row group
    (3, 100, 256, 16, 2, 3, 0),       # micro test geometry
    (2, 37, 384, 16, 2, 2, 2),        # whisper-tiny
    (2, 300, 512, 32, 1, 2, 0),       # whisper-base, r32, one set
    (64, 128, 768, 16, 2, 4, 0),      # many short utterances: CTAs cross utterance boundaries, A_k re-staged
]


def test_layernorm_lora_u_support_matrix(libsar):
    """d <= 768 with 16 or 32 rank columns over all sets; everything else keeps LayerNorm + the tcgen05 U pass."""
    assert ops.layernorm_lora_u_supported(768, 16, 2) and ops.layernorm_lora_u_supported(384, 32, 1)
    assert ops.layernorm_lora_u_supported(768, 16, 1) and ops.layernorm_lora_u_supported(256, 16, 2)
    assert not ops.layernorm_lora_u_supported(1024, 32, 2) and not ops.layernorm_lora_u_supported(1280, 64, 2)
    assert not ops.layernorm_lora_u_supported(768, 32, 2) and not ops.layernorm_lora_u_supported(768, 24, 1)


@pytest.mark.parametrize("B,T,d,r,n_sets,n,bo", LNU_CASES)
def test_layernorm_lora_u_fused_matches_layernorm_and_the_u_pass(cuda_dev, B, T, d, r, n_sets, n, bo):
    """sar_layernorm_lora_u_fwd: x must be LayerNorm(h) (same arithmetic as sar_layernorm_fwd: at most a last-bit bf16
    difference from the different fp32 summation order) and U = bf16(scale · x · A_kᵀ) for every LoRA set — the operand the
    split path's dense kernel consumes (reference ops: nn.LayerNorm + PEFT lora_A at q_proj / v_proj)."""
    assert ops.layernorm_lora_u_supported(d, r, n_sets)
    g = torch.Generator().manual_seed(d + T)
    h = (torch.randn(B, T, d, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    gamma = (1.0 + 0.1 * torch.randn(d, generator=g)).to(torch.bfloat16)
    beta = (0.1 * torch.randn(d, generator=g)).to(torch.bfloat16)
    A = ((torch.rand(n_sets * n, r, d, generator=g) * 2 - 1) / d ** 0.5).to(torch.bfloat16)
    idx = torch.randint(0, n, (B,), generator=g).to(torch.int32)
    if bo:
        idx[::bo] = -1
    scale = 2.0
    x, u = ops.layernorm_lora_u_fwd(h.to(cuda_dev), gamma.to(cuda_dev), beta.to(cuda_dev), A.to(cuda_dev),
                                    idx.to(cuda_dev), n_sets, scale, 1e-5)
    x, u = x.cpu(), u.cpu()
    ref_ln = torch.nn.functional.layer_norm(h.float(), (d,), gamma.float(), beta.float(), 1e-5)
    assert (x.float() - ref_ln).abs().max().item() <= 2.0 ** -5
    x_plain = ops.layernorm_fwd(h.to(cuda_dev), gamma.to(cuda_dev), beta.to(cuda_dev), 1e-5).cpu()
    diff = x.float() - x_plain.float()
    assert diff.abs().max().item() <= 2.0 ** -7 * x_plain.float().abs().max().item()   # one bf16 ulp of the largest value
    assert (diff != 0).float().mean().item() <= 2e-3
    assert u.shape == (n_sets, B, T, r)
    for s in range(n_sets):
        for b in range(B):
            k = int(idx[b])
            if k < 0:
                continue                                             # base-only utterance: U rows are never read
            want = (scale * (x[b].float() @ A[s * n + k].float().t())).to(torch.bfloat16)
            assert rel_err(u[s, b], want) <= TIGHT, (s, b)
    # and against the tcgen05 U pass of the split path on the same x
    if n_sets * B * T >= 1:
        u_pass = ops.lora_u_fwd(x.to(cuda_dev), A.to(cuda_dev), idx.to(cuda_dev), n_sets, scale, d).cpu()
        for b in range(B):
            if int(idx[b]) >= 0:
                assert rel_err(u[:, b], u_pass[:, b]) <= TIGHT


def test_projection_with_precomputed_u_equals_the_two_launch_split_path(cuda_dev):
    """sar_attn_proj_fwd with SAR_FLAG_U_READY (U from the fused LayerNorm) vs its own U pass: same dense launch, U equal
    up to summation order -> outputs within one bf16 rounding."""
    B, T, d, r, n = 4, 1500, 768, 16, 4
    g = torch.Generator().manual_seed(5)
    dev = cuda_dev
    h = torch.randn(B, T, d, generator=g).to(torch.bfloat16).to(dev)
    gamma = torch.ones(d, dtype=torch.bfloat16, device=dev)
    beta = torch.zeros(d, dtype=torch.bfloat16, device=dev)
    W = (torch.randn(3 * d, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    bias = (torch.randn(3 * d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    A = ((torch.rand(2 * n, r, d, generator=g) * 2 - 1) / d ** 0.5).to(torch.bfloat16).to(dev)
    Bp = ops.pack_lora_b((torch.randn(2 * n, d, r, generator=g) * 0.02).to(torch.bfloat16).to(dev))
    idx = torch.tensor([0, 3, -1, 2], dtype=torch.int32, device=dev)
    x, u = ops.layernorm_lora_u_fwd(h, gamma, beta, A, idx, 2, 2.0)
    want = ops.attn_proj_fwd(x, W, bias, A, Bp, idx, [0, -1, 1], [1.0, 1.0, 1.0], 2, 2.0, split=True)
    got = ops.attn_proj_fwd(x, W, bias, A, Bp, idx, [0, -1, 1], [1.0, 1.0, 1.0], 2, 2.0, u=u)
    for a, b in zip(got, want):
        assert rel_err(a, b) <= TIGHT
    assert torch.equal(got[1], want[1])            # k_proj carries no adapter: bit-identical


# ------------------------------------------------------------------------------------------------ operand refresh
def test_operand_refresh_rederives_every_block_bit_exactly(cuda_dev):
    """sar_operand_refresh: dst = bf16(scale·src) over strided / transposed blocks of padded stacks, fp32 and bf16
    sources; padding stays untouched.  Bit-exact against the same arithmetic in torch (fp32 multiply, one rounding)."""
    import ctypes

    from speech_adapter_routing_b200 import _lib
    from speech_adapter_routing_b200.operand_refresh import _Desc

    dev = cuda_dev
    g = torch.Generator().manual_seed(5)
    r, rp, d_in, d_out = 12, 16, 384, 768
    wA = torch.randn(r, d_in, generator=g).to(dev)
    wB = (torch.randn(d_out, r, generator=g) * 0.02).to(dev)
    wB16 = wB.to(torch.bfloat16)
    A = torch.full((2, rp, d_in), 7.0, dtype=torch.bfloat16, device=dev)
    Bp = torch.full((2, d_out, 64), 7.0, dtype=torch.bfloat16, device=dev)
    At = torch.full((2, d_in, 64), 7.0, dtype=torch.bfloat16, device=dev)
    Bt = torch.full((2, rp, d_out), 7.0, dtype=torch.bfloat16, device=dev)
    jobs = [(wA, A[1, :r, :], 1.0), (wB, Bp[0, :, :r], 2.0), (wA.t(), At[1, :, :r], 1.0), (wB.t(), Bt[0, :r, :], 0.25),
            (wB16, Bp[1, :, :r], 0.125), (wB16.t(), Bt[1, :r, :], 1.0)]
    descs = []
    for src, dst, s in jobs:
        dt = _lib.SAR_DTYPE_F32 if src.dtype == torch.float32 else _lib.SAR_DTYPE_BF16
        descs.append(_Desc(src.data_ptr(), dst.data_ptr(), dst.shape[0], dst.shape[1], src.stride(0), src.stride(1),
                           dst.stride(0), dst.stride(1), s, dt))
    table = torch.frombuffer(bytearray(bytes((_Desc * len(descs))(*descs))), dtype=torch.uint8).to(dev)
    _lib.check(_lib.lib().sar_operand_refresh(table.data_ptr(), len(descs), max(d.numel() for _, d, _ in jobs),
                                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for src, dst, s in jobs:
        assert torch.equal(dst, (src.float() * s).to(torch.bfloat16))
    seven = torch.tensor(7.0, dtype=torch.bfloat16, device=dev)
    assert (A[0] == seven).all() and (A[1, r:] == seven).all() and (Bp[:, :, r:] == seven).all()
    assert (At[0] == seven).all() and (At[1, :, r:] == seven).all() and (Bt[:, r:] == seven).all()
    # empty table is a no-op; a null table with entries is rejected
    _lib.check(_lib.lib().sar_operand_refresh(None, 0, 1, torch.cuda.current_stream().cuda_stream))
    assert _lib.lib().sar_operand_refresh(None, 3, 16, torch.cuda.current_stream().cuda_stream) == _lib.SAR_EINVAL


@pytest.mark.parametrize("M,d", [(24000, 768), (37, 384), (5, 1280)])
def test_layernorm_with_statistics_matches_aten(cuda_dev, M, d):
    """sar_layernorm_fwd_stats: y bit-identical to sar_layernorm_fwd, mean / rstd equal to torch.native_layer_norm's
    (fp32 two-pass on the same bf16 input) within 1e-5 relative."""
    g = torch.Generator().manual_seed(d + M)
    x = (torch.randn(M, d, generator=g) * 2 + 0.5).to(torch.bfloat16).to(cuda_dev)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(torch.bfloat16).to(cuda_dev)
    b = (0.1 * torch.randn(d, generator=g)).to(torch.bfloat16).to(cuda_dev)
    y, mean, rstd = ops.layernorm_fwd_stats(x, w, b, 1e-5)
    assert torch.equal(y, ops.layernorm_fwd(x, w, b, 1e-5))
    ry, rmean, rrstd = torch.native_layer_norm(x, (d,), w, b, 1e-5)
    assert mean.shape == rmean.shape and rstd.shape == rrstd.shape and mean.dtype == rmean.dtype == torch.float32
    assert rel_err(mean, rmean) <= 1e-5 and rel_err(rstd, rrstd) <= 1e-5
    assert rel_err(y, ry) <= TIGHT
