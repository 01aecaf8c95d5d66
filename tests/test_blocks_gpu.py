"""GPU parity tests of the fused Whisper-block entry points (sar_attn_proj_fwd, sar_linear_fwd, sar_dense_fwd,
sar_layernorm_fwd) through the C ABI against oracle/blocks.py, and of the re-bound HF layer bodies.

Tolerances (bf16 storage, fp32 accumulation; written where used):
  projections / dense layers vs the same-rounding oracle: max|err| <= 2^-7 * max|ref|
  LayerNorm: one bf16 rounding of values of magnitude <= 8  ->  max|err| <= 2^-5 absolute
  GELU element-wise: |err| <= 2^-8 |ref| + 4e-5  (bf16 rounding + the 2.5e-5 fit error of the logistic-form erf-GELU)
  fused layer bodies vs HF's own bodies over the same K1 module slots: logits max|err| <= 3e-2 * max|ref|
"""
import pytest
import torch

from oracle import blocks as oblocks, fixtures
from speech_adapter_routing_b200 import ops

pytestmark = pytest.mark.gpu

TIGHT = 2.0 ** -7


def rel_err(y, ref):
    y, ref = y.float().cpu(), ref.float().cpu()
    return ((y - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def head_major_to_rows(y):
    B, h, T, e = y.shape
    return y.permute(0, 2, 1, 3).reshape(B, T, h * e)


# ------------------------------------------------------------------------------------------------ sar_attn_proj_fwd
PROJ_CASES = [
    # name, B, T, d, r, n, segments (lora?, out scale), grid override, base_only_every
    ("cross q", 3, 300, 768, 16, 4, [(True, 0.125)], 0, 0),
    ("cross k|v", 3, 300, 768, 16, 4, [(False, 1.0), (True, 1.0)], 0, 0),
    ("self q|k|v", 3, 300, 768, 16, 4, [(True, 0.125), (False, 1.0), (True, 1.0)], 0, 0),
    ("self q|k|v, ranges start mid-unit", 3, 256, 768, 16, 4, [(True, 0.125), (False, 1.0), (True, 1.0)], 10, 0),
    ("self q|k|v, base-only utterances", 6, 200, 768, 16, 4, [(True, 0.125), (False, 1.0), (True, 1.0)], 0, 3),
    ("self q|k|v, no adapters at all", 2, 256, 768, 16, 0, [(False, 0.125), (False, 1.0), (False, 1.0)], 0, 0),
    ("medium r32", 4, 200, 1024, 32, 4, [(True, 0.125), (False, 1.0), (True, 1.0)], 0, 0),
    ("large-v3 r64, 8 adapters", 4, 130, 1280, 64, 8, [(True, 0.125), (False, 1.0), (True, 1.0)], 0, 0),
    ("decoder length 128", 8, 128, 768, 16, 4, [(True, 0.125), (False, 1.0), (True, 1.0)], 0, 0),
    ("tiny d=384", 4, 100, 384, 16, 2, [(True, 0.125), (False, 1.0), (True, 1.0)], 0, 0),
    ("single frame T=1", 5, 1, 768, 16, 4, [(True, 0.125), (False, 1.0), (True, 1.0)], 0, 0),
]


def _proj_case(B, T, d, r, n, segs, bo, seed=4321):
    cases = [fixtures.make_lora_case(B, T, d, d, r, max(n, 1), seed=seed + 10 * i, base_only_every=bo)
             for i in range(len(segs))]
    x, idx = cases[0].x, cases[0].utt_adapter
    refs = oblocks.attn_projections(x, [c.W for c in cases], [c.bias for c in cases],
                                    [(c.A_stack, c.B_stack) if lo else None for c, (lo, _) in zip(cases, segs)],
                                    cases[0].scaling, idx, [s for _, s in segs], d // 64)
    seg_set, As, Bps = [], [], []
    for c, (lo, _) in zip(cases, segs):
        if lo:
            seg_set.append(len(As)); As.append(c.A_stack); Bps.append(ops.pack_lora_b(c.B_stack))
        else:
            seg_set.append(-1)
    return cases, x, idx, refs, seg_set, As, Bps


@pytest.mark.parametrize("name,B,T,d,r,n,segs,grid,bo", PROJ_CASES, ids=[c[0] for c in PROJ_CASES])
@pytest.mark.parametrize("head_major,split", [(True, True), (False, True), (True, False), (False, False)])
def test_attn_proj_matches_oracle(cuda_dev, name, B, T, d, r, n, segs, grid, bo, head_major, split):
    """split=True: U through a workspace + dense kernel with one extra K block; split=False: single launch, U in smem."""
    cases, x, idx, refs, seg_set, As, Bps = _proj_case(B, T, d, r, n, segs, bo)
    dev = cuda_dev
    W = torch.cat([c.W for c in cases], 0).to(dev)
    bias = torch.cat([c.bias for c in cases], 0).to(dev)
    A = torch.cat(As, 0).to(dev) if As else None
    Bp = torch.cat(Bps, 0).to(dev) if As else None
    ys = ops.attn_proj_fwd(x.to(dev), W, bias, A, Bp, idx.to(dev) if As else None, seg_set, [s for _, s in segs],
                           max(len(As), 1), cases[0].scaling, y_head_major=head_major, grid=grid, split=split)
    for y, ref in zip(ys, refs):
        if head_major:
            assert y.shape == (B, d // 64, T, 64)
            assert rel_err(y, ref) <= TIGHT
        else:
            assert y.shape == (B, T, d)
            assert rel_err(y, head_major_to_rows(ref)) <= TIGHT


def test_attn_proj_full_size_utterances_are_independent(cuda_dev):
    """BASELINE config 2 size (B=64, T=1500, d=768, 4 adapters r16): an utterance projected alone gives bit-identical
    rows, whatever the tile-to-SM schedule of the full batch was."""
    segs = [(True, 1.0), (False, 1.0), (True, 1.0)]
    cases, x, idx, _, seg_set, As, Bps = _proj_case(64, 1500, 768, 16, 4, segs, 5, seed=77)
    dev = cuda_dev
    W = torch.cat([c.W for c in cases], 0).to(dev)
    bias = torch.cat([c.bias for c in cases], 0).to(dev)
    A, Bp = torch.cat(As, 0).to(dev), torch.cat(Bps, 0).to(dev)
    full = ops.attn_proj_fwd(x.to(dev), W, bias, A, Bp, idx.to(dev), seg_set, [1.0] * 3, 2, 2.0)
    for b in (0, 5, 31, 63):
        one = ops.attn_proj_fwd(x[b:b + 1].to(dev), W, bias, A, Bp, idx[b:b + 1].to(dev), seg_set, [1.0] * 3, 2, 2.0)
        for yf, yo in zip(full, one):
            assert torch.equal(yf[b], yo[0])
    assert all(torch.isfinite(y.float()).all() for y in full)
    # the split path (U via workspace, 256-wide dense tiles) and the single-launch path (U in shared memory, 192-wide
    # tiles) round at the same points: identical bits
    fused = ops.attn_proj_fwd(x.to(dev), W, bias, A, Bp, idx.to(dev), seg_set, [1.0] * 3, 2, 2.0, split=False)
    for ys, yf in zip(full, fused):
        assert torch.equal(ys, yf)


def test_attn_proj_rejects_bad_arguments(cuda_dev):
    x = torch.zeros(1, 8, 768, dtype=torch.bfloat16, device=cuda_dev)
    W = torch.zeros(768 + 64, 768, dtype=torch.bfloat16, device=cuda_dev)   # d_out = 832 is not a multiple of 128
    with pytest.raises(Exception, match="multiple"):
        ops.attn_proj_fwd(x, W, None, None, None, None, [-1], [1.0], 1, 1.0)


# ------------------------------------------------------------------------------------------------ sar_linear_fwd
LINEAR_CASES = [
    # name, B, T, d_in, d_out, gelu, residual, head-major x, grid, in place
    ("plain", 2, 300, 768, 768, False, False, False, 0, False),
    ("fc1 + GELU", 2, 300, 768, 3072, True, False, False, 0, False),
    ("fc2 + residual, K=3072", 2, 300, 3072, 768, False, True, False, 0, False),
    ("out_proj from SDPA layout + residual", 3, 300, 768, 768, False, True, True, 0, False),
    ("out_proj, few pairs", 3, 256, 768, 768, False, True, True, 10, False),
    ("flattened rows", 1, 8192, 768, 3072, True, False, False, 0, False),
    ("large-v3 fc1", 2, 200, 1280, 5120, True, False, False, 0, False),
    ("medium fc2", 2, 200, 4096, 1024, False, True, False, 0, False),
    ("in-place residual", 2, 300, 768, 768, False, True, False, 0, True),
    ("one row", 1, 1, 768, 768, False, True, False, 0, False),
]


@pytest.mark.parametrize("name,B,T,d_in,d_out,gelu,res,hm,grid,inplace", LINEAR_CASES, ids=[c[0] for c in LINEAR_CASES])
def test_linear_matches_oracle(cuda_dev, name, B, T, d_in, d_out, gelu, res, hm, grid, inplace):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, d_in, generator=g).to(torch.bfloat16)
    W = (torch.randn(d_out, d_in, generator=g) * 0.02).to(torch.bfloat16)
    b = (torch.randn(d_out, generator=g) * 0.02).to(torch.bfloat16)
    r = torch.randn(B, T, d_out, generator=g).to(torch.bfloat16) if res else None
    ref = oblocks.dense(x, W, b, r, gelu)
    xd = x.to(cuda_dev)
    if hm:
        xd = xd.view(B, T, d_in // 64, 64).permute(0, 2, 1, 3).contiguous()
    rd = None if r is None else r.to(cuda_dev)
    y = ops.linear_fwd(xd, W.to(cuda_dev), b.to(cuda_dev), rd, int(gelu), x_head_major=hm,
                       out=rd if inplace else None, grid=grid)
    assert y.shape == (B, T, d_out)
    assert rel_err(y, ref) <= TIGHT


def test_gelu_epilogue_elementwise_accuracy(cuda_dev):
    """W = I makes the GEMM exact, so y = bf16(gelu(x)) element by element over the whole useful range and beyond."""
    d = 128
    v = torch.cat([torch.linspace(-12, 12, 256 * d - 8), torch.tensor([0.0, -0.0, 30.0, -30.0, 1e-3, -1e-3, 100.0, -100.0])])
    x = v.to(torch.bfloat16).view(1, -1, d)
    y = ops.linear_fwd(x.to(cuda_dev), torch.eye(d, dtype=torch.bfloat16, device=cuda_dev), None, None, 1).float().cpu()
    ref = torch.nn.functional.gelu(x.float())
    err = (y - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 4e-5).all()), err.max().item()


# ------------------------------------------------------------------------------------------------ sar_dense_fwd
@pytest.mark.parametrize("M,d,V", [(300, 768, 5001), (1024, 768, 51865), (130, 1280, 2050)])
def test_dense_lm_head_ragged_vocab(cuda_dev, M, d, V):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(M, d, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, d, generator=g) * 0.02).to(torch.bfloat16)
    ldy = (V + 7) // 8 * 8
    buf = torch.full((M + 1, ldy), 7.0, dtype=torch.bfloat16, device=cuda_dev)    # one guard row after the output
    ops.dense_fwd(x.to(cuda_dev), d, 0, W.to(cuda_dev), None, buf, ldy, 0, 1, M, d, V)
    ref = (x.to(cuda_dev).float() @ W.to(cuda_dev).float().t()).cpu()             # fp32 oracle evaluated on the device
    assert rel_err(buf[:M, :V], ref) <= TIGHT
    pad = buf[:M, V:]
    assert bool(((pad == 7.0) | (pad == 0.0)).all())     # TMA clips stores at 16-byte granularity
    assert bool((buf[M] == 7.0).all())                   # nothing past the last row


@pytest.mark.parametrize("B,C,L,d", [(2, 80, 3000, 384), (3, 80, 3000, 768), (2, 128, 3000, 1280)])
def test_dense_conv_frontend_as_gemm(cuda_dev, B, C, L, d):
    """conv1 / conv2 of the Whisper encoder as GEMMs over overlapping rows of zero-padded channels-last frame buffers."""
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, C, L, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(d, C, 3, generator=g) * (C * 3) ** -0.5).to(torch.bfloat16)
    b1 = (torch.randn(d, generator=g) * 0.1).to(torch.bfloat16)
    w2 = (torch.randn(d, d, 3, generator=g) * (d * 3) ** -0.5).to(torch.bfloat16)
    b2 = (torch.randn(d, generator=g) * 0.1).to(torch.bfloat16)
    pos = (torch.randn(L // 2, d, generator=g) * 0.1).to(torch.bfloat16)
    ref = oblocks.conv_frontend(x, w1, b1, w2, b2, pos)
    dev = cuda_dev
    buf1 = torch.zeros(B, L + 2, C, dtype=torch.bfloat16, device=dev)
    buf2 = torch.zeros(B, L + 2, d, dtype=torch.bfloat16, device=dev)
    buf1[:, 1:L + 1].copy_(x.to(dev).transpose(1, 2))
    W1 = w1.to(dev).permute(0, 2, 1).reshape(d, -1).contiguous()
    W2 = w2.to(dev).permute(0, 2, 1).reshape(d, -1).contiguous()
    ops.dense_fwd(buf1, C, (L + 2) * C, W1, b1.to(dev), buf2[:, 1:], d, (L + 2) * d, B, L, 3 * C, d, act=1)
    h = torch.empty(B, L // 2, d, dtype=torch.bfloat16, device=dev)
    ops.dense_fwd(buf2, 2 * d, (L + 2) * d, W2, b2.to(dev), h, d, (L // 2) * d, B, L // 2, 3 * d, d, act=1,
                  residual=pos.to(dev), ldr=d, res_broadcast=True)
    assert rel_err(h, ref) <= TIGHT
    assert bool((buf2[:, 0] == 0).all()) and bool((buf2[:, L + 1] == 0).all())     # the padding frames stay zero


# ------------------------------------------------------------------------------------------------ sar_layernorm_fwd
@pytest.mark.parametrize("M,d", [(7, 64), (1000, 384), (96000, 768), (515, 1024), (300, 1280), (33, 2048)])
def test_layernorm_matches_oracle(cuda_dev, M, d):
    g = torch.Generator().manual_seed(12)
    x = (torch.randn(M, d, generator=g) * 2 + 0.5).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(torch.bfloat16)
    b = (0.1 * torch.randn(d, generator=g)).to(torch.bfloat16)
    y = ops.layernorm_fwd(x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev), 1e-5)
    ref = oblocks.layer_norm(x, w, b)
    assert (y.float().cpu() - ref).abs().max().item() <= 2.0 ** -5
    with pytest.raises(Exception, match="multiple of 8"):
        ops.layernorm_fwd(torch.zeros(4, 12, dtype=torch.bfloat16, device=cuda_dev),
                          torch.zeros(12, dtype=torch.bfloat16, device=cuda_dev),
                          torch.zeros(12, dtype=torch.bfloat16, device=cuda_dev))


# ------------------------------------------------------------------------------------------------ re-bound layer bodies
def test_fused_layer_bodies_match_hf_bodies(cuda_dev):
    """Whole model, mixed adapters incl. a base-only utterance: the fused blocks (conv-as-GEMM front-end, fused
    projections, epilogue-fused residual / GELU, own LayerNorm, padded lm head) vs HF's bodies over the same K1 slots."""
    import speech_adapter_routing_b200 as sar
    from speech_adapter_routing_b200 import whisper_blocks
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    cfg = WhisperConfig(vocab_size=1001, num_mel_bins=80, d_model=384, encoder_layers=2, decoder_layers=2,
                        encoder_attention_heads=6, decoder_attention_heads=6, encoder_ffn_dim=1536,
                        decoder_ffn_dim=1536, max_source_positions=1500, max_target_positions=448,
                        pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=3)
    torch.manual_seed(0)
    model = WhisperForConditionalGeneration(cfg).to(torch.bfloat16).to(cuda_dev).eval()
    langs = ["a", "b", "c"]
    for l in langs:
        sar.inject_lora(model, sar.LoraConfig(r=16, lora_alpha=32, target_modules=["q_proj", "v_proj"]), adapter_name=l)
    g = torch.Generator().manual_seed(3)
    for m in sar.lora_modules(model).values():
        for l in langs:
            m.lora_B[l].weight.data.copy_((torch.randn(m.out_features, 16, generator=g) * 0.05).to(cuda_dev))
    x = torch.randn(5, 80, 3000, generator=g).to(torch.bfloat16).to(cuda_dev)
    dec = torch.randint(4, 1001, (5, 37), generator=g).to(cuda_dev)
    idx = torch.tensor([0, 2, -1, 1, 2], dtype=torch.int32, device=cuda_dev)
    try:
        with torch.no_grad(), sar.route(idx):
            ops.reset_counters()
            fused = model(input_features=x, decoder_input_ids=dec, use_cache=False).logits.float()
            counts = dict(ops.LAUNCHES)
            whisper_blocks.FUSED_BLOCKS_ENABLED = False
            ops.reset_counters()
            plain = model(input_features=x, decoder_input_ids=dec, use_cache=False).logits.float()
            plain_counts = dict(ops.LAUNCHES)
    finally:
        whisper_blocks.FUSED_BLOCKS_ENABLED = True
    assert counts["proj"] >= 2 + 2 * 3 and counts["k1"] == 0        # >= one launch per fused projection call
    assert plain_counts["proj"] == 0 and plain_counts["k1"] == 12   # 2*2 + 2*4 module-slot calls
    assert fused.shape == plain.shape == (5, 37, 1001)
    assert rel_err(fused, plain) <= 3e-2


# ------------------------------------------------------------------------------------------------ decode-step rows form
@pytest.mark.parametrize("M,d,r,n,segs", [
    (64, 768, 16, 4, [(True, 0.125), (False, 1.0), (True, 1.0)]),      # self-attention q|k|v of a decode step, B = 64
    (5, 768, 16, 4, [(True, 0.125)]),                                   # cross-attention q
    (130, 1280, 64, 8, [(True, 0.125), (False, 1.0), (True, 1.0)]),     # more rows than one 128-row tile half
    (16, 1024, 32, 4, [(False, 1.0), (True, 1.0)]),
])
def test_attn_proj_rows_matches_oracle(cuda_dev, M, d, r, n, segs):
    """sar_attn_proj_fwd_rows: one token per utterance, rows of different adapters (and base-only rows) share a tile."""
    cases, x, idx, refs, seg_set, As, Bps = _proj_case(M, 1, d, r, n, segs, 5)
    dev = cuda_dev
    W = torch.cat([c.W for c in cases], 0).to(dev)
    bias = torch.cat([c.bias for c in cases], 0).to(dev)
    A, Bp = torch.cat(As, 0).to(dev), torch.cat(Bps, 0).to(dev)
    ys = ops.attn_proj_fwd_rows(x.view(M, d).to(dev), W, bias, A, Bp, idx.to(dev), seg_set, [s for _, s in segs],
                                len(As), cases[0].scaling)
    for y, ref in zip(ys, refs):
        assert y.shape == (M, d)
        # two bf16 roundings (base GEMM result, then + low-rank update): 2^-6 instead of 2^-7
        assert rel_err(y.view(M, d // 64, 1, 64), ref) <= 2.0 ** -6


def test_attn_proj_large_v3_full_size(cuda_dev):
    """BASELINE config 4 shape on one GPU (whisper-large-v3: d = 1280, 8 adapters r64, 64 clips x 1500 frames): the
    split and single-launch paths agree bit for bit, utterances are independent, base-only utterances equal the dense
    projection."""
    segs = [(True, 1.0), (False, 1.0), (True, 1.0)]
    B, T, d, r, n = 64, 1500, 1280, 64, 8
    g = torch.Generator().manual_seed(99)
    dev = cuda_dev
    x = torch.randn(B, T, d, generator=g).to(torch.bfloat16).to(dev)
    W = (torch.randn(3 * d, d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    bias = (torch.randn(3 * d, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    A = ((torch.rand(2 * n, r, d, generator=g) * 2 - 1) * d ** -0.5).to(torch.bfloat16).to(dev)
    Bp = ops.pack_lora_b((torch.randn(2 * n, d, r, generator=g) * 0.02).to(torch.bfloat16).to(dev))
    idx = torch.randint(0, n, (B,), generator=g).to(torch.int32)
    idx[::7] = -1
    idx = idx.to(dev)
    split = ops.attn_proj_fwd(x, W, bias, A, Bp, idx, [0, -1, 1], [1.0] * 3, 2, 2.0, split=True)
    fused = ops.attn_proj_fwd(x, W, bias, A, Bp, idx, [0, -1, 1], [1.0] * 3, 2, 2.0, split=False)
    dense = ops.attn_proj_fwd(x, W, bias, None, None, None, [-1, -1, -1], [1.0] * 3, 1, 2.0)
    for ys, yf, yd in zip(split, fused, dense):
        assert torch.equal(ys, yf)
        assert torch.equal(ys[::7], yd[::7])                       # base-only utterances
        assert torch.isfinite(ys.float()).all()
    assert torch.equal(split[1], dense[1])                         # k_proj carries no adapter
    assert not torch.equal(split[0][1], dense[0][1])               # an adapted utterance differs from the base
    b = 33
    one = ops.attn_proj_fwd(x[b:b + 1], W, bias, A, Bp, idx[b:b + 1], [0, -1, 1], [1.0] * 3, 2, 2.0)
    for ys, yo in zip(split, one):
        assert torch.equal(ys[b], yo[0])
    # fp32 reference of one utterance's q segment, evaluated on the device
    k = int(idx[b])
    u = ((x[b].float() @ A[k].float().t()) * 2.0).to(torch.bfloat16).float()
    ref = x[b].float() @ W[:d].float().t() + bias[:d].float() + (u @ Bp[k, :, :r].float().t() if k >= 0 else 0)
    assert rel_err(head_major_to_rows(split[0][b:b + 1])[0], ref) <= TIGHT


# ------------------------------------------------------------------------------------------------ decode self-attention
@pytest.mark.parametrize("B,H,Tmax,pos", [(3, 6, 64, 0), (64, 12, 128, 37), (5, 20, 448, 447), (2, 12, 256, 130)])
def test_decode_self_attn_matches_sdpa(cuda_dev, B, H, Tmax, pos):
    """sar_decode_self_attn vs an fp32 softmax(q kᵀ) v over positions 0..pos; the cache row at ``pos`` is written."""
    g = torch.Generator().manual_seed(5)
    dev = cuda_dev
    K = torch.randn(B, H, Tmax, 64, generator=g).to(torch.bfloat16).to(dev)
    V = torch.randn(B, H, Tmax, 64, generator=g).to(torch.bfloat16).to(dev)
    q = (torch.randn(B, H, 64, generator=g) * 0.4).to(torch.bfloat16).to(dev)
    kn = torch.randn(B, H, 64, generator=g).to(torch.bfloat16).to(dev)
    vn = torch.randn(B, H, 64, generator=g).to(torch.bfloat16).to(dev)
    K0, V0 = K.clone(), V.clone()
    p = torch.tensor([pos], dtype=torch.long, device=dev)
    out = ops.decode_self_attn(q, kn, vn, K, V, p)
    K0[:, :, pos], V0[:, :, pos] = kn, vn
    assert torch.equal(K, K0) and torch.equal(V, V0)          # exactly one cache row written, bit-exact
    ref = oblocks.attention(q.cpu()[:, :, None], K0.cpu(), V0.cpu(), n_keys=pos + 1).reshape(B, H * 64)
    assert out.shape == (B, H * 64)
    assert rel_err(out, ref) <= TIGHT


@pytest.mark.parametrize("B,H,Tk", [(3, 6, 1), (64, 12, 1500), (5, 20, 1500), (2, 12, 77), (1, 8, 4096)])
def test_decode_cross_attn_matches_fp32_softmax(cuda_dev, B, H, Tk):
    """sar_decode_cross_attn (one query token per utterance over read-only encoder K / V) vs an fp32 softmax(q kᵀ) v;
    K / V must come back untouched."""
    g = torch.Generator().manual_seed(6)
    dev = cuda_dev
    K = torch.randn(B, H, Tk, 64, generator=g).to(torch.bfloat16).to(dev)
    V = torch.randn(B, H, Tk, 64, generator=g).to(torch.bfloat16).to(dev)
    q = (torch.randn(B, H, 1, 64, generator=g) * 0.4).to(torch.bfloat16).to(dev)
    K0, V0 = K.clone(), V.clone()
    out = ops.decode_cross_attn(q, K, V)
    assert torch.equal(K, K0) and torch.equal(V, V0)
    ref = oblocks.attention(q.cpu(), K.cpu(), V.cpu()).reshape(B, H * 64)
    assert out.shape == (B, H * 64)
    assert rel_err(out, ref) <= TIGHT
    out2 = ops.decode_cross_attn(q, K, V)
    assert torch.equal(out, out2)                                # fixed-order reduction: deterministic


# ------------------------------------------------------------------------------------------------ skinny (<= 128 rows)
@pytest.mark.parametrize("M,d_in,d_out,gelu,res", [
    (64, 768, 3072, True, False),      # fc1 + GELU of a decode step, B = 64
    (64, 3072, 768, False, True),      # fc2 + residual: 48 K blocks through a 10-stage ring
    (64, 768, 768, False, True),       # out_proj + residual
    (1, 768, 768, False, False),       # a single row
    (100, 1280, 5120, True, False),    # large-v3 fc1, rows not a multiple of 32
    (128, 1024, 1024, False, True),    # exactly one full tile of rows
    (33, 384, 1536, True, False),      # tiny geometry
])
def test_skinny_dense_matches_oracle(cuda_dev, M, d_in, d_out, gelu, res):
    """<= 128 rows route to the weight-streaming single-CTA kernel (skinny_fwd.cu): same contract as sar_linear_fwd."""
    g = torch.Generator().manual_seed(13)
    x = torch.randn(1, M, d_in, generator=g).to(torch.bfloat16)
    W = (torch.randn(d_out, d_in, generator=g) * 0.02).to(torch.bfloat16)
    b = (torch.randn(d_out, generator=g) * 0.02).to(torch.bfloat16)
    r = torch.randn(1, M, d_out, generator=g).to(torch.bfloat16) if res else None
    ref = oblocks.dense(x, W, b, r, gelu)
    guard = torch.full((1, M + 2, d_out), 3.0, dtype=torch.bfloat16, device=cuda_dev)   # rows past M must stay untouched
    y = ops.linear_fwd(x.to(cuda_dev), W.to(cuda_dev), b.to(cuda_dev), None if r is None else r.to(cuda_dev), int(gelu),
                       out=guard[:, :M])
    assert rel_err(y, ref) <= TIGHT
    assert bool((guard[:, M:] == 3.0).all())
    # identical to the CTA-pair kernel on the same operands up to one bf16 ulp of the largest value
    y2 = ops.linear_fwd(x.to(cuda_dev), W.to(cuda_dev), b.to(cuda_dev), None if r is None else r.to(cuda_dev), int(gelu),
                        block_n=128)
    assert rel_err(y, y2.float()) <= 2.0 ** -8


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,H,Tq,Tk,causal", [
    (1, 1, 128, 64, False), (2, 3, 200, 300, False), (1, 2, 1500, 1500, False), (2, 12, 128, 1500, False),
    (2, 3, 128, 128, True), (2, 2, 37, 37, True), (1, 2, 300, 300, True), (3, 2, 1, 77, False), (1, 20, 448, 448, True),
    # long non-causal: the two-query-tile kernel (attn_fwd2.cu): ragged last key tile, second query tile partly / fully
    # past Tq, cross-attention with Tq != Tk
    (2, 3, 384, 1500, False), (1, 2, 448, 1500, False), (1, 3, 513, 200, False), (1, 2, 1024, 64, False),
    (1, 1, 640, 1, False),
])
def test_attn_fwd_matches_fp32_softmax(cuda_dev, B, H, Tq, Tk, causal):
    """sar_attn_fwd (tcgen05 flash-attention forward, head dim 64) vs fp32 softmax(q kᵀ) v on the same bf16 inputs:
    ragged last key tile, query tiles past Tq, causal diagonal, one-row queries.  Tolerance: P is rounded to bf16 before
    the PV product (like every flash-attention kernel): 2^-7 of max|ref|."""
    g = torch.Generator().manual_seed(31)
    q = (torch.randn(B, H, Tq, 64, generator=g) * 0.35).to(torch.bfloat16).to(cuda_dev)
    k = torch.randn(B, H, Tk, 64, generator=g).to(torch.bfloat16).to(cuda_dev)
    v = torch.randn(B, H, Tk, 64, generator=g).to(torch.bfloat16).to(cuda_dev)
    ref = oblocks.attention(q.cpu(), k.cpu(), v.cpu(), causal=causal)   # oracle pinned against HF's eager attention
    out = ops.attn_fwd(q, k, v, causal)
    assert out.shape == q.shape and out.dtype == torch.bfloat16
    assert rel_err(out, ref) <= TIGHT
    # a (b, h) pair computed alone gives identical bits (tiles never cross heads)
    one = ops.attn_fwd(q[:1, :1].contiguous(), k[:1, :1].contiguous(), v[:1, :1].contiguous(), causal)
    assert torch.equal(one[0, 0], out[0, 0])


@pytest.mark.parametrize("M,d_in,d_out", [(24000, 768, 3072), (300, 384, 1536), (130, 128, 256)])
def test_linear_with_gelu_backward_epilogue(cuda_dev, M, d_in, d_out):
    """SAR_ACT_GELU_BWD: y = (x·Wᵀ) * GELU'(pre) — the dX GEMM of fc2 with the GELU backward in its epilogue — against
    torch's erf-form gelu_backward on the fp32 product (one bf16 rounding of the result: 2^-7 of the largest value)."""
    from speech_adapter_routing_b200._lib import SAR_ACT_GELU_BWD

    g = torch.Generator().manual_seed(M + d_out)
    x = (torch.randn(1, M, d_in, generator=g) * 0.5).to(torch.bfloat16).to(cuda_dev)
    W = (torch.randn(d_out, d_in, generator=g) * 0.05).to(torch.bfloat16).to(cuda_dev)
    pre = (torch.randn(1, M, d_out, generator=g) * 1.5).to(torch.bfloat16).to(cuda_dev)
    y = ops.linear_fwd(x, W, None, residual=pre, act=SAR_ACT_GELU_BWD)
    ref = torch.ops.aten.gelu_backward(x.float() @ W.float().t(), pre.float())
    assert rel_err(y, ref) <= TIGHT
    # far tails: GELU' -> 0 / 1 without NaNs
    pre2 = torch.full_like(pre, 30.0); pre2[..., ::2] = -30.0
    y2 = ops.linear_fwd(x, W, None, residual=pre2, act=SAR_ACT_GELU_BWD)
    ref2 = torch.ops.aten.gelu_backward(x.float() @ W.float().t(), pre2.float())
    assert torch.isfinite(y2.float()).all() and rel_err(y2, ref2) <= TIGHT
