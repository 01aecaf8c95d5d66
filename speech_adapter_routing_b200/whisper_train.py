"""Training path of the fused Whisper blocks: forward AND backward of a whole encoder / decoder layer on libsar.

The reference trains with HF's eager layer bodies under autograd (src/training/trainer.py:251-256, bf16 autocast,
gradient checkpointing on by default, src/models/whisper_lora.py:81-83): ~45 launches per layer forward, twice, plus the
autograd backward.  Only the LoRA tensors train; the base weights are frozen, so the backward needs exactly

    dX through every dense layer                      -> the tcgen05 pair kernel on the transposed weight (sar_linear_fwd)
    GELU backward                                     -> the epilogue of fc2's dX GEMM (SAR_ACT_GELU_BWD)
    dX, dA, dB at q_proj / v_proj                     -> K3 (sar_qv_lora_bwd), accumulating into the flat gradient bucket
    LayerNorm, softmax(QKᵀ)V backward                 -> ATen kernels (library) on the tensors the forward saved

Each layer is ONE ``torch.autograd.Function``: its forward runs the fused launches of inference (row-major projection
outputs viewed as heads, LayerNorm on the own kernel with its statistics written out for ATen's LayerNorm backward) and
hands every tensor the backward needs to ``ctx.save_for_backward`` — so HF's gradient checkpointing really drops and
recomputes them.  The HF modules, their parameters and their attribute paths are untouched; when a precondition fails
(dropout active, masks, KV cache, non-bf16, LoRA on other modules) the layer keeps HF's body over the K1 / K3 module
slots (``REFUSED`` says why).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops
from ._lib import SAR_ACT_GELU_BWD, SAR_ACT_NONE
from .lora_linear import RoutedLoRALinear, _notify_grad_ready
from .routing import current_mix_weights

ENABLED = __import__("os").environ.get("SAR_FUSED_TRAIN", "1") != "0"
# softmax(QKᵀ)V forward / backward: "cudnn" or "flash" — ATen's fused-attention ops, called directly (no nested autograd
# graph: one host call each way, CUDA-graph capturable, and every saved tensor can go through ctx.save_for_backward).
SDPA_IMPL = __import__("os").environ.get("SAR_TRAIN_SDPA", "cudnn")
FUSED_GELU_BWD = __import__("os").environ.get("SAR_TRAIN_FUSED_GELU_BWD", "1") != "0"   # A/B switch
OWN_LN = __import__("os").environ.get("SAR_TRAIN_OWN_LN", "1") != "0"      # A/B switch: LayerNorm forward on sar_layernorm_fwd_stats
# Set by train_graph.GraphedTrainStep around its calls: under CUDA-graph capture HF materialises the decoder's causal mask
# as a tensor instead of passing None (transformers/masking_utils.py:262-275 refuses to skip it while "tracing"); the
# step passes no padding mask, so a square 4-D mask is known to be exactly the causal one and the fused layer may use
# is_causal.
ASSUME_CAUSAL_MASK = False
TRACE = None        # debugging: a list collects (label, |t|.sum()) of intermediates (device scalars: CUDA-graph safe)


def _trace(label: str, *ts) -> None:
    if TRACE is not None:
        for i, t in enumerate(ts):
            if t is not None:
                TRACE.append((f"{label}[{i}]", t.detach().float().abs().sum()))


CALLS = {"encoder_layers": 0, "decoder_layers": 0}   # fused-layer forwards taken (tests assert the path is live)


# ------------------------------------------------------------------------------------------------ small pieces
def _ln_fwd(x: torch.Tensor, ln) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(y, mean, rstd): the own LayerNorm kernel with its statistics written out (12 us for 24 000 x 768 rows; ATen's
    vectorized_layer_norm_kernel needs 28), ATen's for parameter dtypes / widths the kernel does not take."""
    w, b = ln.weight, ln.bias
    if (OWN_LN and w.dtype == torch.bfloat16 and b is not None and b.dtype == torch.bfloat16 and x.shape[-1] % 8 == 0
            and x.shape[-1] <= 2048):
        return ops.layernorm_fwd_stats(x, w.detach(), b.detach(), ln.eps)
    return torch.native_layer_norm(x, (x.shape[-1],), w, b, ln.eps)


def _ln_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, ln) -> torch.Tensor:
    return torch.ops.aten.native_layer_norm_backward(dy.contiguous(), x, [x.shape[-1]], mean, rstd, ln.weight, ln.bias,
                                                     [True, False, False])[0]


def _wt(pack) -> torch.Tensor:
    """bf16 transpose of a dense layer's weight ([d_in, d_out] = the "weight" of the dX GEMM), cached on the pack."""
    p = pack.get()
    if getattr(pack, "_wt_key", None) != pack.key:
        pack._wt_key = pack.key
        pack._wt = p.W.t().contiguous()
    return pack._wt


def _dense_dx(dy: torch.Tensor, pack, head_major_out: bool = False) -> torch.Tensor:
    """dX = dY·W for a frozen dense layer: the pair kernel on the transposed weight, rows flattened to one utterance."""
    B, T, d_out = dy.shape
    dx = ops.linear_fwd(dy.reshape(1, B * T, d_out), _wt(pack), None)
    return dx.view(B, T, -1)


def _to_rows(t_hm: torch.Tensor) -> torch.Tensor:        # [B, h, T, 64] -> [B, T, h*64]
    B, H, T, hd = t_hm.shape
    return t_hm.transpose(1, 2).reshape(B, T, H * hd)


def _to_heads(t: torch.Tensor, H: int) -> torch.Tensor:
    """[B, T, d] -> [B, h, T, 64] as a strided VIEW (no copy).  This is the layout autograd hands dO to SDPA's backward in
    HF's own layer body (grad of ``attn_output.transpose(1, 2).contiguous()``), and it has to be: torch's cuDNN attention
    backward caches its graph per (q, k, v) shape/stride key WITHOUT dO's strides, so a process that mixes dO layouts for
    the same q / k / v layout silently gets wrong gradients from whichever layout came second (measured: 100 % errors in
    HF-body training after a fused step and vice versa).  Matching HF's layout keeps both paths on one valid graph."""
    B, T, d = t.shape
    return t.view(B, T, H, d // H).transpose(1, 2)


def _lora_params(mods: List[RoutedLoRALinear]) -> List[torch.Tensor]:
    ws: List[torch.Tensor] = []
    for m in mods:
        ws += [m.lora_A[n].weight for n in m.adapter_order] + [m.lora_B[n].weight for n in m.adapter_order]
    return ws


def _direct_views(m: RoutedLoRALinear, st, device):
    """(dA, dB) views of the flat gradient bucket that K3 may accumulate into in place for module ``m`` — single adapter
    whose parameters' ``.grad`` are fp32 slices of a dist.FlatGradBucket — or None."""
    names = m.adapter_order
    n, rp = st["A"].shape[0], st["A"].shape[1]
    if not (n == 1 and m.r[names[0]] == rp and st["grad_a_gain"][0] == 1.0 and st["grad_b_gain"][0] == 1.0):
        return None
    wA, wB = m.lora_A[names[0]].weight, m.lora_B[names[0]].weight
    gA, gB = wA.grad, wB.grad
    if (getattr(wA, "_sar_direct_grad", False) and getattr(wB, "_sar_direct_grad", False) and wA.requires_grad
            and wB.requires_grad and gA is not None and gB is not None and gA.dtype == torch.float32
            and gB.dtype == torch.float32 and gA.is_contiguous() and gB.is_contiguous() and gA.device == device):
        return gA.view(1, rp, m.in_features), gB.view(1, m.out_features, rp)
    return None


def _all_direct(mods: List[RoutedLoRALinear], device) -> bool:
    return bool(mods) and all(_direct_views(m, m._stacks(), device) is not None for m in mods)


def _lora_bwd(m: RoutedLoRALinear, dy: torch.Tensor, x: torch.Tensor, u: torch.Tensor, idx: torch.Tensor):
    """K3 for one LoRA'd projection: returns (dx, grads) with ``grads`` aligned to ``_lora_params([m])`` — all None when
    K3 accumulated straight into the parameters' bucket slices (single adapter under dist.FlatGradBucket)."""
    st = m._stacks(backward=True)
    n, rp = st["A"].shape[0], st["A"].shape[1]
    names = m.adapter_order
    direct = _direct_views(m, st, x.device)
    if direct is not None:
        dA, dB = direct
    else:
        dA = torch.zeros(n, rp, m.in_features, dtype=torch.float32, device=x.device)
        dB = torch.zeros(n, m.out_features, rp, dtype=torch.float32, device=x.device)
    dx = ops.qv_lora_bwd(dy.contiguous(), x, u, st["Wt"], st["At"], st["Bt"], idx, dA, dB, st["scale"], need_dx=True)
    if direct is not None:
        _notify_grad_ready(m.lora_A[names[0]].weight, m.lora_B[names[0]].weight)
        return dx, [None] * (2 * n)
    grads: List[Optional[torch.Tensor]] = []
    for k, name in enumerate(names):
        w = m.lora_A[name].weight
        grads.append((dA[k, : m.r[name]] * st["grad_a_gain"][k]).to(w.dtype) if w.requires_grad else None)
    for k, name in enumerate(names):
        w = m.lora_B[name].weight
        grads.append((dB[k, :, : m.r[name]] * st["grad_b_gain"][k]).to(w.dtype) if w.requires_grad else None)
    return dx, grads


# Tensors one attention keeps for its backward, in save order.  They all travel through ctx.save_for_backward, so under
# HF's gradient checkpointing (non-reentrant: the first forward runs with grad enabled) the saved-tensor hooks really drop
# them and the recompute really refills them — Python attributes on ctx would keep every activation alive.
_ATTN_FIELDS = ("x_q", "x_kv", "u_q", "u_kv", "q", "k", "v", "o", "lse", "cq", "ck", "seed", "off", "g_q", "g_kv")
_N_ATTN = len(_ATTN_FIELDS)


def _heads(t: torch.Tensor, H: int) -> torch.Tensor:
    """[B, T, d] -> [B, h, T, 64] strided view (HF's own q / k / v layout: no copy in either direction)."""
    B, T, d = t.shape
    return t.view(B, T, H, d // H).transpose(1, 2)


def _lora_u(proj, x, idx):
    """U = scale·x·A_kᵀ planes of one fused projection call ([n_sets, B, T, r]): the forward's low-rank operand and K3's
    ``u``."""
    if idx is None or proj.A is None:
        return None
    return ops.lora_u_fwd(x, proj.A, idx, proj.n_sets, proj.scale, proj.W.shape[0] // len(proj.mods))


def _active_dropout(m):
    """The nn.Dropout PEFT would apply to this module's LoRA input right now (training mode, p > 0), else None."""
    if not (isinstance(m, RoutedLoRALinear) and m.training and m._dropout_active()):
        return None
    return m.lora_dropout[m.active_adapter]


def _lora_u_dropout(proj, x, idx, u):
    """PEFT drops the input of the LoRA branch only: y = base(x) + s·B(A(drop(x))) (src/models/whisper_lora.py:30, default
    p = 0.1), one independent mask per module.  With g = drop(1) - 1 (-1 where dropped, p/(1-p) where kept) drop(x) =
    x + x∘g, so U gains s·(x∘g)·Aᵀ — a [M, d] x [d, r] product per module on top of the fused U pass; everything
    downstream (the dense launch with the low-rank K block) is unchanged.  Returns the per-set masks g stacked
    [n_sets, B, T, d] (None when no module of the call drops)."""
    drops = [_active_dropout(m) for m in proj.lora_mods]
    if u is None or not any(d is not None for d in drops):
        return None
    B, T, d = x.shape
    gs = torch.empty(len(drops), B, T, d, dtype=x.dtype, device=x.device)
    ones = torch.ones_like(x)                                          # shared by the modules of this call (read-only)
    for i, (m, drop) in enumerate(zip(proj.lora_mods, drops)):
        if drop is None:
            gs[i].zero_()
            continue
        g = torch.sub(drop(ones), 1, out=gs[i])
        st = m._stacks()
        u[i].add_(((x * g).view(B * T, d) @ st["A"][0].t()).view(B, T, -1), alpha=st["scale"])
    return gs


def _lora_dropout_bwd(m, dy, x, g, part, grads, idx):
    """Backward of the extra term of ``_lora_u_dropout`` for one module (K3 already did the drop-free part, including dB,
    which only sees U): dA += vᵀ(x∘g), dx += (v·A)∘g with v = s·dy·B."""
    st = m._stacks(backward=True)
    B, T, d = x.shape
    M = B * T
    v = (dy.reshape(M, -1) @ st["Bt"][0].t()) * st["scale"]           # [M, r]
    dA_c = (v.t() @ (x * g).view(M, d)).float()                        # [r, d]
    part = part + ((v @ st["A"][0]).view(B, T, d) * g)
    direct = _direct_views(m, st, x.device)
    if direct is not None:
        direct[0][0].add_(dA_c)
    else:
        w = m.lora_A[m.adapter_order[0]].weight
        if grads[0] is not None:
            grads[0] = grads[0] + dA_c[: w.shape[0]].to(grads[0].dtype)
    return part, grads


def _sdpa_fwd(q, k, v, is_causal: bool):
    """softmax(QKᵀ)V on ATen's fused-attention ops, called directly (one host call, CUDA-graph capturable).  Returns
    (o, [lse, cum_seq_q, cum_seq_k, seed, offset], (kind, max_q, max_k))."""
    global SDPA_IMPL
    aten = torch.ops.aten
    if SDPA_IMPL == "cudnn":
        try:
            r = aten._scaled_dot_product_cudnn_attention(q, k, v, None, True, 0.0, is_causal, False, scale=1.0)
            return r[0], [r[1], r[2], r[3], r[6], r[7]], ("cudnn", r[4], r[5])
        except (RuntimeError, TypeError):
            SDPA_IMPL = "flash"               # this build / shape has no cuDNN attention: ATen's flash kernels
    r = aten._scaled_dot_product_flash_attention(q, k, v, 0.0, is_causal, False, scale=1.0)
    return r[0], [r[1], r[2], r[3], r[6], r[7]], ("flash", r[4], r[5])


def _sdpa_bwd(do, q, k, v, o, lse, cq, ck, seed, off, meta, is_causal: bool):
    aten = torch.ops.aten
    kind, mq, mk = meta
    if kind == "cudnn":
        return aten._scaled_dot_product_cudnn_attention_backward(do, q, k, v, o, lse, seed, off, None, cq, ck, mq, mk, 0.0,
                                                                 is_causal, scale=1.0)
    return aten._scaled_dot_product_flash_attention_backward(do, q, k, v, o, lse, cq, ck, mq, mk, 0.0, is_causal, seed, off,
                                                             scale=1.0)


def _attn_fwd(proj_q, proj_kv, out_pack, x_q, x_kv, idx, causal: bool, residual: torch.Tensor, H: int):
    """One attention block of a layer: fused projections (row-major outputs, viewed as heads without a copy), SDPA,
    out_proj + residual.  proj_q produces q (and k, v when proj_kv is None: self-attention); proj_kv the cross-attention
    k | v from ``x_kv``.  Returns (h_out, saved tensors in _ATTN_FIELDS order, meta)."""
    u_q = _lora_u(proj_q, x_q, idx)
    g_q = _lora_u_dropout(proj_q, x_q, idx, u_q)
    ys = proj_q(x_q, idx if proj_q.lora_mods else None, u=u_q, head_major=False)
    u_kv = g_kv = None
    if proj_kv is None:
        q, k, v = ys
    else:
        (q,) = ys
        u_kv = _lora_u(proj_kv, x_kv, idx)
        g_kv = _lora_u_dropout(proj_kv, x_kv, idx, u_kv)
        k, v = proj_kv(x_kv, idx if proj_kv.lora_mods else None, u=u_kv, head_major=False)
    is_causal = causal and q.shape[1] > 1
    o, extra, meta = _sdpa_fwd(_heads(q, H), _heads(k, H), _heads(v, H), is_causal)
    _trace("attn.fwd u_q,u_kv,q,k,v,o", u_q, u_kv, q, k, v, o)
    p = out_pack.get()
    out = ops.linear_fwd(_to_rows(o), p.W, p.b, residual, SAR_ACT_NONE)
    return out, [x_q, x_kv, u_q, u_kv, q, k, v, o] + extra + [g_q, g_kv], (meta, is_causal, H)


def _proj_bwd(proj, x, u, idx, dys: List[torch.Tensor], gs: Optional[torch.Tensor] = None):
    """dX and the LoRA gradients of one fused projection call.  ``dys``: row-major output gradients per segment; ``gs``:
    the lora_dropout masks of the forward (``_lora_u_dropout``) or None."""
    dx = None
    grads: List[Optional[torch.Tensor]] = []
    set_i = 0
    for m, s, dy in zip(proj.mods, proj.seg_scale, dys):
        if s != 1.0:
            dy = dy * s                          # q = s·(x Wᵀ + b + Δ): the scale sits outside the projection
        lora = isinstance(m, RoutedLoRALinear) and bool(m.adapter_order)
        if lora and idx is not None:
            B, T, _ = x.shape
            part, g = _lora_bwd(m, dy, x, u[set_i].reshape(B * T, -1), idx)
            if gs is not None and _active_dropout(m) is not None:
                part, g = _lora_dropout_bwd(m, dy, x, gs[set_i], part, g, idx)
            grads += g
            set_i += 1
        else:
            if lora:                             # adapters present but switched off for this call: no gradient
                grads += [None] * (2 * len(m.adapter_order))
            base = m.base_layer if isinstance(m, RoutedLoRALinear) else m
            packs = proj.__dict__.setdefault("_dx_packs", {})
            pack = packs.get(id(m))
            if pack is None:
                pack = packs[id(m)] = _WtPack(base)
            part = _dense_dx(dy.contiguous(), pack)
        dx = part if dx is None else dx + part
    return dx, grads


def _attn_bwd(proj_q, proj_kv, out_pack, idx, saved, meta, dh_out: torch.Tensor):
    """dh_out: gradient of out_proj's output (the residual branch is the caller's).  Returns (dx_q, dx_kv, LoRA grads of
    proj_q, LoRA grads of proj_kv)."""
    x_q, x_kv, u_q, u_kv, q, k, v, o, lse, cq, ck, seed, off, g_q, g_kv = saved
    sd_meta, is_causal, H = meta
    do = _to_heads(_dense_dx(dh_out, out_pack), H)
    dq, dk, dv = _sdpa_bwd(do, _heads(q, H), _heads(k, H), _heads(v, H), o, lse, cq, ck, seed, off, sd_meta, is_causal)
    _trace("attn.bwd do,dq,dk,dv", do, dq, dk, dv)
    dq, dk, dv = _to_rows(dq), _to_rows(dk), _to_rows(dv)
    if proj_kv is None:
        dx, g = _proj_bwd(proj_q, x_q, u_q, idx, [dq, dk, dv], g_q)
        return dx, None, g, []
    dxq, gq = _proj_bwd(proj_q, x_q, u_q, idx, [dq], g_q)
    dxkv, gkv = _proj_bwd(proj_kv, x_kv, u_kv, idx, [dk, dv], g_kv)
    return dxq, dxkv, gq, gkv


class _WtPack:
    """Transposed bf16 weight of a frozen nn.Linear for the dX GEMM (same (pointer, version, epoch) keying as the packs)."""

    def __init__(self, m):
        self.m = m
        self.key = None
        self._wt_key = None

    def get(self):
        from .whisper_blocks import _pver

        key = _pver(self.m.weight)
        if key != self.key:
            self.key = key
            self.W = self.m.weight.detach().to(torch.bfloat16).contiguous()
        return self


def _ffn_fwd(layer, pk, h: torch.Tensor):
    """LayerNorm -> fc1 -> GELU -> fc2 + residual.  Saved: (h, mean, rstd, pre-activation)."""
    x3, mean3, rstd3 = _ln_fwd(h, layer.final_layer_norm)
    p1, p2 = pk["fc1"].get(), pk["fc2"].get()
    B, T, d = h.shape
    pre = ops.linear_fwd(x3.reshape(1, B * T, d), p1.W, p1.b, None, SAR_ACT_NONE)          # [1, M, ffn] pre-activation
    f = F.gelu(pre)
    out = ops.linear_fwd(f, p2.W, p2.b, h.reshape(1, B * T, d), SAR_ACT_NONE).view(B, T, d)
    return out, (h, mean3, rstd3, pre)


def _ffn_bwd(layer, pk, saved, dout: torch.Tensor) -> torch.Tensor:
    h, mean3, rstd3, pre = saved
    B, T, d = h.shape
    if FUSED_GELU_BWD:      # dpre = (dout·W2) * GELU'(pre) in the dX GEMM's epilogue (pre rides in as the residual operand)
        dpre = ops.linear_fwd(dout.reshape(1, B * T, d), _wt(pk["fc2"]), None, residual=pre, act=SAR_ACT_GELU_BWD)
    else:
        df = ops.linear_fwd(dout.reshape(1, B * T, d), _wt(pk["fc2"]), None)               # [1, M, ffn]
        dpre = torch.ops.aten.gelu_backward(df, pre)
    dx3 = ops.linear_fwd(dpre, _wt(pk["fc1"]), None).view(B, T, d)
    return dout + _ln_bwd(dx3, h, mean3, rstd3, layer.final_layer_norm)


# ------------------------------------------------------------------------------------------------ layer functions
def _weight_grads(ctx, grads: List[Optional[torch.Tensor]]):
    """Gradients for the Function's trailing inputs.  Normal mode: one per LoRA weight.  Anchored mode (every weight's
    gradient goes straight into the flat bucket, so the weights were not passed as inputs): nothing may be left over."""
    if ctx.n_ws == len(grads):
        return grads
    if any(g is not None for g in grads):
        raise RuntimeError("a LoRA gradient left the flat bucket between forward and backward (optimizer.zero_grad("
                           "set_to_none=True) after bucket.zero_()?)")
    return [None] * ctx.n_ws


def _fn_inputs(ws: List[torch.Tensor], mods: List[RoutedLoRALinear], h: torch.Tensor, other_requires_grad: bool):
    """Trailing inputs of a layer Function.  When K3 accumulates every gradient of the layer in place (``_all_direct``)
    the weights are NOT handed to autograd: their AccumulateGrad nodes would never receive anything, but they carry the
    stream they were created on, and the engine joins the backward stream with it — which breaks CUDA-graph capture of the
    step whenever an earlier (eager) autograd graph is still alive.  A fresh zero-size leaf keeps the Function
    differentiable when no activation input requires grad (first encoder layer)."""
    if not _all_direct(mods, h.device):
        return ws
    if h.requires_grad or other_requires_grad:
        return []
    return [torch.zeros((), dtype=torch.float32, device=h.device, requires_grad=True)]


def _no_autocast():
    """The reference trainer runs its forward under ``torch.autocast(bf16)`` (src/training/trainer.py:328-331).  Everything
    inside the layer Functions is already bf16 with explicit fp32 statistics; autocast would only re-route the few ATen
    calls (LayerNorm fallback -> fp32 outputs) — so it is switched off for their bodies."""
    return torch.autocast(device_type="cuda", enabled=False)


class _EncoderLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, layer, idx, *lora_ws):
        with _no_autocast():
            return _EncoderLayerFn._forward(ctx, h, layer, idx, *lora_ws)

    @staticmethod
    def _forward(ctx, h, layer, idx, *lora_ws):
        pk = layer._sar_pack
        h = h.contiguous()
        H = layer.self_attn.num_heads
        x1, mean1, rstd1 = _ln_fwd(h, layer.self_attn_layer_norm)
        h2, sa, sa_meta = _attn_fwd(pk["self"].qkv.get(), None, pk["self"].out, x1, None, idx, False, h, H)
        out, ffn_saved = _ffn_fwd(layer, pk, h2)
        ctx.layer, ctx.idx, ctx.sa_meta, ctx.n_ws = layer, idx, sa_meta, len(lora_ws)
        ctx.save_for_backward(h, mean1, rstd1, *sa, *ffn_saved)
        return out

    @staticmethod
    def backward(ctx, dout):
        layer = ctx.layer
        pk = layer._sar_pack
        t = ctx.saved_tensors
        h, mean1, rstd1 = t[:3]
        sa, ffn_saved = t[3:3 + _N_ATTN], t[3 + _N_ATTN:]
        dh2 = _ffn_bwd(layer, pk, ffn_saved, dout.contiguous())
        dx1, _, g, _ = _attn_bwd(pk["self"].qkv.get(), None, pk["self"].out, ctx.idx, sa, ctx.sa_meta, dh2)
        dh = dh2 + _ln_bwd(dx1, h, mean1, rstd1, layer.self_attn_layer_norm)
        return (dh, None, None, *_weight_grads(ctx, g))


class _DecoderLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, enc, layer, idx, *lora_ws):
        with _no_autocast():
            return _DecoderLayerFn._forward(ctx, h, enc, layer, idx, *lora_ws)

    @staticmethod
    def _forward(ctx, h, enc, layer, idx, *lora_ws):
        pk = layer._sar_pack
        h = h.contiguous()
        enc = enc.contiguous()
        H = layer.self_attn.num_heads
        x1, mean1, rstd1 = _ln_fwd(h, layer.self_attn_layer_norm)
        h2, sa, sa_meta = _attn_fwd(pk["self"].qkv.get(), None, pk["self"].out, x1, None, idx, True, h, H)
        x2, mean2, rstd2 = _ln_fwd(h2, layer.encoder_attn_layer_norm)
        h3, ca, ca_meta = _attn_fwd(pk["cross"].q.get(), pk["cross"].kv.get(), pk["cross"].out, x2, enc, idx, False, h2, H)
        out, ffn_saved = _ffn_fwd(layer, pk, h3)
        _trace("dec.fwd h,enc,h2,h3,out", h, enc, h2, h3, out)
        ctx.layer, ctx.idx, ctx.sa_meta, ctx.ca_meta, ctx.n_ws = layer, idx, sa_meta, ca_meta, len(lora_ws)
        ctx.save_for_backward(h, mean1, rstd1, h2, mean2, rstd2, *sa, *ca, *ffn_saved)
        return out

    @staticmethod
    def backward(ctx, dout):
        layer = ctx.layer
        pk = layer._sar_pack
        t = ctx.saved_tensors
        h, mean1, rstd1, h2, mean2, rstd2 = t[:6]
        sa, ca, ffn_saved = t[6:6 + _N_ATTN], t[6 + _N_ATTN:6 + 2 * _N_ATTN], t[6 + 2 * _N_ATTN:]
        dh3 = _ffn_bwd(layer, pk, ffn_saved, dout.contiguous())
        dx2, denc, gq, gkv = _attn_bwd(pk["cross"].q.get(), pk["cross"].kv.get(), pk["cross"].out, ctx.idx, ca, ctx.ca_meta,
                                       dh3)
        _trace("dec.bwd dout,dh3,dx2,denc", dout, dh3, dx2, denc)
        dh2 = dh3 + _ln_bwd(dx2, h2, mean2, rstd2, layer.encoder_attn_layer_norm)
        dx1, _, gs, _ = _attn_bwd(pk["self"].qkv.get(), None, pk["self"].out, ctx.idx, sa, ctx.sa_meta, dh2)
        dh = dh2 + _ln_bwd(dx1, h, mean1, rstd1, layer.self_attn_layer_norm)
        _trace("dec.bwd dh2,dx1,dh", dh2, dx1, dh)
        return (dh, denc, None, None, *_weight_grads(ctx, [*gs, *gq, *gkv]))


# ------------------------------------------------------------------------------------------------ LM head
class _LMHeadFn(torch.autograd.Function):
    """logits = x·Eᵀ with the frozen (tied) embedding E [V, d] ($HF/modeling_whisper.py:1135 proj_out): forward and dX on the
    pair kernel's dense entry.  V = 51865 is odd: logits and their gradient live in buffers whose row stride is padded to 8
    elements (TMA needs 16-byte aligned rows); the caller slices the pad columns off, so their gradient arrives as zeros."""

    @staticmethod
    def forward(ctx, x2, head, ldy):
        W = head.weight.detach()
        M, d = x2.shape
        buf = torch.empty(M, ldy, dtype=torch.bfloat16, device=x2.device)
        ops.dense_fwd(x2, d, 0, W, None, buf, ldy, 0, 1, M, d, W.shape[0])   # pad columns: never read (sliced off)
        ctx.head, ctx.d = head, d
        return buf

    @staticmethod
    def backward(ctx, dbuf):
        head = ctx.head
        M, ldy = dbuf.shape
        Wt = _lm_head_wt(head, ldy)                                   # [d, ldy] = Eᵀ, zero pad columns
        dx = torch.empty(M, ctx.d, dtype=torch.bfloat16, device=dbuf.device)
        ops.dense_fwd(dbuf.contiguous(), ldy, 0, Wt, None, dx, ctx.d, 0, 1, M, ldy, ctx.d)
        return dx, None, None


def _lm_head_wt(head, ldy: int) -> torch.Tensor:
    from .whisper_blocks import _pver

    key = (_pver(head.weight), ldy)
    c = head.__dict__.get("_sar_wt")
    if c is None or c[0] != key:
        W = head.weight.detach()
        Wt = torch.zeros(W.shape[1], ldy, dtype=torch.bfloat16, device=W.device)
        Wt[:, : W.shape[0]] = W.t()
        c = (key, Wt)
        head.__dict__["_sar_wt"] = c
    return c[1]


def lm_head_train(x2: torch.Tensor, head, ldy: int) -> torch.Tensor:
    CALLS["lm_head"] = CALLS.get("lm_head", 0) + 1
    return _LMHeadFn.apply(x2, head, ldy)


# ------------------------------------------------------------------------------------------------ entry points
REFUSED = {"encoder": "", "decoder": ""}     # why the last layer call kept HF's body ("" = it did not); for tests / debugging


def _train_refusal(layer, h: torch.Tensor, kwargs) -> str:
    from .whisper_blocks import FUSED_BLOCKS_ENABLED, _layer_dropout_active

    if not (ENABLED and FUSED_BLOCKS_ENABLED):
        return "switched off"
    if current_mix_weights() is not None:
        return "soft_fused mix weights active (inference only)"
    if not (h.is_cuda and h.dtype == torch.bfloat16 and h.dim() == 3 and h.shape[-1] % 128 == 0):
        return f"hidden states {tuple(h.shape)} {h.dtype} {h.device.type}"
    if kwargs.get("output_attentions", False):
        return "output_attentions"
    if _layer_dropout_active(layer):          # Whisper's own dropouts; lora_dropout is handled (_lora_u_dropout)
        return "dropout active"
    return ""


def _packs_refusal(*projs) -> str:
    for p in projs:
        if not p.ok:
            return "projection pack unsupported"
        for m in p.lora_mods:
            r = next(iter(m.r.values()))
            if m.training and m._dropout_active() and (len(m.adapter_order) != 1 or r % 16):
                return "lora_dropout with several adapters / a padded rank"   # the module slot's correction term applies
            if r % 16 or r > 64:
                return f"rank {r}"
    return ""


def encoder_layer_train(layer, hidden_states: torch.Tensor, kwargs) -> Optional[torch.Tensor]:
    """Fused forward + backward of a WhisperEncoderLayer under autograd, or None when HF's body has to run."""
    why = _train_refusal(layer, hidden_states, kwargs)
    if not why:
        qkv = layer._sar_pack["self"].qkv.get()
        why = _packs_refusal(qkv)
    if not why:
        ws = _lora_params(qkv.lora_mods)
        if not hidden_states.requires_grad and not any(w.requires_grad for w in ws):
            why = "nothing requires grad"
    REFUSED["encoder"] = why
    if why:
        return None
    idx = qkv.resolve_index(hidden_states.shape[0], hidden_states.device)
    CALLS["encoder_layers"] += 1
    return _EncoderLayerFn.apply(hidden_states, layer, idx, *_fn_inputs(ws, qkv.lora_mods, hidden_states, False))


_ENC_CAST: List = [None, None]        # (weakref to the fp32 encoder output, its bf16 cast): one cast per step, not per layer


def autocast_to_bf16(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Under ``torch.autocast(bf16)`` (the reference trainer's forward, src/training/trainer.py:328-331) HF's LayerNorms
    return fp32, and the next projection would cast that input to bf16 anyway: do that cast here, once, so the fused
    layers see what their kernels take.  Anything else is returned unchanged."""
    if (t is None or t.dtype != torch.float32 or not t.is_cuda or not torch.is_autocast_enabled("cuda")
            or torch.get_autocast_dtype("cuda") != torch.bfloat16):
        return t
    import weakref

    src = _ENC_CAST[0]() if _ENC_CAST[0] is not None else None
    if src is not t:
        _ENC_CAST[0], _ENC_CAST[1] = weakref.ref(t), t.to(torch.bfloat16)
    return _ENC_CAST[1]


def decoder_layer_train(layer, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor],
                        kwargs) -> Optional[torch.Tensor]:
    e = autocast_to_bf16(encoder_hidden_states)
    if e is None:
        why = "no encoder states"
    elif not (e.is_cuda and e.dtype == torch.bfloat16):
        why = f"encoder states {e.dtype} {e.device.type}"
    else:
        why = _train_refusal(layer, hidden_states, kwargs)
    if not why:
        pk = layer._sar_pack
        qkv, cq, ckv = pk["self"].qkv.get(), pk["cross"].q.get(), pk["cross"].kv.get()
        why = _packs_refusal(qkv, cq, ckv)
    if not why:
        ws = _lora_params(qkv.lora_mods) + _lora_params(cq.lora_mods) + _lora_params(ckv.lora_mods)
        if not (hidden_states.requires_grad or e.requires_grad or any(w.requires_grad for w in ws)):
            why = "nothing requires grad"
    REFUSED["decoder"] = why
    if why:
        return None
    idx = qkv.resolve_index(hidden_states.shape[0], hidden_states.device)
    CALLS["decoder_layers"] += 1
    mods = qkv.lora_mods + cq.lora_mods + ckv.lora_mods
    return _DecoderLayerFn.apply(hidden_states, e, layer, idx, *_fn_inputs(ws, mods, hidden_states, e.requires_grad))
