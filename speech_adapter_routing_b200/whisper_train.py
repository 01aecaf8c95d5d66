"""Training path of the fused Whisper blocks: forward AND backward of a whole encoder / decoder layer on libsar.

The reference trains with HF's eager layer bodies under autograd (src/training/trainer.py:251-256, bf16 autocast,
gradient checkpointing on by default, src/models/whisper_lora.py:81-83): ~45 launches per layer forward, twice, plus the
autograd backward.  Only the LoRA tensors train; the base weights are frozen, so the backward needs exactly

    dX through every dense layer                      -> the tcgen05 pair kernel on the transposed weight (sar_linear_fwd)
    dX, dA, dB at q_proj / v_proj                     -> K3 (sar_qv_lora_bwd), accumulating into the flat gradient bucket
    LayerNorm, GELU, softmax(QKᵀ)V backward           -> ATen / SDPA kernels (library) on the tensors the forward saved

Each layer is ONE ``torch.autograd.Function``: its forward runs the same fused launches as inference (LayerNorm statistics
come from ATen's LayerNorm so that its backward can reuse them) and saves the handful of tensors the backward needs; under
HF's gradient checkpointing that forward is the recompute and the saved tensors live only until the layer's backward.
The HF modules, their parameters and their attribute paths are untouched; when a precondition fails (dropout active,
masks, KV cache, non-bf16, LoRA on other modules) the layer keeps HF's body over the K1 / K3 module slots.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops
from ._lib import SAR_ACT_NONE
from .lora_linear import RoutedLoRALinear, _notify_grad_ready

ENABLED = __import__("os").environ.get("SAR_FUSED_TRAIN", "1") != "0"
# softmax(QKᵀ)V forward / backward: "cudnn" and "flash" call ATen's fused-attention ops directly (no nested autograd
# graph, so the step is CUDA-graph capturable and costs one host call each way); "autograd" differentiates F.sdpa.
SDPA_IMPL = __import__("os").environ.get("SAR_TRAIN_SDPA", "cudnn")
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
    return torch.native_layer_norm(x, (x.shape[-1],), ln.weight, ln.bias, ln.eps)


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


class _Attn:
    """Forward of one attention (projections on the fused kernel, SDPA under a private autograd graph) that keeps what
    the backward needs."""

    def __init__(self, proj_q, proj_kv, out_pack, x_q, x_kv, idx, causal: bool):
        # proj_q: _ProjPack producing q (and k, v when proj_kv is None: self-attention); proj_kv: cross-attention k | v
        self.proj_q, self.proj_kv, self.out_pack, self.causal = proj_q, proj_kv, out_pack, causal
        self.x_q, self.x_kv, self.idx = x_q, x_kv, idx
        self.u_q = self._u(proj_q, x_q)
        ys = proj_q(x_q, idx if proj_q.lora_mods else None, u=self.u_q)
        if proj_kv is None:
            q, k, v = ys
            self.u_kv = None
        else:
            (q,) = ys
            self.u_kv = self._u(proj_kv, x_kv)
            k, v = proj_kv(x_kv, idx if proj_kv.lora_mods else None, u=self.u_kv)
        self.is_causal = causal and q.shape[2] > 1
        self._sdpa_fwd(q, k, v)
        _trace("attn.fwd u_q,u_kv,q,k,v,o", self.u_q, self.u_kv, q, k, v, self.o)

    def _sdpa_fwd(self, q, k, v) -> None:
        global SDPA_IMPL
        aten = torch.ops.aten
        if SDPA_IMPL == "cudnn":
            try:
                r = aten._scaled_dot_product_cudnn_attention(q, k, v, None, True, 0.0, self.is_causal, False, scale=1.0)
                self.q, self.k, self.v, self.o, self.sdpa = q, k, v, r[0], ("cudnn",) + tuple(r[1:8])
                return
            except (RuntimeError, TypeError):
                SDPA_IMPL = "flash"               # this build / shape has no cuDNN attention: ATen's flash kernels
        if SDPA_IMPL == "flash":
            r = aten._scaled_dot_product_flash_attention(q, k, v, 0.0, self.is_causal, False, scale=1.0)
            self.q, self.k, self.v, self.o, self.sdpa = q, k, v, r[0], ("flash",) + tuple(r[1:8])
            return
        with torch.enable_grad():
            self.q, self.k, self.v = (t.detach().requires_grad_(True) for t in (q, k, v))
            self.o = F.scaled_dot_product_attention(self.q, self.k, self.v, is_causal=self.is_causal, scale=1.0)
            self.sdpa = ("autograd",)

    def _sdpa_bwd(self, do: torch.Tensor):
        aten = torch.ops.aten
        kind = self.sdpa[0]
        if kind == "autograd":
            return torch.autograd.grad(self.o, (self.q, self.k, self.v), do)
        lse, cq, ck, mq, mk, seed, off = self.sdpa[1:]
        if kind == "cudnn":
            return aten._scaled_dot_product_cudnn_attention_backward(do, self.q, self.k, self.v, self.o, lse, seed, off,
                                                                     None, cq, ck, mq, mk, 0.0, self.is_causal, scale=1.0)
        return aten._scaled_dot_product_flash_attention_backward(do, self.q, self.k, self.v, self.o, lse, cq, ck, mq, mk,
                                                                 0.0, self.is_causal, seed, off, scale=1.0)

    def _u(self, proj, x):
        """U = scale·x·A_kᵀ planes for this call ([n_sets, B, T, r]) — the forward's low-rank operand and K3's ``u``."""
        if self.idx is None or proj.A is None:
            return None
        return ops.lora_u_fwd(x, proj.A, self.idx, proj.n_sets, proj.scale, proj.W.shape[0] // len(proj.mods))

    def out(self, residual: torch.Tensor) -> torch.Tensor:
        p = self.out_pack.get()
        return ops.linear_fwd(self.o.detach(), p.W, p.b, residual, SAR_ACT_NONE, x_head_major=True)

    def _proj_bwd(self, proj, x, u, dys_hm: List[torch.Tensor]):
        """dX and the LoRA gradients of one fused projection call.  ``dys_hm``: head-major output gradients per segment."""
        dx = None
        grads: List[Optional[torch.Tensor]] = []
        set_i = 0
        for m, s, dy_hm in zip(proj.mods, proj.seg_scale, dys_hm):
            dy = _to_rows(dy_hm)
            if s != 1.0:
                dy = dy * s                      # q = s·(x Wᵀ + b + Δ): the scale sits outside the projection
            lora = isinstance(m, RoutedLoRALinear) and bool(m.adapter_order)
            if lora and self.idx is not None:
                B, T, _ = x.shape
                part, g = _lora_bwd(m, dy, x, u[set_i].reshape(B * T, -1), self.idx)
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
                part = _dense_dx(dy, pack)
            dx = part if dx is None else dx + part
        return dx, grads

    def backward(self, dh_out: torch.Tensor):
        """dh_out: gradient of out_proj's output (the residual branch is handled by the caller).  Returns
        (dx_q, dx_kv, lora grads of proj_q, lora grads of proj_kv)."""
        H = self.q.shape[1]
        do = _to_heads(_dense_dx(dh_out, self.out_pack), H)
        dq, dk, dv = self._sdpa_bwd(do)
        _trace("attn.bwd do,dq,dk,dv", do, dq, dk, dv)
        if self.proj_kv is None:
            dx, g = self._proj_bwd(self.proj_q, self.x_q, self.u_q, [dq, dk, dv])
            return dx, None, g, []
        dxq, gq = self._proj_bwd(self.proj_q, self.x_q, self.u_q, [dq])
        dxkv, gkv = self._proj_bwd(self.proj_kv, self.x_kv, self.u_kv, [dk, dv])
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
    df = ops.linear_fwd(dout.reshape(1, B * T, d), _wt(pk["fc2"]), None)                   # [1, M, ffn]
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


class _EncoderLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, layer, idx, *lora_ws):
        pk = layer._sar_pack
        h = h.contiguous()
        x1, mean1, rstd1 = _ln_fwd(h, layer.self_attn_layer_norm)
        attn = _Attn(pk["self"].qkv.get(), None, pk["self"].out, x1, None, idx, causal=False)
        h2 = attn.out(h)
        out, ffn_saved = _ffn_fwd(layer, pk, h2)
        ctx.layer, ctx.attn, ctx.ffn_saved, ctx.n_ws = layer, attn, ffn_saved, len(lora_ws)
        ctx.save_for_backward(h, mean1, rstd1)
        return out

    @staticmethod
    def backward(ctx, dout):
        layer, attn = ctx.layer, ctx.attn
        pk = layer._sar_pack
        h, mean1, rstd1 = ctx.saved_tensors
        dh2 = _ffn_bwd(layer, pk, ctx.ffn_saved, dout.contiguous())
        dx1, _, g, _ = attn.backward(dh2)
        dh = dh2 + _ln_bwd(dx1, h, mean1, rstd1, layer.self_attn_layer_norm)
        ctx.attn = ctx.ffn_saved = None
        return (dh, None, None, *_weight_grads(ctx, g))


class _DecoderLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, enc, layer, idx, *lora_ws):
        pk = layer._sar_pack
        h = h.contiguous()
        enc = enc.contiguous()
        x1, mean1, rstd1 = _ln_fwd(h, layer.self_attn_layer_norm)
        sa = _Attn(pk["self"].qkv.get(), None, pk["self"].out, x1, None, idx, causal=True)
        h2 = sa.out(h)
        x2, mean2, rstd2 = _ln_fwd(h2, layer.encoder_attn_layer_norm)
        ca = _Attn(pk["cross"].q.get(), pk["cross"].kv.get(), pk["cross"].out, x2, enc, idx, causal=False)
        h3 = ca.out(h2)
        out, ffn_saved = _ffn_fwd(layer, pk, h3)
        _trace("dec.fwd h,enc,h2,h3,out", h, enc, h2, h3, out)
        ctx.layer, ctx.sa, ctx.ca, ctx.ffn_saved, ctx.n_ws = layer, sa, ca, ffn_saved, len(lora_ws)
        ctx.save_for_backward(h, mean1, rstd1, h2, mean2, rstd2)
        return out

    @staticmethod
    def backward(ctx, dout):
        layer, sa, ca = ctx.layer, ctx.sa, ctx.ca
        pk = layer._sar_pack
        h, mean1, rstd1, h2, mean2, rstd2 = ctx.saved_tensors
        dh3 = _ffn_bwd(layer, pk, ctx.ffn_saved, dout.contiguous())
        dx2, denc, gq, gkv = ca.backward(dh3)
        _trace("dec.bwd dout,dh3,dx2,denc", dout, dh3, dx2, denc)
        dh2 = dh3 + _ln_bwd(dx2, h2, mean2, rstd2, layer.encoder_attn_layer_norm)
        dx1, _, gs, _ = sa.backward(dh2)
        dh = dh2 + _ln_bwd(dx1, h, mean1, rstd1, layer.self_attn_layer_norm)
        _trace("dec.bwd dh2,dx1,dh", dh2, dx1, dh)
        ctx.sa = ctx.ca = ctx.ffn_saved = None
        return (dh, denc, None, None, *_weight_grads(ctx, [*gs, *gq, *gkv]))


# ------------------------------------------------------------------------------------------------ entry points
REFUSED = {"encoder": "", "decoder": ""}     # why the last layer call kept HF's body ("" = it did not); for tests / debugging


def _train_refusal(layer, h: torch.Tensor, kwargs) -> str:
    from .whisper_blocks import FUSED_BLOCKS_ENABLED, _dropout_active

    if not (ENABLED and FUSED_BLOCKS_ENABLED):
        return "switched off"
    if not (h.is_cuda and h.dtype == torch.bfloat16 and h.dim() == 3 and h.shape[-1] % 128 == 0):
        return f"hidden states {tuple(h.shape)} {h.dtype} {h.device.type}"
    if kwargs.get("output_attentions", False):
        return "output_attentions"
    if _dropout_active(layer):
        return "dropout active"
    return ""


def _packs_refusal(*projs) -> str:
    for p in projs:
        if not p.ok:
            return "projection pack unsupported"
        for m in p.lora_mods:
            if m.training and m._dropout_active():       # lora_dropout > 0: the module slot's correction term applies
                return "lora_dropout active"
            r = next(iter(m.r.values()))
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


def decoder_layer_train(layer, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor],
                        kwargs) -> Optional[torch.Tensor]:
    e = encoder_hidden_states
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
