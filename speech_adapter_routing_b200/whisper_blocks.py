"""Fused Whisper encoder / decoder blocks on libsar.

HF's ``WhisperEncoderLayer`` / ``WhisperDecoderLayer`` stay where they are (same parameters, same state-dict keys,
same attribute paths the reference walks); only their ``forward`` is re-bound so that, for inference on a B200, one
layer is 7 (encoder) / 12 (decoder) launches instead of ~25 / ~45 eager ones:

    LayerNorm                         sar_layernorm_fwd
    q‖k‖v (+ routed LoRA on q, v)     sar_attn_proj_fwd   — x read once, query scale and the [B,T,d]→[B,h,T,64]
                                                            transpose folded into the epilogue's TMA store
    softmax(q kᵀ) v                   torch SDPA (library: cuDNN / flash kernels), on the head-major tensors
    out_proj + residual               sar_linear_fwd      — reads SDPA's [B,h,T,64] output in place (K block = head)
    LayerNorm                         sar_layernorm_fwd
    fc1 + GELU                        sar_linear_fwd
    fc2 + residual                    sar_linear_fwd

replacing $HF/models/whisper/modeling_whisper.py:284-357 (WhisperAttention.forward), :376-414 and :452-506 (layer
bodies).  The fused path is taken when its preconditions hold (CUDA bf16, no autograd, no KV cache, no
``output_attentions``, head_dim 64, GELU); otherwise the layer's original HF forward runs — whose q_proj / v_proj are
still RoutedLoRALinear (libsar K1/K3), so neither branch is a CPU or eager-LoRA fallback.
"""
from __future__ import annotations

import types
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import SAR_ACT_GELU, SAR_ACT_NONE
from .lora_linear import RoutedLoRALinear

FUSED_BLOCKS_ENABLED = True   # debug switch: False restores HF's layer bodies everywhere


# ------------------------------------------------------------------------------------------------ operand packing
def _pver(*ts) -> Tuple:
    return tuple((t.data_ptr(), t._version) if t is not None else None for t in ts)


def _lin_params(m: nn.Module) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    base = m.base_layer if isinstance(m, RoutedLoRALinear) else m
    return base.weight, base.bias


def _lora_key(m: nn.Module) -> Tuple:
    return m._key() if isinstance(m, RoutedLoRALinear) else ()


class _ProjPack:
    """Concatenated operands of one sar_attn_proj_fwd call over projections ``mods`` (in output order)."""

    def __init__(self, mods: List[nn.Module], seg_scale: List[float]):
        self.mods = mods
        self.seg_scale = seg_scale
        self.key = None

    @torch.no_grad()
    def get(self):
        key = tuple(_pver(*_lin_params(m)) + _lora_key(m) for m in self.mods)
        if key == self.key:
            return self
        self.key = key
        Ws, bs = [], []
        for m in self.mods:
            W, b = _lin_params(m)
            Ws.append(W.detach().to(torch.bfloat16))
            bs.append(torch.zeros(W.shape[0], dtype=torch.bfloat16, device=W.device) if b is None
                      else b.detach().to(torch.bfloat16))
        self.W = torch.cat(Ws, 0).contiguous()
        self.bias = torch.cat(bs, 0).contiguous()
        self.seg_set: List[int] = []
        As, Bps, scales = [], [], []
        self.lora_mods: List[RoutedLoRALinear] = []
        for m in self.mods:
            if isinstance(m, RoutedLoRALinear) and m.adapter_order:
                st = m._stacks()
                self.seg_set.append(len(As))
                As.append(st["A"])
                Bps.append(st["Bp"])
                scales.append(st["scale"])
                self.lora_mods.append(m)
            else:
                self.seg_set.append(-1)
        self.n_sets = max(len(As), 1)
        self.ok = True
        if As:
            shapes = {tuple(a.shape) for a in As}
            orders = {tuple(m.adapter_order) for m in self.lora_mods}
            if len(shapes) != 1 or len(orders) != 1 or any(abs(s - scales[0]) > 1e-12 for s in scales):
                self.ok = False   # q and v disagree on rank / adapter set / scaling: the layer keeps HF's body
            else:
                self.A = torch.cat(As, 0).contiguous()
                self.Bp = torch.cat(Bps, 0).contiguous()
                self.scale = float(scales[0])
        else:
            self.A = self.Bp = None
            self.scale = 0.0
        return self

    def resolve_index(self, B: int, device) -> Optional[torch.Tensor]:
        return self.lora_mods[0].resolve_index(B, device) if self.lora_mods else None

    def __call__(self, x: torch.Tensor, idx: Optional[torch.Tensor]) -> List[torch.Tensor]:
        lora = idx is not None and self.A is not None
        return ops.attn_proj_fwd(x, self.W, self.bias, self.A if lora else None, self.Bp if lora else None,
                                 idx if lora else None, self.seg_set if lora else [-1] * len(self.seg_set),
                                 self.seg_scale, self.n_sets if lora else 1, self.scale, y_head_major=True)


class _DensePack:
    """bf16 weight / bias of a plain nn.Linear or nn.LayerNorm, refreshed when the parameter changes."""

    def __init__(self, m: nn.Module):
        self.m = m
        self.key = None

    @torch.no_grad()
    def get(self):
        key = _pver(self.m.weight, self.m.bias)
        if key != self.key:
            self.key = key
            self.W = self.m.weight.detach().to(torch.bfloat16).contiguous()
            b = self.m.bias
            self.b = None if b is None else b.detach().to(torch.bfloat16).contiguous()
        return self


class _AttnPack:
    def __init__(self, attn: nn.Module, cross: bool):
        s = float(attn.scaling)
        if cross:
            self.q = _ProjPack([attn.q_proj], [s])
            self.kv = _ProjPack([attn.k_proj, attn.v_proj], [1.0, 1.0])
        else:
            self.qkv = _ProjPack([attn.q_proj, attn.k_proj, attn.v_proj], [s, 1.0, 1.0])
        self.out = _DensePack(attn.out_proj)


def _supported_attn(attn: nn.Module) -> bool:
    if attn.head_dim != 64 or attn.embed_dim % 128:
        return False
    for name in ("q_proj", "k_proj", "v_proj"):
        m = getattr(attn, name)
        if not isinstance(m, (nn.Linear, RoutedLoRALinear)):
            return False
    return isinstance(attn.out_proj, nn.Linear)


def _is_gelu(layer: nn.Module) -> bool:
    fn = layer.activation_fn
    return isinstance(fn, nn.GELU) and getattr(fn, "approximate", "none") == "none" or \
        type(fn).__name__ == "GELUActivation" and not getattr(fn, "use_gelu_python", False)


def _fast_path_ok(layer: nn.Module, h: torch.Tensor, kwargs) -> bool:
    return (FUSED_BLOCKS_ENABLED and h.is_cuda and h.dtype == torch.bfloat16 and h.dim() == 3
            and not torch.is_grad_enabled() and not kwargs.get("output_attentions", False))


def _ln(x: torch.Tensor, pack: _DensePack, eps: float) -> torch.Tensor:
    p = pack.get()
    return ops.layernorm_fwd(x, p.W, p.b, eps)


def _dense(x: torch.Tensor, pack: _DensePack, residual: Optional[torch.Tensor] = None, act: int = SAR_ACT_NONE,
           head_major: bool = False, inplace: bool = False) -> torch.Tensor:
    """act(x·Wᵀ + b) + residual.  Row-major inputs are flattened to one [1, B·T, d] "utterance" (no LoRA term, so
    tiles may span utterances: no padding rows when T is not a multiple of the 256-row pair tile)."""
    p = pack.get()
    if head_major:
        return ops.linear_fwd(x, p.W, p.b, residual, act, x_head_major=True, out=residual if inplace else None)
    B, T, d = x.shape
    r2 = None if residual is None else residual.view(1, B * T, -1)
    y = ops.linear_fwd(x.view(1, B * T, d), p.W, p.b, r2, act, out=r2 if inplace else None)
    return y.view(B, T, -1)


def _sdpa(q, k, v, mask=None, causal=False):
    # q is pre-scaled inside the projection epilogue (HF applies `* self.scaling` to q_proj's output, then scale=1)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, is_causal=causal and mask is None and q.shape[2] > 1,
                                          scale=1.0)


# ------------------------------------------------------------------------------------------------ layer bodies
def _encoder_layer_forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, **kwargs):
    pk = self._sar_pack
    if attention_mask is not None or not _fast_path_ok(self, hidden_states, kwargs):
        return self._sar_hf_forward(hidden_states, attention_mask, **kwargs)
    qkv = pk["self"].qkv.get()
    if not qkv.ok:
        return self._sar_hf_forward(hidden_states, attention_mask, **kwargs)
    h = hidden_states.contiguous()
    B = h.shape[0]
    idx = qkv.resolve_index(B, h.device)
    x = _ln(h, pk["ln1"], self.self_attn_layer_norm.eps)
    q, k, v = qkv(x, idx)
    o = _sdpa(q, k, v)
    h = _dense(o, pk["self"].out, residual=h, head_major=True)
    x = _ln(h, pk["ln3"], self.final_layer_norm.eps)
    f = _dense(x, pk["fc1"], act=SAR_ACT_GELU)
    return _dense(f, pk["fc2"], residual=h, inplace=True)


def _decoder_layer_forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                           encoder_hidden_states: Optional[torch.Tensor] = None,
                           encoder_attention_mask: Optional[torch.Tensor] = None, past_key_values=None,
                           use_cache: Optional[bool] = True, **kwargs):
    pk = self._sar_pack
    fast = (past_key_values is None and encoder_attention_mask is None and _fast_path_ok(self, hidden_states, kwargs)
            and (encoder_hidden_states is None or
                 (encoder_hidden_states.is_cuda and encoder_hidden_states.dtype == torch.bfloat16)))
    if fast:
        qkv = pk["self"].qkv.get()
        fast = qkv.ok
        if fast and encoder_hidden_states is not None:
            cq, ckv = pk["cross"].q.get(), pk["cross"].kv.get()
            fast = cq.ok and ckv.ok
    if not fast:
        return self._sar_hf_forward(hidden_states, attention_mask, encoder_hidden_states,
                                    encoder_attention_mask=encoder_attention_mask, past_key_values=past_key_values,
                                    use_cache=use_cache, **kwargs)
    h = hidden_states.contiguous()
    B = h.shape[0]
    idx = qkv.resolve_index(B, h.device)
    x = _ln(h, pk["ln1"], self.self_attn_layer_norm.eps)
    q, k, v = qkv(x, idx)
    o = _sdpa(q, k, v, mask=attention_mask, causal=True)
    h = _dense(o, pk["self"].out, residual=h, head_major=True)
    if encoder_hidden_states is not None:
        x = _ln(h, pk["ln2"], self.encoder_attn_layer_norm.eps)
        (q,) = cq(x, cq.resolve_index(B, h.device))
        k, v = ckv(encoder_hidden_states.contiguous(), ckv.resolve_index(B, h.device))
        o = _sdpa(q, k, v)
        h = _dense(o, pk["cross"].out, residual=h, head_major=True, inplace=True)
    x = _ln(h, pk["ln3"], self.final_layer_norm.eps)
    f = _dense(x, pk["fc1"], act=SAR_ACT_GELU)
    return _dense(f, pk["fc2"], residual=h, inplace=True)


# ------------------------------------------------------------------------------------------------ installation
def _layer_supported(layer: nn.Module, decoder: bool) -> bool:
    if not _is_gelu(layer) or not _supported_attn(layer.self_attn):
        return False
    if decoder and not _supported_attn(layer.encoder_attn):
        return False
    if layer.fc1.out_features % 128 or layer.fc1.in_features % 128:
        return False
    lns = [layer.self_attn_layer_norm, layer.final_layer_norm] + ([layer.encoder_attn_layer_norm] if decoder else [])
    return all(isinstance(m, nn.LayerNorm) and m.elementwise_affine and m.bias is not None for m in lns)


def install_fused_blocks(model: nn.Module) -> int:
    """Re-bind ``forward`` of every supported WhisperEncoderLayer / WhisperDecoderLayer under ``model`` (idempotent).
    Call again after module surgery (e.g. after LoRA injection replaced q_proj / v_proj).  Returns #layers bound."""
    from transformers.models.whisper.modeling_whisper import WhisperDecoderLayer, WhisperEncoderLayer

    n = 0
    for layer in model.modules():
        dec = isinstance(layer, WhisperDecoderLayer)
        if not dec and not isinstance(layer, WhisperEncoderLayer):
            continue
        if not _layer_supported(layer, dec):
            continue
        if not hasattr(layer, "_sar_hf_forward"):
            object.__setattr__(layer, "_sar_hf_forward", layer.forward)   # the bound class method (HF's body)
        pack: Dict[str, object] = {
            "self": _AttnPack(layer.self_attn, cross=False),
            "ln1": _DensePack(layer.self_attn_layer_norm),
            "ln3": _DensePack(layer.final_layer_norm),
            "fc1": _DensePack(layer.fc1),
            "fc2": _DensePack(layer.fc2),
        }
        if dec:
            pack["cross"] = _AttnPack(layer.encoder_attn, cross=True)
            pack["ln2"] = _DensePack(layer.encoder_attn_layer_norm)
        object.__setattr__(layer, "_sar_pack", pack)
        object.__setattr__(layer, "forward",
                           types.MethodType(_decoder_layer_forward if dec else _encoder_layer_forward, layer))
        n += 1
    return n


def uninstall_fused_blocks(model: nn.Module) -> None:
    for layer in model.modules():
        if hasattr(layer, "_sar_hf_forward"):
            object.__delattr__(layer, "forward")
            object.__delattr__(layer, "_sar_hf_forward")
            object.__delattr__(layer, "_sar_pack")
