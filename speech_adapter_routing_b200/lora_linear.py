"""RoutedLoRALinear — the module that occupies every ``q_proj`` / ``v_proj`` slot of HF Whisper.

It takes the place of PEFT's ``lora.Linear`` (which the reference installs through ``get_peft_model`` at
src/models/whisper_lora.py:88-98) and keeps PEFT's sub-module / parameter names so that state-dict keys and the
``adapter_model.safetensors`` layout are unchanged:

    base_layer.{weight,bias}     lora_A.<adapter>.weight [r, d_in]     lora_B.<adapter>.weight [d_out, r]

Unlike PEFT it holds *all* language adapters at once, stacked for the fused kernel, and applies adapter
``utt_adapter[b]`` to utterance ``b`` (routing context, see routing.py).  The arithmetic is libsar's K1 / K3
(sm_100a); there is no eager fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .routing import current_mix_weights, current_utt_adapter, operand_epoch, refresh_operands, routing_base_only


class _QVLoRAFn(torch.autograd.Function):
    """y = K1(x; W, A_k, B_k);  backward = K3 (dx, dA, dB) — base W frozen."""

    @staticmethod
    def forward(ctx, x, module, utt_adapter, *lora_weights):
        st = module._stacks()
        y, u = ops.qv_lora_fwd(x, st["W"], st["bias"], st["A"], st["Bp"], utt_adapter, st["scale"], save_u=True)
        ctx.module = module
        ctx.save_for_backward(x, u, utt_adapter)
        ctx.n_weights = len(lora_weights)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, u, utt_adapter = ctx.saved_tensors
        m: "RoutedLoRALinear" = ctx.module
        st = m._stacks(backward=True)
        n, rp = st["A"].shape[0], st["A"].shape[1]
        names = m.adapter_order
        # Single-adapter training (the reference trainer's case) under dist.FlatGradBucket: K3 accumulates STRAIGHT into
        # the parameters' ``.grad`` tensors, which are slices of the flat NCCL bucket — the gradient never exists anywhere
        # else ("accumulating into caller-provided offsets of one flat bucket").  Opt-in per parameter
        # (``_sar_direct_grad``, set by the bucket): plain ``torch.autograd.grad`` users keep autograd's semantics.
        direct = None
        if n == 1 and m.r[names[0]] == rp and st["grad_a_gain"][0] == 1.0 and st["grad_b_gain"][0] == 1.0:
            wA, wB = m.lora_A[names[0]].weight, m.lora_B[names[0]].weight
            gA, gB = wA.grad, wB.grad
            if (getattr(wA, "_sar_direct_grad", False) and getattr(wB, "_sar_direct_grad", False)
                    and wA.requires_grad and wB.requires_grad and gA is not None and gB is not None
                    and gA.dtype == torch.float32 and gB.dtype == torch.float32 and gA.is_contiguous()
                    and gB.is_contiguous() and gA.device == x.device and not gA.requires_grad and not gB.requires_grad):
                direct = (gA.view(1, rp, m.in_features), gB.view(1, m.out_features, rp))
        if direct is not None:
            dA, dB = direct
        else:
            dA = torch.zeros(n, rp, m.in_features, dtype=torch.float32, device=x.device)
            dB = torch.zeros(n, m.out_features, rp, dtype=torch.float32, device=x.device)
        dx = ops.qv_lora_bwd(dy.to(torch.bfloat16), x, u, st["Wt"], st["At"], st["Bt"], utt_adapter, dA, dB,
                             st["scale"], need_dx=ctx.needs_input_grad[0])
        if direct is not None:
            _notify_grad_ready(m.lora_A[names[0]].weight, m.lora_B[names[0]].weight)
            return (dx, None, None, None, None)          # the gradients are already in place
        grads: List[Optional[torch.Tensor]] = []
        for k, name in enumerate(names):        # lora_A weights, in stack order
            w = m.lora_A[name].weight
            grads.append((dA[k, : m.r[name]] * st["grad_a_gain"][k]).to(w.dtype) if w.requires_grad else None)
        for k, name in enumerate(names):        # lora_B weights
            w = m.lora_B[name].weight
            grads.append((dB[k, :, : m.r[name]] * st["grad_b_gain"][k]).to(w.dtype) if w.requires_grad else None)
        return (dx, None, None, *grads)


# Parameters whose gradient K3 wrote in place tell whoever registered here (dist.FlatGradBucket's overlapped all-reduce):
# autograd's AccumulateGrad — and with it every post-accumulate hook — never runs for them.
GRAD_READY_LISTENERS: List = []      # callables or weakref.WeakMethod objects (dead ones are dropped on the next notification)


def _notify_grad_ready(*params) -> None:
    import weakref

    dead = False
    for ref in GRAD_READY_LISTENERS:
        fn = ref() if isinstance(ref, weakref.WeakMethod) else ref
        if fn is None:
            dead = True
        else:
            fn(params)
    if dead:
        GRAD_READY_LISTENERS[:] = [r for r in GRAD_READY_LISTENERS if not (isinstance(r, weakref.WeakMethod) and r() is None)]


class RoutedLoRALinear(nn.Module):
    def __init__(self, base_layer: nn.Linear, adapter_name: str = "default", r: int = 16, lora_alpha: float = 32,
                 lora_dropout: float = 0.0):
        super().__init__()
        if not isinstance(base_layer, nn.Linear):
            raise TypeError("RoutedLoRALinear wraps nn.Linear")
        self.base_layer = base_layer
        self.in_features = base_layer.in_features
        self.out_features = base_layer.out_features
        self.r: Dict[str, int] = {}
        self.lora_alpha: Dict[str, float] = {}
        self.scaling: Dict[str, float] = {}
        self.lora_dropout = nn.ModuleDict()
        self.lora_A = nn.ModuleDict()
        self.lora_B = nn.ModuleDict()
        self.adapter_order: List[str] = []   # adapter name -> index k used by utt_adapter
        self.active_adapter: Optional[str] = None
        self.disable_adapters = False
        self._cache: Dict[str, object] = {}
        for p in base_layer.parameters():
            p.requires_grad = False
        if adapter_name is not None:
            self.add_adapter(adapter_name, r, lora_alpha, lora_dropout)

    # ------------------------------------------------------------------ adapter management
    def add_adapter(self, name: str, r: int, lora_alpha: float, lora_dropout: float = 0.0) -> None:
        if name in self.lora_A:
            raise ValueError(f"adapter {name!r} already exists")
        if r <= 0 or r > ops.SAR_RPAD:
            raise ValueError(f"LoRA rank must be in [1, {ops.SAR_RPAD}], got {r}")
        dev = self.base_layer.weight.device
        A = nn.Linear(self.in_features, r, bias=False, device=dev, dtype=torch.float32)
        B = nn.Linear(r, self.out_features, bias=False, device=dev, dtype=torch.float32)
        nn.init.kaiming_uniform_(A.weight, a=math.sqrt(5))   # PEFT default init for lora_A
        nn.init.zeros_(B.weight)                             # PEFT default init for lora_B
        self.lora_A[name] = A
        self.lora_B[name] = B
        self.lora_dropout[name] = nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()
        self.r[name] = r
        self.lora_alpha[name] = lora_alpha
        self.scaling[name] = lora_alpha / r
        self.adapter_order.append(name)
        if self.active_adapter is None:
            self.active_adapter = name
        self._cache.clear()
        refresh_operands()

    def set_adapter(self, name: str) -> None:
        if name not in self.lora_A:
            raise KeyError(name)
        self.active_adapter = name

    def adapter_index(self, name: str) -> int:
        return self.adapter_order.index(name)

    # ------------------------------------------------------------------ kernel operand stacks
    def _key(self):
        ws = [self.base_layer.weight, self.base_layer.bias] + [self.lora_A[n].weight for n in self.adapter_order] + \
             [self.lora_B[n].weight for n in self.adapter_order]
        return tuple((w.data_ptr(), w._version) if w is not None else None for w in ws) + (operand_epoch(),)

    @torch.no_grad()
    def _stacks(self, backward: bool = False) -> Dict[str, object]:
        """bf16 operand stacks for K1 (and K3 when ``backward``), rebuilt only when a parameter changed."""
        key = self._key()
        c = self._cache
        if c.get("key") != key:
            c.clear()
            c["key"] = key
            W = self.base_layer.weight
            dev = W.device
            c["W"] = W.detach().to(torch.bfloat16).contiguous()
            b = self.base_layer.bias
            c["bias"] = None if b is None else b.detach().to(torch.bfloat16).contiguous()
            names = self.adapter_order
            if names:
                rp = (max(self.r[n] for n in names) + 15) // 16 * 16
                scal = [self.scaling[n] for n in names]
                uniform = all(abs(s - scal[0]) < 1e-12 for s in scal)
                c["scale"] = float(scal[0]) if uniform else 1.0
                A = torch.zeros(len(names), rp, self.in_features, dtype=torch.float32, device=dev)
                Bm = torch.zeros(len(names), self.out_features, rp, dtype=torch.float32, device=dev)
                for k, n in enumerate(names):
                    A[k, : self.r[n]] = self.lora_A[n].weight.detach().float()
                    Bm[k, :, : self.r[n]] = self.lora_B[n].weight.detach().float() * (1.0 if uniform else scal[k])
                c["A"] = A.to(torch.bfloat16)
                c["Bp"] = ops.pack_lora_b(Bm)
                c["_Bm"] = Bm
                c["grad_a_gain"] = [1.0] * len(names)
                c["grad_b_gain"] = [1.0 if uniform else scal[k] for k in range(len(names))]
            else:
                c["A"] = c["Bp"] = None
                c["scale"] = 0.0
        if backward and "Wt" not in c:
            c["Wt"] = c["W"].t().contiguous()
            c["At"] = ops.pack_lora_b(c["A"].transpose(1, 2).contiguous())      # [n, d_in, 64]
            c["Bt"] = self._bm().to(torch.bfloat16).transpose(1, 2).contiguous()   # [n, r, d_out]
        return c

    @torch.no_grad()
    def _bm(self) -> torch.Tensor:
        """fp32 [n, d_out, rp] stack of scaling-folded lora_B (source of Bt and of the merged soft_fused operands).  An
        in-place operand refresh (operand_refresh.py) drops it; it is rebuilt from the parameters on demand."""
        c = self._cache
        if c.get("_Bm") is None:
            names = self.adapter_order
            rp = c["A"].shape[1]
            scal = [self.scaling[n] for n in names]
            uniform = all(abs(s - scal[0]) < 1e-12 for s in scal)
            Bm = torch.zeros(len(names), self.out_features, rp, dtype=torch.float32, device=c["A"].device)
            for k, n in enumerate(names):
                Bm[k, :, : self.r[n]] = self.lora_B[n].weight.detach().float() * (1.0 if uniform else scal[k])
            c["_Bm"] = Bm
        return c["_Bm"]

    def _default_index(self, B: int, device) -> Optional[torch.Tensor]:
        if self.disable_adapters or self.active_adapter is None or not self.adapter_order:
            return None
        k = self.adapter_index(self.active_adapter)
        ck = ("idx", B, k, str(device))
        t = self._cache.get(ck)
        if t is None:
            t = torch.full((B,), k, dtype=torch.int32, device=device)
            self._cache[ck] = t
        return t

    def resolve_index(self, B: int, device) -> Optional[torch.Tensor]:
        """int32 [B] adapter index per utterance for this call: the routing context if one is active, else the
        module's active adapter for every utterance; None = base weights only."""
        if routing_base_only():
            return None
        idx = current_utt_adapter()
        if idx is None:
            return self._default_index(B, device)
        if self.disable_adapters or not self.adapter_order:
            return None
        if idx.numel() != B:
            if B % idx.numel():
                raise ValueError(f"routing context has {idx.numel()} utterances but the batch has {B}")
            idx = idx.repeat_interleave(B // idx.numel())   # beam search expands the batch
        return idx

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("RoutedLoRALinear runs on libsar's sm_100a kernels only (no CPU fallback); "
                               "move the model and inputs to a B200")
        if current_mix_weights() is not None:
            raise NotImplementedError("soft_fused routing runs on the fused Whisper blocks only (bf16 CUDA inference with "
                                      "head_dim 64); this projection was reached through HF's layer body")
        in_dtype = x.dtype
        lead = x.shape[:-1]
        if x.dim() == 2:
            x3 = x.unsqueeze(0)
        elif x.dim() == 3:
            x3 = x
        else:
            x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
        if x3.dtype != torch.bfloat16:
            x3 = x3.to(torch.bfloat16)
        B, T = x3.shape[0], x3.shape[1]
        idx = self.resolve_index(B, x3.device)
        drop_fix = self.training and idx is not None and self._dropout_active()
        need_grad = torch.is_grad_enabled() and idx is not None and (
            x3.requires_grad or any(p.requires_grad for p in self.lora_A.parameters()) or
            any(p.requires_grad for p in self.lora_B.parameters()))
        if need_grad:
            ws = [self.lora_A[n].weight for n in self.adapter_order] + [self.lora_B[n].weight for n in self.adapter_order]
            y = _QVLoRAFn.apply(x3.contiguous(), self, idx, *ws)
        else:
            st = self._stacks()
            if T == 1 and B > 1:   # decode step: rows of different adapters share a tile
                y = ops.qv_lora_fwd_rows(x3.reshape(B, -1), st["W"], st["bias"], st["A"] if idx is not None else None,
                                         st["Bp"] if idx is not None else None, idx, st["scale"]).unsqueeze(1)
            else:
                y, _ = ops.qv_lora_fwd(x3, st["W"], st["bias"], st["A"] if idx is not None else None,
                                       st["Bp"] if idx is not None else None, idx, st["scale"])
        if drop_fix:
            y = y + self._dropout_correction(x3, idx).to(y.dtype)
        y = y.reshape(*lead, self.out_features)
        return y if in_dtype == torch.bfloat16 else y.to(in_dtype)

    # ------------------------------------------------------------------ lora_dropout > 0 (training mode only)
    def _dropout_active(self) -> bool:
        d = self.lora_dropout[self.active_adapter] if self.active_adapter in self.lora_dropout else None
        return isinstance(d, nn.Dropout) and d.p > 0

    def _dropout_correction(self, x3: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """PEFT drops the input of the A branch only: y = base(x) + s·B(A(drop(x))) (reference default p = 0.1,
        src/models/whisper_lora.py:30, scripts/train_lora.py:55).  The fused kernels compute base(x) + s·B(A(x)); this adds
        the zero-mean remainder s·B(A(drop(x) - x)) with plain autograd ops (two skinny GEMMs per adapter in use), so
        K1 / K3 stay on the main path and the sum is exactly PEFT's formula.  Utterance b uses adapter idx[b] (< 0: none);
        the mask is drawn once per call from the active adapter's Dropout module."""
        drop = self.lora_dropout[self.active_adapter]
        rem = drop(x3) - x3                                            # [B, T, d_in]: -x where dropped, x·p/(1-p) where kept
        out = None
        for k, name in enumerate(self.adapter_order):
            A, Bw = self.lora_A[name].weight, self.lora_B[name].weight
            sel = (idx == k).to(A.dtype).view(-1, 1, 1)                # no host sync: adapters not in the batch add zeros
            term = torch.nn.functional.linear(torch.nn.functional.linear(rem.to(A.dtype), A), Bw) * (self.scaling[name] * sel)
            out = term if out is None else out + term
        return out

    # ------------------------------------------------------------------ merge (PEFT merge_and_unload semantics)
    @torch.no_grad()
    def merged_linear(self, adapter: Optional[str] = None) -> nn.Linear:
        """nn.Linear with W + scaling·B·A of one adapter folded in (what PEFT's merge_and_unload returns)."""
        name = adapter or self.active_adapter
        lin = nn.Linear(self.in_features, self.out_features, bias=self.base_layer.bias is not None,
                        device=self.base_layer.weight.device, dtype=self.base_layer.weight.dtype)
        W = self.base_layer.weight.detach().float()
        if name is not None and not self.disable_adapters:
            W = W + self.scaling[name] * (self.lora_B[name].weight.float() @ self.lora_A[name].weight.float())
        lin.weight.copy_(W.to(lin.weight.dtype))
        if lin.bias is not None:
            lin.bias.copy_(self.base_layer.bias)
        lin.weight.requires_grad = False
        return lin

    def extra_repr(self) -> str:
        return (f"in={self.in_features}, out={self.out_features}, adapters={self.adapter_order}, "
                f"r={self.r}, active={self.active_adapter}")
