"""CPU oracle of the Whisper-block ops around the routed q/v projections (TEST INFRASTRUCTURE ONLY — imported by
tests/, smoke() and bench.py's CPU legs, never by the product path).

Each function restates, in fp32 on the CPU, what HF transformers' Whisper computes at the cited line of
$HF/models/whisper/modeling_whisper.py (the reference repo delegates its whole forward to that code through PEFT:
src/models/whisper_lora.py:114-143), with the adapter term of oracle.lora at q_proj / v_proj.
PARITY: pinned against the installed HF modules themselves in tests/test_oracle_cpu.py (transformers 5.5.0 is the
backbone the reference would run here); the LoRA term is unpinned as explained in oracle/lora.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import lora as olora


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm over the last dim (:391, :401, :470, :486, :498, :643, :797), fp32 statistics."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight.float(), bias.float(), eps)


def dense(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor] = None,
          gelu: bool = False) -> torch.Tensor:
    """act(x Wᵀ + b) + residual: out_proj + residual add (:352-353, :399), fc1 + erf-GELU (:403), fc2 + residual (:405-407)."""
    y = F.linear(x.float(), W.float(), None if bias is None else bias.float())
    if gelu:
        y = F.gelu(y)   # ACT2FN["gelu"]: exact erf form
    if residual is not None:
        y = y + residual.float()
    return y


def attn_projections(x: torch.Tensor, Ws: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]],
                     loras: Sequence[Optional[tuple]], scaling: float, utt_adapter: torch.Tensor,
                     out_scales: Sequence[float], heads: int) -> List[torch.Tensor]:
    """q / k / v of one WhisperAttention.forward (:310-312, :331-334): every projection of the same x, the routed LoRA
    term on those that carry one (``loras[i] = (A_stack, B_stack)``), ``* self.scaling`` on q, and the
    view(B, T, h, d/h).transpose(1, 2) head-major layout.  bf16-valued inputs, fp32 arithmetic with K1's rounding of u."""
    B, T, _ = x.shape
    none = torch.full((B,), -1, dtype=torch.int32)
    outs = []
    for W, b, lo, s in zip(Ws, biases, loras, out_scales):
        if lo is None:
            y = olora.lora_linear_routed_k1_rounding(x, W, b, W[:0].reshape(0, 1, W.shape[1]), W[:0].reshape(0, W.shape[0], 1),
                                                     scaling, none).float()
        else:
            y = olora.lora_linear_routed_k1_rounding(x, W, b, lo[0], lo[1], scaling, utt_adapter).float()
        y = y * s
        outs.append(y.view(B, T, heads, -1).transpose(1, 2))
    return outs


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False,
              n_keys: Optional[int] = None) -> torch.Tensor:
    """softmax(q kᵀ) v per head — eager_attention_forward (:215-238) with scaling = 1 (the query is already scaled by
    the projection) and HF's additive causal mask written as an explicit ``key <= query`` rule; ``n_keys`` limits the
    keys to the first n (a static decode cache attended up to the current position).  q [B, h, Tq, dh], k / v
    [B, h, Tk, dh], fp32 arithmetic; returns the head-major [B, h, Tq, dh] (HF transposes it back before out_proj)."""
    q, k, v = q.float(), k.float(), v.float()
    if n_keys is not None:
        k, v = k[:, :, :n_keys], v[:, :, :n_keys]
    w = torch.matmul(q, k.transpose(2, 3))
    if causal:
        Tq, Tk = q.shape[2], k.shape[2]
        allowed = torch.arange(Tk)[None, :] <= (torch.arange(Tq)[:, None] + (Tk - Tq))
        w = w.masked_fill(~allowed, float("-inf"))
    return torch.matmul(torch.softmax(w, dim=-1), v)


def conv_frontend(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                  pos: torch.Tensor, round_mid_to_bf16: bool = True) -> torch.Tensor:
    """gelu(conv1) -> gelu(conv2) -> permute -> + embed_positions (:626-633).  ``round_mid_to_bf16`` rounds conv1's
    activation to bf16 like a bf16 model stores it (HF) / like the GEMM path writes its frame buffer."""
    h = F.gelu(F.conv1d(x.float(), w1.float(), b1.float(), padding=1))
    if round_mid_to_bf16:
        h = h.to(torch.bfloat16).float()
    h = F.gelu(F.conv1d(h, w2.float(), b2.float(), stride=2, padding=1))
    return h.permute(0, 2, 1) + pos.float()


def lm_head(x: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """proj_out (:1135): logits = h Wᵀ, no bias."""
    return F.linear(x.float(), W.float())
