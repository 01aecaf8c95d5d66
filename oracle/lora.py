"""CPU oracle of the LoRA linear that the reference installs at every q_proj / v_proj (TEST INFRASTRUCTURE ONLY).

Restates PEFT ``lora.Linear.forward`` (third-party, not vendored; reference pin ``peft>=0.7.0``,
requirements.txt:6) as configured by the reference at src/models/whisper_lora.py:88-95
(r, lora_alpha, dropout, target_modules=["q_proj","v_proj"], bias="none"):

    y = x Wᵀ + b + (dropout(x) Aᵀ) Bᵀ · (lora_alpha / r)

and the per-utterance adapter selection of src/models/adapter_router.py:610-622 (utterance i uses adapter
``idx[i]`` for every LoRA'd module).  PARITY UNPINNED: the reference holds no expected tensor for this op.
"""
from __future__ import annotations

from typing import Optional

import torch


def lora_linear(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], A: Optional[torch.Tensor],
                B: Optional[torch.Tensor], scaling: float) -> torch.Tensor:
    """Single-adapter PEFT formula in the dtype of the inputs (fp32 for the check set).  A [r,d_in], B [d_out,r]."""
    y = torch.nn.functional.linear(x, W, bias)
    if A is not None:
        y = y + torch.nn.functional.linear(torch.nn.functional.linear(x, A), B) * scaling
    return y


def lora_linear_routed(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], A_stack: torch.Tensor,
                       B_stack: torch.Tensor, scaling: float, utt_adapter: torch.Tensor) -> torch.Tensor:
    """Reference semantics of hard routing at one module: utterance b is computed with adapter utt_adapter[b]
    exactly as the batch-1 loop of adapter_router.py:610-622 would (-1 = base model, no adapter).
    x [B,T,d_in]; A_stack [n,r,d_in]; B_stack [n,d_out,r]."""
    outs = []
    for b in range(x.shape[0]):
        k = int(utt_adapter[b])
        if k < 0:
            outs.append(lora_linear(x[b:b + 1], W, bias, None, None, scaling))
        else:
            outs.append(lora_linear(x[b:b + 1], W, bias, A_stack[k], B_stack[k], scaling))
    return torch.cat(outs, dim=0)


def lora_linear_routed_k1_rounding(x_bf16: torch.Tensor, W_bf16: torch.Tensor, bias_bf16: Optional[torch.Tensor],
                                   A_stack_bf16: torch.Tensor, B_stack_bf16: torch.Tensor, scaling: float,
                                   utt_adapter: torch.Tensor) -> torch.Tensor:
    """Same op with the fused kernel's rounding points, evaluated in fp32 on the bf16-valued inputs:
    u = bf16(scaling · x Aᵀ) (it is an MMA operand), everything else accumulated in fp32, one final bf16 rounding.
    Used for the tight (same-rounding) tolerance in the parity tests."""
    x = x_bf16.float()
    W = W_bf16.float()
    bias = None if bias_bf16 is None else bias_bf16.float()
    outs = []
    for b in range(x.shape[0]):
        y = torch.nn.functional.linear(x[b], W, bias)
        k = int(utt_adapter[b])
        if k >= 0:
            u = (torch.nn.functional.linear(x[b], A_stack_bf16[k].float()) * scaling).to(torch.bfloat16).float()
            y = y + torch.nn.functional.linear(u, B_stack_bf16[k].float())
        outs.append(y)
    return torch.stack(outs, dim=0).to(torch.bfloat16)


def lora_linear_backward(dy: torch.Tensor, x: torch.Tensor, W: torch.Tensor, A_stack: torch.Tensor,
                         B_stack: torch.Tensor, scaling: float, utt_adapter: torch.Tensor):
    """fp32 autograd of ``lora_linear_routed`` w.r.t. x, A_stack, B_stack (base W frozen) — what the reference
    trainer's loss.backward() computes through PEFT (src/training/trainer.py:251-256)."""
    x = x.detach().clone().requires_grad_(True)
    A = A_stack.detach().clone().requires_grad_(True)
    B = B_stack.detach().clone().requires_grad_(True)
    y = lora_linear_routed(x, W, None, A, B, scaling, utt_adapter)
    y.backward(dy)
    return x.grad, A.grad, B.grad
