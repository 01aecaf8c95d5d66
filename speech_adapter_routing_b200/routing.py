"""Per-batch adapter-index bookkeeping.

HF's WhisperAttention passes nothing down to ``q_proj`` / ``v_proj`` except the hidden states, so the routed
adapter index reaches the projection modules out of band: a thread-local context that the router / wrapper sets
for the duration of one forward.  The value is a *device* int32 tensor ``utt_adapter[B]`` (−1 = base weights
only), written by the K2 router kernel — nothing is copied to the host, unlike the reference's
``[self.languages[l.item()] for l in labels]`` (src/models/adapter_router.py:565).
"""
from __future__ import annotations

import threading
from contextlib import contextmanager
from typing import Optional

import torch


class _RoutingState(threading.local):
    def __init__(self) -> None:
        self.utt_adapter: Optional[torch.Tensor] = None
        self.active = False
        self.base_only = False
        self.mix_weights: Optional[torch.Tensor] = None


_state = _RoutingState()


def current_utt_adapter() -> Optional[torch.Tensor]:
    """int32 [B] device tensor set by ``route(...)``, or None when no routing context is active."""
    return _state.utt_adapter if _state.active else None


def routing_active() -> bool:
    return _state.active


def routing_base_only() -> bool:
    """True inside ``route_base()``: every LoRA'd projection runs on its base weights (dense kernel, no adapter
    operands streamed) — the LID feature pass of the reference runs on the un-adapted encoder (adapter_router.py:441-474)."""
    return _state.base_only


@contextmanager
def route_base():
    prev = (_state.utt_adapter, _state.active, _state.base_only, _state.mix_weights)
    _state.utt_adapter, _state.active, _state.base_only, _state.mix_weights = None, False, True, None
    try:
        yield
    finally:
        _state.utt_adapter, _state.active, _state.base_only, _state.mix_weights = prev


def current_mix_weights() -> Optional[torch.Tensor]:
    """fp32 [B, n_adapters] device tensor set by ``route_mix(...)`` (soft_fused routing), or None."""
    return _state.mix_weights


@contextmanager
def route_mix(weights: torch.Tensor):
    """Run the enclosed forward with a per-utterance WEIGHTED MIX of all stacked adapters inside every LoRA'd projection:
    y = base(x) + Σ_k weights[b, k] · s · B_k A_k x (lid_router.AdapterRouter strategy "soft_fused")."""
    w = weights.detach().to(torch.float32).contiguous()
    prev = (_state.utt_adapter, _state.active, _state.base_only, _state.mix_weights)
    _state.utt_adapter = torch.zeros(w.shape[0], dtype=torch.int32, device=w.device)   # the one merged adapter
    _state.active, _state.base_only, _state.mix_weights = True, False, w
    try:
        yield
    finally:
        _state.utt_adapter, _state.active, _state.base_only, _state.mix_weights = prev


@contextmanager
def route(utt_adapter: Optional[torch.Tensor]):
    """Run the enclosed forward with per-utterance adapter indices.  ``None`` restores module defaults
    (every utterance uses the module's active adapter)."""
    if utt_adapter is not None:
        if utt_adapter.dtype != torch.int32:
            utt_adapter = utt_adapter.to(torch.int32)
        utt_adapter = utt_adapter.contiguous()
    prev = (_state.utt_adapter, _state.active, _state.base_only, _state.mix_weights)
    _state.utt_adapter, _state.active, _state.base_only, _state.mix_weights = utt_adapter, utt_adapter is not None, False, None
    try:
        yield
    finally:
        _state.utt_adapter, _state.active, _state.base_only, _state.mix_weights = prev


def base_only(batch_size: int, device) -> torch.Tensor:
    """utt_adapter selecting no adapter for every utterance (the LID feature pass runs on base weights)."""
    return torch.full((batch_size,), -1, dtype=torch.int32, device=device)


# ---- operand-cache epoch ---------------------------------------------------------------------------------------
# Every bf16 operand cache of the package (RoutedLoRALinear._stacks, the fused blocks' packs, the K2 parameter pack,
# the decode graph) is keyed on (data_ptr, tensor._version) of its source parameters.  In-place ops under no_grad
# (optimizer steps, ``copy_``, ``load_state_dict``) bump ``_version``; writes through ``param.data`` do NOT.  The
# epoch below is part of every key: ``refresh_operands()`` bumps it, which makes every cache rebuild on next use.
# load_adapter / add_adapter / load_state_dict paths of this package call it; user code that writes through ``.data``
# (or through a raw pointer) must call it too.
_operand_epoch = 0


def operand_epoch() -> int:
    return _operand_epoch


def refresh_operands() -> int:
    """Invalidate every cached bf16 operand pack and captured decode graph (call after ``param.data`` writes)."""
    global _operand_epoch
    _operand_epoch += 1
    return _operand_epoch
