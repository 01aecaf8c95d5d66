"""Seeded synthetic fixtures shared by the parity tests, smoke() and bench.py (TEST INFRASTRUCTURE ONLY).

Everything is generated on the CPU with an explicit ``torch.Generator`` so the CPU oracle and the GPU path see
identical bits.  Nothing here reads /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(seed)
    return g


@dataclass
class LoraCase:
    x: torch.Tensor            # bf16 [B,T,d_in]
    W: torch.Tensor            # bf16 [d_out,d_in]
    bias: Optional[torch.Tensor]  # bf16 [d_out]
    A_stack: torch.Tensor      # bf16 [n,r,d_in]
    B_stack: torch.Tensor      # bf16 [n,d_out,r]
    utt_adapter: torch.Tensor  # int32 [B]
    scaling: float


def make_lora_case(B: int, T: int, d_in: int, d_out: int, r: int, n_adapters: int, seed: int = 1234,
                   mix: str = "uniform", with_bias: bool = True, base_only_every: int = 0) -> LoraCase:
    """x ~ N(0,1); W ~ N(0, 0.02²) (HF Whisper init std); A ~ U(±1/sqrt(d_in)) (kaiming-uniform a=√5, PEFT's lora_A
    init); B ~ N(0, 0.02²) (PEFT's zero init would make parity vacuous); scaling = lora_alpha/r = 2 (reference scripts
    pass alpha = 2r, slurm_jobs/train_lora_array.sh:85).  `mix`: uniform | skewed | single | sorted."""
    g = _gen(seed)
    x = torch.randn(B, T, d_in, generator=g).to(torch.bfloat16)
    W = (torch.randn(d_out, d_in, generator=g) * 0.02).to(torch.bfloat16)
    bias = (torch.randn(d_out, generator=g) * 0.02).to(torch.bfloat16) if with_bias else None
    bound = 1.0 / math.sqrt(d_in)
    A = ((torch.rand(n_adapters, r, d_in, generator=g) * 2 - 1) * bound).to(torch.bfloat16)
    Bm = (torch.randn(n_adapters, d_out, r, generator=g) * 0.02).to(torch.bfloat16)
    if mix == "uniform":
        idx = torch.randint(0, n_adapters, (B,), generator=g)
    elif mix == "skewed":
        p = torch.full((n_adapters,), 0.3 / max(n_adapters - 1, 1))
        p[0] = 0.7 if n_adapters > 1 else 1.0
        idx = torch.multinomial(p, B, replacement=True, generator=g)
    elif mix == "single":
        idx = torch.full((B,), n_adapters - 1)
    elif mix == "sorted":
        idx = torch.sort(torch.randint(0, n_adapters, (B,), generator=g)).values
    else:
        raise ValueError(mix)
    idx = idx.to(torch.int32)
    if base_only_every:
        idx[::base_only_every] = -1
    return LoraCase(x, W, bias, A, Bm, idx, 2.0)


ROUTER_KEYS = ["layer_norm.weight", "layer_norm.bias", "classifier.0.weight", "classifier.0.bias",
               "classifier.1.weight", "classifier.1.bias", "classifier.4.weight", "classifier.4.bias",
               "classifier.5.weight", "classifier.5.bias", "classifier.8.weight", "classifier.8.bias"]


def make_router_state_dict(d: int, num_classes: int, hidden=(256, 128), seed: int = 1334,
                           final_gain: float = 8.0) -> Dict[str, torch.Tensor]:
    """State dict with the reference LanguageClassifier's key names and nn.Linear/LayerNorm default init
    (kaiming-uniform(a=√5) weights, U(±1/sqrt(fan_in)) biases, LN affine perturbed away from 1/0 so that the
    affine terms are exercised); the final layer is scaled by `final_gain` to spread the logits."""
    g = _gen(seed)

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b, (torch.rand(o, generator=g) * 2 - 1) * b

    def ln(n):
        return 1.0 + 0.1 * torch.randn(n, generator=g), 0.1 * torch.randn(n, generator=g)

    sd: Dict[str, torch.Tensor] = {}
    sd["layer_norm.weight"], sd["layer_norm.bias"] = ln(d)
    sd["classifier.0.weight"], sd["classifier.0.bias"] = lin(hidden[0], d)
    sd["classifier.1.weight"], sd["classifier.1.bias"] = ln(hidden[0])
    sd["classifier.4.weight"], sd["classifier.4.bias"] = lin(hidden[1], hidden[0])
    sd["classifier.5.weight"], sd["classifier.5.bias"] = ln(hidden[1])
    w, b = lin(num_classes, hidden[1])
    sd["classifier.8.weight"], sd["classifier.8.bias"] = w * final_gain, b * final_gain
    return sd


def make_encoder_states(B: int, T: int, d: int, num_classes: int, seed: int = 1434,
                        langs: Optional[List[int]] = None, dtype=torch.bfloat16) -> (torch.Tensor, torch.Tensor):
    """Synthetic encoder hidden states with a per-language direction so the LID has something separable:
    h[b,t,:] = N(0,1) + 1.5·template[lang_b].  Returns (h, langs)."""
    g = _gen(seed)
    templates = torch.randn(num_classes, d, generator=g)
    if langs is None:
        langs_t = torch.randint(0, num_classes, (B,), generator=g)
    else:
        langs_t = torch.tensor(langs)
    h = torch.randn(B, T, d, generator=g) + 1.5 * templates[langs_t][:, None, :]
    return h.to(dtype), langs_t


def language_mix(B: int, num_classes: int, kind: str, seed: int = 7, interleaved: bool = True) -> List[int]:
    """Language id per utterance for the routed fixtures: uniform | skewed (70/10/10/10) | single."""
    g = _gen(seed)
    if kind == "uniform":
        ids = [i % num_classes for i in range(B)]
    elif kind == "skewed":
        n0 = int(round(0.7 * B))
        rest = [1 + (i % max(num_classes - 1, 1)) for i in range(B - n0)] if num_classes > 1 else [0] * (B - n0)
        ids = [0] * n0 + rest
    elif kind == "single":
        ids = [num_classes - 1] * B
    else:
        raise ValueError(kind)
    if interleaved:
        perm = torch.randperm(B, generator=g).tolist()
        ids = [ids[p] for p in perm]
    else:
        ids = sorted(ids)
    return ids
