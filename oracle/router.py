"""CPU oracle of the language-ID router head (TEST INFRASTRUCTURE ONLY).

Restates, with plain tensor ops and the reference's state-dict key names, the default-configuration
``LanguageClassifier`` of /root/reference/src/models/adapter_router.py:
  forward  :251-293   layer_norm (:268) -> mean pooling (:229) -> classifier MLP (:84-97, :281) -> softmax (:282)
  predict  :295-312   argmax over probs (:311)
and the adapter-index bookkeeping that AdapterRouter.detect_language (:550-566) keeps as a Python list.

PINNED against the reference's own class: tests/golden/router_golden.pt is produced by importing
adapter_router.py by file path (tests/golden/make_golden.py) and tests/test_oracle_router.py checks this
restatement against it.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.LayerNorm default, used by layer_norm and classifier.{1,5}


def classifier_forward(h: torch.Tensor, sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """h [B,T,d] -> {"logits","probs"} in fp32.  `sd` uses the reference state-dict keys
    (layer_norm.*, classifier.{0,1,4,5,8}.*).  Dropout layers (classifier.{3,7}) are identity in eval mode."""
    h = h.float()
    d = h.shape[-1]
    f = F.layer_norm(h, (d,), sd["layer_norm.weight"].float(), sd["layer_norm.bias"].float(), EPS)  # :268
    pooled = f.mean(dim=1)                                                                           # :229
    z = F.linear(pooled, sd["classifier.0.weight"].float(), sd["classifier.0.bias"].float())
    z = F.relu(F.layer_norm(z, (z.shape[-1],), sd["classifier.1.weight"].float(), sd["classifier.1.bias"].float(), EPS))
    z = F.linear(z, sd["classifier.4.weight"].float(), sd["classifier.4.bias"].float())
    z = F.relu(F.layer_norm(z, (z.shape[-1],), sd["classifier.5.weight"].float(), sd["classifier.5.bias"].float(), EPS))
    logits = F.linear(z, sd["classifier.8.weight"].float(), sd["classifier.8.bias"].float())       # :281
    probs = F.softmax(logits, dim=-1)                                                                # :282
    return {"logits": logits, "probs": probs}


def predict(h: torch.Tensor, sd: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(labels int64 [B], probs [B,C]) — LanguageClassifier.predict (:295-312)."""
    out = classifier_forward(h, sd)
    return out["probs"].argmax(dim=-1), out["probs"]


def segments(idx: torch.Tensor, num_classes: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Adapter-index bookkeeping: utterances stably sorted by adapter id, and per-adapter segment offsets.
    perm int32 [B]; seg_starts int32 [C+1] with segment k = perm[seg_starts[k]:seg_starts[k+1]]."""
    idx = idx.to(torch.int64)
    perm = torch.sort(idx, stable=True).indices.to(torch.int32)
    counts = torch.bincount(idx, minlength=num_classes)
    seg = torch.zeros(num_classes + 1, dtype=torch.int32)
    seg[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return perm, seg


def top2_margin(logits: torch.Tensor) -> torch.Tensor:
    """Per-utterance gap between the best and second-best logit (fixtures must keep it well above fp32 noise so
    that bit-exact index parity is a property of the kernel, not of luck)."""
    top = logits.topk(2, dim=-1).values
    return top[:, 0] - top[:, 1]
