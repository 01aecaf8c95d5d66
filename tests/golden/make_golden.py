"""Generate golden vectors from the REFERENCE's own router implementation.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/src/models/adapter_router.py *by file path* (the module needs only torch; importing the
``src.models`` package would pull in ``peft``, which is not installed), instantiates the unmodified
``LanguageClassifier`` with its default architecture, and records inputs, state dict and outputs of
``forward`` / ``predict`` for a few seeded cases.  The committed result (router_golden.pt) pins oracle/router.py
and, through it, the K2 kernel.
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import torch

REF = Path("/root/reference/src/models/adapter_router.py")
OUT = Path(__file__).resolve().parent / "router_golden.pt"

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import fixtures  # noqa: E402


def load_reference_module():
    spec = importlib.util.spec_from_file_location("ref_adapter_router", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_module()
    cases = []
    # (B, T, d, C, seed, input dtype)
    for (B, T, d, C, seed, dt) in [(6, 50, 128, 4, 11, torch.float32), (5, 37, 256, 8, 12, torch.float32),
                                   (4, 64, 128, 4, 13, torch.bfloat16), (3, 40, 192, 4, 14, torch.float32)]:
        torch.manual_seed(seed)
        clf = ref.LanguageClassifier(input_dim=d, num_classes=C, languages=[f"l{i}" for i in range(C)])
        # default init gives near-tied logits on random features: perturb the LN affines and spread
        # the final layer so the argmax has a real margin.  The class itself is untouched.
        with torch.no_grad():
            clf.layer_norm.weight.add_(0.1 * torch.randn(d))
            clf.layer_norm.bias.add_(0.1 * torch.randn(d))
            clf.classifier[8].weight.mul_(8.0)
        clf.eval()
        h, langs = fixtures.make_encoder_states(B, T, d, C, seed=seed + 100, dtype=dt)
        with torch.no_grad():
            out = clf.forward(h.float())          # the reference head runs in fp32 (fp32 params)
            labels, probs = clf.predict(h.float())
        sd = {k: v.clone() for k, v in clf.state_dict().items()}
        cases.append({"B": B, "T": T, "d": d, "C": C, "h": h, "state_dict": sd, "logits": out["logits"].clone(),
                      "probs": out["probs"].clone(), "labels": labels.clone(), "predict_probs": probs.clone()})
        print(f"case d={d} C={C}: labels={labels.tolist()} min margin="
              f"{(out['logits'].topk(2).values[:,0]-out['logits'].topk(2).values[:,1]).min().item():.3f}")
    torch.save({"source": str(REF), "torch": str(torch.__version__), "cases": cases}, OUT)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
