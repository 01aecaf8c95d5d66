"""CPU oracle of the routed multi-adapter Whisper forward (TEST INFRASTRUCTURE ONLY).

What the reference computes for a mixed-language batch (src/models/adapter_router.py):
  1. LID features: a full encoder pass on the BASE weights, no adapter         (:585, :441-474)
  2. language per utterance: LanguageClassifier.predict                         (:588, :550-566)
  3. hard routing: for every utterance i, a batch-1 forward of the WhisperLoRA of language_i
     (LoRA r/alpha on every q_proj and v_proj, src/models/whisper_lora.py:88-98); logits concatenated, loss =
     mean over utterances of each utterance's token-mean CE                      (:599-625, :695-713)
  4. generate: per-utterance greedy generate, right-padded with token id 0      (:715-761)

The backbone is the installed HF ``transformers`` Whisper (the reference's own dependency) built from a config
with random-init weights; the LoRA linear is oracle.lora (PEFT's formula).  Instead of n_adapters+1 copies of
Whisper this oracle keeps one copy whose LoRA modules switch adapter per utterance — arithmetically identical to
the reference's separate copies, because the base weights are shared and frozen.

PINNED against the reference run here: tests/golden/make_routed_golden.py loads the reference's unmodified
``AdapterRouter`` / ``EncoderFeatureExtractor`` / ``LanguageClassifier`` by file path, gives them the installed HF
Whisper as ``base_model`` and one full model copy per language (PEFT's formula at q_proj / v_proj) as ``adapters``, and
records forward (hard / soft / threshold, loss aggregation), detect_language and generate (EOS handling, zero
right-padding, ``language=``) — tests/golden/routed_forward_golden.pt.  tests/test_oracle_cpu.py checks that this
restatement reproduces those outputs bit for bit.  What stays unpinned is only PEFT's ``lora.Linear`` arithmetic itself
(``peft`` is neither vendored nor installable offline; oracle/lora.py restates its published formula).
"""
from __future__ import annotations

import json
import math
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from transformers import WhisperConfig, WhisperForConditionalGeneration

from . import lora as olora
from . import router as orouter

GEOMETRY = {  # d_model, layers, heads, ffn, mel bins, vocab
    "micro": (256, 2, 4, 512, 80, 1024),      # test-only geometry, head_dim 64 like every real Whisper (the fused blocks run)
    "micro32": (128, 2, 4, 256, 80, 1024),    # test-only geometry, head_dim 32: HF layer bodies over the K1 module slots
    "tiny": (384, 4, 6, 1536, 80, 51865),
    "base": (512, 6, 8, 2048, 80, 51865),
    "small": (768, 12, 12, 3072, 80, 51865),
    "medium": (1024, 24, 16, 4096, 80, 51865),
    "large-v3": (1280, 32, 20, 5120, 128, 51866),
    # two-layer cuts of the medium / large-v3 geometries (same d_model, heads, ffn, mel bins, vocabulary): every kernel
    # shape of BASELINE configs 3 / 4 at a depth the fp32 CPU oracle finishes in seconds
    "medium-2l": (1024, 2, 16, 4096, 80, 51865),
    "large-v3-2l": (1280, 2, 20, 5120, 128, 51866),
}


def make_config(geometry: str) -> WhisperConfig:
    d, layers, heads, ffn, mels, vocab = GEOMETRY[geometry]
    kw = dict(vocab_size=vocab, num_mel_bins=mels, d_model=d, encoder_layers=layers, decoder_layers=layers,
              encoder_attention_heads=heads, decoder_attention_heads=heads, encoder_ffn_dim=ffn,
              decoder_ffn_dim=ffn, max_source_positions=1500, max_target_positions=448)
    if vocab < 51865:  # test-only vocabulary: keep the special token ids inside it
        kw.update(pad_token_id=1, bos_token_id=2, eos_token_id=3, decoder_start_token_id=4,
                  begin_suppress_tokens=None, suppress_tokens=None)
    cfg = WhisperConfig(**kw)
    cfg.forced_decoder_ids = None
    cfg.suppress_tokens = []
    return cfg


def build_whisper(geometry: str, seed: int = 1234) -> WhisperForConditionalGeneration:
    """Random-init fp32 HF Whisper on the CPU (default HF init under the seed; no hub access)."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        model = WhisperForConditionalGeneration(make_config(geometry))
    model.eval()
    if getattr(model, "generation_config", None) is not None:
        model.generation_config.forced_decoder_ids = None
        model.generation_config.suppress_tokens = []
        model.generation_config.begin_suppress_tokens = None
    return model


def lora_module_paths(model: nn.Module, targets=("q_proj", "v_proj")) -> List[str]:
    """Every nn.Linear whose name ends in a target — the set PEFT injects (whisper_lora.py:61-62, :92)."""
    return [n for n, m in model.named_modules()
            if isinstance(m, nn.Linear) and n.rsplit(".", 1)[-1] in targets]


def make_adapter_weights(model: nn.Module, r: int, n_adapters: int, seed: int = 1235,
                         targets=("q_proj", "v_proj")) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """path -> (A [n,r,d_in], B [n,d_out,r]); A ~ kaiming-uniform(a=√5), B ~ N(0, 0.02²), values rounded to bf16 so
    the bf16 GPU path and the fp32 oracle start from identical numbers."""
    out = {}
    for i, path in enumerate(lora_module_paths(model, targets)):
        lin = model.get_submodule(path)
        g = torch.Generator().manual_seed(seed + 17 * i)
        bound = 1.0 / math.sqrt(lin.in_features)
        A = ((torch.rand(n_adapters, r, lin.in_features, generator=g) * 2 - 1) * bound).to(torch.bfloat16).float()
        B = (torch.randn(n_adapters, lin.out_features, r, generator=g) * 0.02).to(torch.bfloat16).float()
        out[path] = (A, B)
    return out


def write_peft_adapter(directory: Path, weights: Dict[str, Tuple[torch.Tensor, torch.Tensor]], k: int, r: int,
                       lora_alpha: float, base_name: str = "openai/whisper-small",
                       targets=("q_proj", "v_proj")) -> None:
    """Write adapter k in PEFT's on-disk layout (adapter_config.json + adapter_model.safetensors, keys
    ``base_model.model.<path>.lora_{A,B}.weight``) — an independent writer that the product's reader is tested
    against."""
    from safetensors.torch import save_file

    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    cfg = {"peft_type": "LORA", "task_type": None, "base_model_name_or_path": base_name, "r": r,
           "lora_alpha": lora_alpha, "lora_dropout": 0.0, "target_modules": list(targets), "bias": "none",
           "fan_in_fan_out": False, "inference_mode": True, "modules_to_save": None, "use_rslora": False,
           "use_dora": False, "init_lora_weights": True}
    (directory / "adapter_config.json").write_text(json.dumps(cfg, indent=2))
    sd = {}
    for path, (A, B) in weights.items():
        sd[f"base_model.model.{path}.lora_A.weight"] = A[k].contiguous()
        sd[f"base_model.model.{path}.lora_B.weight"] = B[k].contiguous()
    save_file(sd, str(directory / "adapter_model.safetensors"), metadata={"format": "pt"})


class OracleLoRALinear(nn.Module):
    """PEFT lora.Linear semantics for one *selected* adapter (``active`` = index, −1 = base only)."""

    def __init__(self, base: nn.Linear, A: torch.Tensor, B: torch.Tensor, scaling: float):
        super().__init__()
        self.base = base
        self.A, self.B, self.scaling = A, B, scaling
        self.active = -1
        self.mix: Optional[torch.Tensor] = None      # [n_adapters] weights of the current utterance (soft_fused)
        self.capture: Optional[list] = None

    def forward(self, x):
        k = self.active
        if self.mix is not None:
            # soft_fused (the product's opt-in strategy): y = base(x) + Σ_k w_k · s · B_k A_k x
            y = olora.lora_linear(x, self.base.weight, self.base.bias, None, None, self.scaling)
            for j in range(self.A.shape[0]):
                y = y + float(self.mix[j]) * self.scaling * F.linear(F.linear(x, self.A[j].to(x.dtype)), self.B[j].to(x.dtype))
            if self.capture is not None:
                self.capture.append(y.detach())
            return y
        y = olora.lora_linear(x, self.base.weight, self.base.bias, None if k < 0 else self.A[k].to(x.dtype),
                              None if k < 0 else self.B[k].to(x.dtype), self.scaling)
        if self.capture is not None:
            self.capture.append(y.detach())
        return y


class RoutedWhisperOracle:
    """The reference's routed forward, restated (see module docstring)."""

    def __init__(self, model: WhisperForConditionalGeneration,
                 weights: Dict[str, Tuple[torch.Tensor, torch.Tensor]], r: int, lora_alpha: float,
                 router_sd: Dict[str, torch.Tensor]):
        self.model = model
        self.router_sd = router_sd
        self.mods: Dict[str, OracleLoRALinear] = {}
        for path, (A, B) in weights.items():
            parent = model.get_submodule(path.rsplit(".", 1)[0])
            leaf = path.rsplit(".", 1)[-1]
            mod = OracleLoRALinear(getattr(parent, leaf), A, B, lora_alpha / r)
            setattr(parent, leaf, mod)
            self.mods[path] = mod

    def _select(self, k: int) -> None:
        for m in self.mods.values():
            m.active = k

    @torch.no_grad()
    def lid_features(self, input_features: torch.Tensor) -> torch.Tensor:
        self._select(-1)                                            # base weights (adapter_router.py:585)
        return self.model.model.encoder(input_features).last_hidden_state

    @torch.no_grad()
    def detect(self, input_features: torch.Tensor):
        h = self.lid_features(input_features)
        out = orouter.classifier_forward(h, self.router_sd)
        return out["probs"].argmax(-1), out, h

    @torch.no_grad()
    def forward_hard(self, input_features: torch.Tensor, decoder_input_ids: Optional[torch.Tensor],
                     labels: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None,
                     capture: bool = False):
        """Per-utterance batch-1 loop (adapter_router.py:610-622).  Returns dict(logits, loss, idx[, captured]).
        ``decoder_input_ids=None`` with labels: HF builds them from the labels (shift right), as when the reference is
        called with labels only."""
        if idx is None:
            idx, _, _ = self.detect(input_features)
        logits, losses = [], []
        cap: Dict[str, list] = {p: [] for p in self.mods} if capture else {}
        for i in range(input_features.shape[0]):
            self._select(int(idx[i]))
            if capture:
                for p, m in self.mods.items():
                    m.capture = []
            if decoder_input_ids is None:
                out = self.model(input_features=input_features[i:i + 1], labels=labels[i:i + 1])
                losses.append(out.loss)
            else:
                out = self.model(input_features=input_features[i:i + 1], decoder_input_ids=decoder_input_ids[i:i + 1])
                if labels is not None:
                    V = out.logits.shape[-1]
                    losses.append(F.cross_entropy(out.logits.reshape(-1, V), labels[i].reshape(-1), ignore_index=-100))
            logits.append(out.logits)
            if capture:
                for p, m in self.mods.items():
                    cap[p].append(m.capture[0])
                    m.capture = None
        res = {"logits": torch.cat(logits, 0), "idx": idx,
               "loss": torch.stack(losses).mean() if losses else None}       # (:707)
        if capture:
            res["captured"] = {p: torch.cat(v, 0) for p, v in cap.items()}
        return res

    @torch.no_grad()
    def forward_soft(self, input_features: torch.Tensor, labels: Optional[torch.Tensor] = None,
                     decoder_input_ids: Optional[torch.Tensor] = None, probs: Optional[torch.Tensor] = None):
        """Every adapter on the WHOLE batch, logits mixed by LID probability; loss = Σ_k mean_b(p[b,k]) · loss_k with
        loss_k HF's batch token-mean (adapter_router.py:627-670)."""
        if probs is None:
            _, out, _ = self.detect(input_features)
            probs = out["probs"]
        weighted, loss = None, None
        for k in range(probs.shape[1]):
            self._select(k)
            o = self.model(input_features=input_features, labels=labels, decoder_input_ids=decoder_input_ids)
            term = probs[:, k:k + 1, None] * o.logits
            weighted = term if weighted is None else weighted + term
            if labels is not None:
                l = probs[:, k].mean() * o.loss
                loss = l if loss is None else loss + l
        return {"logits": weighted, "loss": loss, "probs": probs}

    @torch.no_grad()
    def forward_soft_fused(self, input_features: torch.Tensor, decoder_input_ids: torch.Tensor, weights: torch.Tensor,
                           labels: Optional[torch.Tensor] = None):
        """The product's opt-in ``soft_fused`` strategy (not a reference strategy): per utterance ONE forward whose every
        LoRA'd projection applies Σ_k weights[b,k]·s·B_k A_k x; loss aggregated like hard routing."""
        logits, losses = [], []
        for i in range(input_features.shape[0]):
            for m in self.mods.values():
                m.mix = weights[i]
            out = self.model(input_features=input_features[i:i + 1], decoder_input_ids=decoder_input_ids[i:i + 1])
            logits.append(out.logits)
            if labels is not None:
                V = out.logits.shape[-1]
                losses.append(F.cross_entropy(out.logits.reshape(-1, V), labels[i].reshape(-1), ignore_index=-100))
        for m in self.mods.values():
            m.mix = None
        return {"logits": torch.cat(logits, 0), "loss": torch.stack(losses).mean() if losses else None}

    @torch.no_grad()
    def forward_threshold(self, input_features: torch.Tensor, threshold: float, labels: Optional[torch.Tensor] = None,
                          decoder_input_ids: Optional[torch.Tensor] = None):
        """Hard when every utterance's top probability exceeds the threshold, else soft (adapter_router.py:672-693: the
        mixed case also falls back to soft)."""
        idx, out, _ = self.detect(input_features)
        if bool((out["probs"].max(dim=-1).values > threshold).all()):
            r = self.forward_hard(input_features, decoder_input_ids, labels, idx=idx)
            return {"logits": r["logits"], "loss": r["loss"]}
        return self.forward_soft(input_features, labels, decoder_input_ids, probs=out["probs"])

    @torch.no_grad()
    def generate_language(self, input_features: torch.Tensor, k: int, **kwargs) -> torch.Tensor:
        """``generate(language=...)``: the named adapter's batched generate, returned unchanged (adapter_router.py:735-738)."""
        self._select(k)
        return self.model.generate(input_features=input_features, **kwargs)

    @torch.no_grad()
    def generate_hard(self, input_features: torch.Tensor, max_new_tokens: int = 16,
                      idx: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        """Per-utterance greedy generate, right-padded with 0 (adapter_router.py:744-761).  HF's Whisper ``generate``
        returns the NEW tokens without the closing EOS, so a finished row is its tokens before EOS, then zeros."""
        if idx is None:
            idx, _, _ = self.detect(input_features)
        outs = []
        for i in range(input_features.shape[0]):
            self._select(int(idx[i]))
            outs.append(self.model.generate(input_features=input_features[i:i + 1], max_new_tokens=max_new_tokens,
                                            num_beams=1, do_sample=False, **kwargs))
        L = max(o.shape[1] for o in outs)
        outs = [torch.cat([o, torch.zeros(1, L - o.shape[1], dtype=o.dtype)], 1) if o.shape[1] < L else o
                for o in outs]
        return torch.cat(outs, 0)


def make_input_features(B: int, n_mels: int, langs: List[int], n_langs: int, seed: int = 2234,
                        frames: int = 3000, template_seed: int = 4234) -> torch.Tensor:
    """Synthetic log-mel clips: 0.5·N(0,1) + a per-language band profile.  The band profiles
    depend only on ``template_seed`` so prototype clips and test clips share them."""
    gt = torch.Generator().manual_seed(template_seed)
    templates = torch.rand(n_langs, n_mels, generator=gt) * 2 - 1
    g = torch.Generator().manual_seed(seed)
    x = 0.5 * torch.randn(B, n_mels, frames, generator=g)
    return x + templates[torch.tensor(langs)][:, :, None]


def make_decoder_inputs(B: int, T_dec: int, vocab: int, start_id: int, seed: int = 3234):
    """decoder_input_ids [B,T_dec] (start token + uniform ids) and the matching labels (shifted, no −100)."""
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(5, vocab, (B, T_dec), generator=g)
    dec = torch.cat([torch.full((B, 1), start_id), body[:, :-1]], dim=1)
    return dec, body


def fit_router_head(router_sd: Dict[str, torch.Tensor], feats_per_lang: torch.Tensor,
                    gain: float = 8.0) -> Dict[str, torch.Tensor]:
    """Closed-form 'training' of the LID output layer so that the synthetic languages are separable with a wide
    margin: classifier.8 becomes a nearest-centroid rule on the penultimate activations of one prototype clip per
    language.  feats_per_lang: encoder states [C,T,d] of the prototypes (computed by the caller)."""
    sd = dict(router_sd)
    h = feats_per_lang.float()
    d = h.shape[-1]
    f = F.layer_norm(h, (d,), sd["layer_norm.weight"], sd["layer_norm.bias"], orouter.EPS).mean(1)
    z = F.linear(f, sd["classifier.0.weight"], sd["classifier.0.bias"])
    z = F.relu(F.layer_norm(z, (z.shape[-1],), sd["classifier.1.weight"], sd["classifier.1.bias"], orouter.EPS))
    z = F.linear(z, sd["classifier.4.weight"], sd["classifier.4.bias"])
    z = F.relu(F.layer_norm(z, (z.shape[-1],), sd["classifier.5.weight"], sd["classifier.5.bias"], orouter.EPS))
    c = z - z.mean(0, keepdim=True)
    norms = c.norm(dim=1, keepdim=True).clamp_min(1e-6)
    w = gain * (c / norms) / norms.mean()      # a prototype of language k scores ~gain on its own row
    sd["classifier.8.weight"] = w.contiguous()
    sd["classifier.8.bias"] = (-w @ z.mean(0)).contiguous()
    return sd
