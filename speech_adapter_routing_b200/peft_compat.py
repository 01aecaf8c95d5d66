"""Self-contained stand-ins for the slice of ``peft`` the reference uses (``peft`` is not installed in this image).

Mirrors, by name and behaviour, exactly what src/models/whisper_lora.py touches:
  LoraConfig(r, lora_alpha, lora_dropout, target_modules, bias)         :88-95
  get_peft_model(model, config)                                         :98
  PeftModel.from_pretrained(base, path) / .save_pretrained(path)        :228-231, :216, :309
  PeftModel.merge_and_unload()                                          :240
  attribute path  peft_model.base_model.model  (the HF model)           :168-184
and writes / reads PEFT's on-disk adapter layout unchanged:
  <dir>/adapter_config.json + <dir>/adapter_model.safetensors with keys
  ``base_model.model.<hf module path>.lora_{A,B}.weight`` (adapter name elided).
The injected module is RoutedLoRALinear (libsar K1/K3) instead of PEFT's eager lora.Linear.
"""
from __future__ import annotations

import json
import re
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Union

import torch
import torch.nn as nn

from .lora_linear import RoutedLoRALinear
from .routing import refresh_operands

ADAPTER_CONFIG = "adapter_config.json"
ADAPTER_WEIGHTS = "adapter_model.safetensors"
ADAPTER_WEIGHTS_BIN = "adapter_model.bin"


@dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: float = 8
    lora_dropout: float = 0.0
    target_modules: Optional[List[str]] = None
    bias: str = "none"
    task_type: Optional[str] = None
    base_model_name_or_path: Optional[str] = None
    inference_mode: bool = False
    fan_in_fan_out: bool = False
    modules_to_save: Optional[List[str]] = None
    use_rslora: bool = False
    use_dora: bool = False
    init_lora_weights: bool = True
    peft_type: str = "LORA"
    extra: Dict[str, object] = field(default_factory=dict)   # unknown JSON fields, round-tripped verbatim

    def to_json_dict(self) -> Dict[str, object]:
        d = asdict(self)
        extra = d.pop("extra")
        d.update({k: v for k, v in extra.items() if k not in d})
        if isinstance(d.get("target_modules"), (set, tuple)):
            d["target_modules"] = sorted(d["target_modules"])
        return d

    @classmethod
    def from_json_dict(cls, d: Dict[str, object]) -> "LoraConfig":
        known = {f for f in cls.__dataclass_fields__ if f != "extra"}
        kw = {k: v for k, v in d.items() if k in known}
        cfg = cls(**kw)
        cfg.extra = {k: v for k, v in d.items() if k not in known}
        return cfg


def _is_target(name: str, targets: Iterable[str]) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def inject_lora(model: nn.Module, config: LoraConfig, adapter_name: str = "default") -> List[str]:
    """Replace every nn.Linear whose qualified name ends in one of ``target_modules`` by a RoutedLoRALinear (in
    place, like PEFT), or add ``adapter_name`` to an already injected module.  Returns the module paths."""
    if config.bias != "none":
        raise NotImplementedError("only bias='none' is supported (the reference's setting)")
    if config.use_dora or config.use_rslora or config.fan_in_fan_out:
        raise NotImplementedError("DoRA / rsLoRA / fan_in_fan_out adapters are outside the reference's configuration")
    targets = list(config.target_modules or [])
    hit: List[str] = []
    for name, module in list(model.named_modules()):
        if not _is_target(name, targets):
            continue
        if isinstance(module, RoutedLoRALinear):
            if adapter_name not in module.lora_A:
                module.add_adapter(adapter_name, config.r, config.lora_alpha, config.lora_dropout)
            hit.append(name)
        elif isinstance(module, nn.Linear):
            parent = model.get_submodule(name.rsplit(".", 1)[0]) if "." in name else model
            setattr(parent, name.rsplit(".", 1)[-1],
                    RoutedLoRALinear(module, adapter_name, config.r, config.lora_alpha, config.lora_dropout))
            hit.append(name)
    if not hit:
        raise ValueError(f"target_modules {targets} matched no nn.Linear in the model")
    from .whisper_blocks import install_fused_blocks
    install_fused_blocks(model)   # (re)bind the fused Whisper block bodies over the new q_proj / v_proj modules
    return hit


def lora_modules(model: nn.Module) -> Dict[str, RoutedLoRALinear]:
    return {n: m for n, m in model.named_modules() if isinstance(m, RoutedLoRALinear)}


class LoraModel(nn.Module):
    """``peft_model.base_model`` — holds the HF model as ``.model`` (PEFT's LoraModel does the same)."""

    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model

    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == "model":
                raise
            return getattr(self.model, name)


class PeftModel(nn.Module):
    """``get_peft_model`` result: ``.base_model.model`` is the HF WhisperForConditionalGeneration whose q_proj /
    v_proj are RoutedLoRALinear; only LoRA parameters require grad."""

    def __init__(self, model: nn.Module, peft_config: LoraConfig, adapter_name: str = "default"):
        super().__init__()
        self.base_model = LoraModel(model)
        self.peft_config: Dict[str, LoraConfig] = {adapter_name: peft_config}
        self.active_adapter = adapter_name

    # ---- delegation -----------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        return self.base_model(*args, **kwargs)

    def generate(self, *args, **kwargs):
        return self.base_model.model.generate(*args, **kwargs)

    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == "base_model":
                raise
            return getattr(self.base_model, name)

    def get_base_model(self) -> nn.Module:
        return self.base_model.model

    # ---- adapters -------------------------------------------------------------------------------------
    def add_adapter(self, adapter_name: str, config: LoraConfig) -> None:
        inject_lora(self.base_model.model, config, adapter_name)
        self.peft_config[adapter_name] = config

    def set_adapter(self, adapter_name: str) -> None:
        for m in lora_modules(self.base_model.model).values():
            m.set_adapter(adapter_name)
        self.active_adapter = adapter_name

    def adapter_names(self) -> List[str]:
        mods = lora_modules(self.base_model.model)
        return list(next(iter(mods.values())).adapter_order) if mods else []

    def print_trainable_parameters(self) -> None:
        t = sum(p.numel() for p in self.parameters() if p.requires_grad)
        a = sum(p.numel() for p in self.parameters())
        print(f"trainable params: {t:,} || all params: {a:,} || trainable%: {100 * t / a:.4f}")

    # ---- PEFT checkpoint layout -----------------------------------------------------------------------
    def adapter_state_dict(self, adapter_name: Optional[str] = None) -> Dict[str, torch.Tensor]:
        name = adapter_name or self.active_adapter
        out: Dict[str, torch.Tensor] = {}
        tag = f".{name}.weight"
        for k, v in self.state_dict().items():
            if (".lora_A." in k or ".lora_B." in k) and k.endswith(tag):
                out[k[: -len(tag)] + ".weight"] = v.detach().cpu().contiguous()
        return out

    def save_pretrained(self, save_directory: Union[str, Path], adapter_name: Optional[str] = None,
                        safe_serialization: bool = True) -> None:
        name = adapter_name or self.active_adapter
        path = Path(save_directory)
        path.mkdir(parents=True, exist_ok=True)
        cfg = self.peft_config[name]
        d = cfg.to_json_dict()
        d["inference_mode"] = True   # PEFT writes inference_mode=true on save
        if d.get("base_model_name_or_path") is None:
            d["base_model_name_or_path"] = getattr(self.base_model.model.config, "_name_or_path", None) or None
        (path / ADAPTER_CONFIG).write_text(json.dumps(d, indent=2, sort_keys=True))
        sd = self.adapter_state_dict(name)
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, str(path / ADAPTER_WEIGHTS), metadata={"format": "pt"})
        else:
            torch.save(sd, path / ADAPTER_WEIGHTS_BIN)

    def load_adapter(self, model_id: Union[str, Path], adapter_name: str = "default",
                     is_trainable: bool = False) -> None:
        path = Path(model_id)
        cfg = LoraConfig.from_json_dict(json.loads((path / ADAPTER_CONFIG).read_text()))
        cfg.inference_mode = not is_trainable
        mods = lora_modules(self.base_model.model)
        if not mods or adapter_name not in next(iter(mods.values())).lora_A:
            inject_lora(self.base_model.model, cfg, adapter_name)
        self.peft_config[adapter_name] = cfg
        sd = load_adapter_weights(path)
        own = dict(self.named_parameters())
        missing = []
        for k, v in sd.items():
            m = re.match(r"(.*\.lora_[AB])\.weight$", k)
            if not m:
                continue
            full = f"{m.group(1)}.{adapter_name}.weight"
            if full not in own:
                missing.append(k)
                continue
            with torch.no_grad():
                own[full].copy_(v.to(own[full].dtype))
        if missing:
            raise KeyError(f"adapter tensors without a matching module: {missing[:4]} ...")
        refresh_operands()   # new adapter values: every cached bf16 operand pack / decode graph rebuilds on next use
        for n, p in own.items():
            if f".{adapter_name}." in n and (".lora_A." in n or ".lora_B." in n):
                p.requires_grad = is_trainable

    @classmethod
    def from_pretrained(cls, model: nn.Module, model_id: Union[str, Path], adapter_name: str = "default",
                        is_trainable: bool = False) -> "PeftModel":
        if isinstance(model, LoraModel):   # reference passes self.model.base_model (whisper_lora.py:228-231)
            model = model.model
        path = Path(model_id)
        cfg = LoraConfig.from_json_dict(json.loads((path / ADAPTER_CONFIG).read_text()))
        pm = cls.__new__(cls)
        nn.Module.__init__(pm)
        pm.base_model = LoraModel(model)
        pm.peft_config = {}
        pm.active_adapter = adapter_name
        for p in model.parameters():
            p.requires_grad = False
        pm.load_adapter(path, adapter_name, is_trainable=is_trainable)
        pm.set_adapter(adapter_name)
        return pm

    @torch.no_grad()
    def merge_and_unload(self, adapter_name: Optional[str] = None) -> nn.Module:
        """Fold one adapter into the base weights and return the plain HF model (PEFT semantics)."""
        name = adapter_name or self.active_adapter
        model = self.base_model.model
        for path, m in list(lora_modules(model).items()):
            parent = model.get_submodule(path.rsplit(".", 1)[0]) if "." in path else model
            setattr(parent, path.rsplit(".", 1)[-1], m.merged_linear(name))
        return model


def load_adapter_weights(path: Path) -> Dict[str, torch.Tensor]:
    if (path / ADAPTER_WEIGHTS).exists():
        from safetensors.torch import load_file
        return load_file(str(path / ADAPTER_WEIGHTS))
    if (path / ADAPTER_WEIGHTS_BIN).exists():
        return torch.load(path / ADAPTER_WEIGHTS_BIN, map_location="cpu")
    raise FileNotFoundError(f"no {ADAPTER_WEIGHTS} / {ADAPTER_WEIGHTS_BIN} in {path}")


def get_peft_model(model: nn.Module, peft_config: LoraConfig, adapter_name: str = "default") -> PeftModel:
    """Freeze the base model, inject RoutedLoRALinear at ``target_modules`` and wrap (PEFT's entry point)."""
    for p in model.parameters():
        p.requires_grad = False
    inject_lora(model, peft_config, adapter_name)
    return PeftModel(model, peft_config, adapter_name)
