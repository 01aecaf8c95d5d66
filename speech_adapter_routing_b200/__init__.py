"""speech_adapter_routing_b200 — B200-native (sm_100a) hot path of dhruv0811/speech-adapter-routing:
routed multi-adapter LoRA forward inside Whisper (q_proj / v_proj), the language-ID router head, and the
LoRA-only backward.  Python keeps the reference's public API; the arithmetic runs in libsar.so (include/sar.h).

Reference module  ->  module here
  src/models/whisper_lora.py     ->  whisper_adapters.py  (WhisperLoRA, create_whisper_lora, load_whisper_lora_from_checkpoint)
  src/models/adapter_router.py   ->  lid_router.py        (LanguageClassifier, EncoderFeatureExtractor, AdapterRouter)
  src/data/dataset.py:124-128    ->  logmel.py            (log_mel_spectrogram: the feature extractor, batched, on the GPU)
  src/models/base.py             ->  whisper_base.py      (load_base_model, get_processor, get_model_name, get_model_info)
  peft (third party)             ->  peft_compat.py       (LoraConfig, get_peft_model, PeftModel) + lora_linear.py
"""
from .whisper_base import (LANGUAGE_CODES, MODEL_NAME_MAP, get_model_info, get_model_name, get_processor,
                           load_base_model, whisper_config)
from .lora_linear import RoutedLoRALinear
from .peft_compat import LoraConfig, PeftModel, get_peft_model, inject_lora, lora_modules
from .whisper_adapters import WhisperLoRA, create_whisper_lora, load_whisper_lora_from_checkpoint
from .lid_router import AdapterRouter, EncoderFeatureExtractor, LanguageClassifier
from .routing import base_only, current_utt_adapter, refresh_operands, route, route_base, route_mix
from .whisper_blocks import install_fused_blocks, uninstall_fused_blocks
from .logmel import log_mel_spectrogram
from .train_graph import GraphedTrainStep
from .dist import FlatGradBucket

__all__ = [
    "WhisperLoRA", "create_whisper_lora", "load_whisper_lora_from_checkpoint", "load_base_model", "get_processor",
    "get_model_name", "get_model_info", "whisper_config", "MODEL_NAME_MAP", "LANGUAGE_CODES",
    "LanguageClassifier", "EncoderFeatureExtractor", "AdapterRouter", "RoutedLoRALinear", "LoraConfig", "PeftModel",
    "get_peft_model", "inject_lora", "lora_modules", "route", "route_base", "base_only", "current_utt_adapter",
    "install_fused_blocks", "uninstall_fused_blocks", "log_mel_spectrogram", "refresh_operands", "route_mix", "GraphedTrainStep", "FlatGradBucket",
]
