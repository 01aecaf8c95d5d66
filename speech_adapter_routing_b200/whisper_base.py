"""Base-model utilities — same public surface as the reference's src/models/base.py
(MODEL_NAME_MAP :16-24, LANGUAGE_CODES :27-36, get_model_name :39, get_processor :44, load_base_model :77,
get_model_info :142), plus an offline path: this image has no network, so ``load_base_model`` can build the
architecture from a geometry table with random-init weights (``random_init=True`` or ``SAR_RANDOM_INIT=1``).
"""
from __future__ import annotations

import logging
import os
from typing import Optional

import torch
from transformers import WhisperConfig, WhisperForConditionalGeneration

logger = logging.getLogger(__name__)

MODEL_NAME_MAP = {
    "whisper-tiny": "openai/whisper-tiny",
    "whisper-base": "openai/whisper-base",
    "whisper-small": "openai/whisper-small",
    "whisper-medium": "openai/whisper-medium",
    "whisper-large": "openai/whisper-large-v3",
    "whisper-large-v2": "openai/whisper-large-v2",
    "whisper-large-v3": "openai/whisper-large-v3",
}

LANGUAGE_CODES = {
    "hindi": "hi", "italian": "it", "punjabi": "pa", "telugu": "te",
    "english": "en", "german": "de", "french": "fr", "spanish": "es",
}

# d_model, layers, heads, ffn, mel bins, vocab — the published OpenAI geometries (reference
# configs/model_configs/whisper.yaml:3-28 lists the same widths/depths)
_GEOMETRY = {
    "openai/whisper-tiny": (384, 4, 6, 1536, 80, 51865),
    "openai/whisper-base": (512, 6, 8, 2048, 80, 51865),
    "openai/whisper-small": (768, 12, 12, 3072, 80, 51865),
    "openai/whisper-medium": (1024, 24, 16, 4096, 80, 51865),
    "openai/whisper-large-v2": (1280, 32, 20, 5120, 80, 51865),
    "openai/whisper-large-v3": (1280, 32, 20, 5120, 128, 51866),
}


def get_model_name(model_id: str) -> str:
    return MODEL_NAME_MAP.get(model_id, model_id)


def whisper_config(model_name: str, **overrides) -> WhisperConfig:
    """Hub-free WhisperConfig for a known geometry."""
    name = get_model_name(model_name)
    if name not in _GEOMETRY:
        raise KeyError(f"no built-in geometry for {model_name!r}")
    d, layers, heads, ffn, mels, vocab = _GEOMETRY[name]
    kw = dict(vocab_size=vocab, num_mel_bins=mels, d_model=d, encoder_layers=layers, decoder_layers=layers,
              encoder_attention_heads=heads, decoder_attention_heads=heads, encoder_ffn_dim=ffn,
              decoder_ffn_dim=ffn, max_source_positions=1500, max_target_positions=448)
    kw.update(overrides)
    return WhisperConfig(**kw)


def get_processor(model_name: str, language: Optional[str] = None, task: str = "transcribe",
                  cache_dir: Optional[str] = None):
    """WhisperProcessor.from_pretrained with the reference's language-name mapping (base.py:44-74)."""
    from transformers import WhisperProcessor

    model_name = get_model_name(model_name)
    if language:
        language = LANGUAGE_CODES.get(language.lower(), language)
    return WhisperProcessor.from_pretrained(model_name, language=language, task=task, cache_dir=cache_dir)


def _random_init_requested(flag: Optional[bool]) -> bool:
    return bool(flag) if flag is not None else os.environ.get("SAR_RANDOM_INIT", "0") == "1"


def load_base_model(model_name: str, device: Optional[str] = None, dtype: Optional[torch.dtype] = None,
                    cache_dir: Optional[str] = None, use_flash_attention: bool = False,
                    random_init: Optional[bool] = None, seed: int = 1234) -> WhisperForConditionalGeneration:
    """Reference semantics (base.py:77-139): bf16 on CUDA / fp32 on CPU unless ``dtype`` is given;
    ``forced_decoder_ids`` cleared and ``suppress_tokens`` emptied; model moved to ``device``."""
    model_name = get_model_name(model_name)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if dtype is None:
        if str(device).startswith("cuda"):
            dtype = torch.bfloat16 if torch.cuda.is_bf16_supported() else torch.float16
        else:
            dtype = torch.float32
    logger.info("Loading %s to %s with dtype %s", model_name, device, dtype)
    if _random_init_requested(random_init):
        cfg = whisper_config(model_name)
        if use_flash_attention:
            cfg._attn_implementation = "flash_attention_2"
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(seed)
            model = WhisperForConditionalGeneration(cfg)
        model = model.to(dtype)
    else:
        kwargs = {"cache_dir": cache_dir, "torch_dtype": dtype}
        if use_flash_attention:
            kwargs["attn_implementation"] = "flash_attention_2"
        model = WhisperForConditionalGeneration.from_pretrained(model_name, **kwargs)
    model.config.forced_decoder_ids = None
    model.config.suppress_tokens = []
    if getattr(model, "generation_config", None) is not None:
        model.generation_config.forced_decoder_ids = None
        model.generation_config.suppress_tokens = []
    model.to(device)
    from .whisper_blocks import install_fused_blocks
    install_fused_blocks(model)
    logger.info("Loaded model with %.1fM parameters", sum(p.numel() for p in model.parameters()) / 1e6)
    return model


def get_model_info(model_name: str) -> dict:
    model_name = get_model_name(model_name)
    try:
        config = whisper_config(model_name)
    except KeyError:
        config = WhisperConfig.from_pretrained(model_name)
    return {
        "name": model_name,
        "hidden_size": config.d_model,
        "encoder_layers": config.encoder_layers,
        "decoder_layers": config.decoder_layers,
        "encoder_attention_heads": config.encoder_attention_heads,
        "decoder_attention_heads": config.decoder_attention_heads,
        "encoder_ffn_dim": config.encoder_ffn_dim,
        "decoder_ffn_dim": config.decoder_ffn_dim,
        "vocab_size": config.vocab_size,
        "max_source_positions": config.max_source_positions,
        "max_target_positions": config.max_target_positions,
    }
