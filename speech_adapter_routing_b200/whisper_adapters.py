"""WhisperLoRA — drop-in for the reference's src/models/whisper_lora.py (same constructor, methods, attributes and
attribute paths), with the q_proj / v_proj LoRA path running on libsar's fused sm_100a kernels.

Reference surface mirrored here (file:line in /root/reference/src/models/whisper_lora.py):
  WhisperLoRA.__init__ :25-101, forward :114-143, generate :145-186, decode :188-205, save_adapter :207-217,
  load_adapter :219-232, merge_and_unload :234-240, train/eval :242-250,
  create_whisper_lora :253-280, load_whisper_lora_from_checkpoint :283-325.

Documented deviations:
  * ``load_whisper_lora_from_checkpoint`` builds the wrapper with a proper ``nn.Module.__init__`` — the
    reference's ``__new__`` + attribute assignment raises on current torch.
  * additional keyword ``random_init`` (and env SAR_RANDOM_INIT=1) builds the architecture without the hub; the
    processor is then ``None`` unless the tokenizer files are cached locally.
"""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn

from .whisper_base import _random_init_requested, get_model_name, get_processor, load_base_model
from .peft_compat import LoraConfig, PeftModel, get_peft_model

logger = logging.getLogger(__name__)


def _try_processor(model_name, language, task, cache_dir=None, offline: bool = False):
    """Processor as the reference loads it (base.py:44-74).  In offline / random-init mode only locally cached
    tokenizer files are considered and a missing processor is tolerated (it is not on the compute path)."""
    try:
        if offline:
            from transformers import WhisperProcessor
            from .whisper_base import LANGUAGE_CODES
            lang = LANGUAGE_CODES.get(language.lower(), language) if language else None
            return WhisperProcessor.from_pretrained(get_model_name(model_name), language=lang, task=task,
                                                    cache_dir=cache_dir, local_files_only=True)
        return get_processor(model_name, language=language, task=task, cache_dir=cache_dir)
    except Exception as e:
        if not offline:
            raise
        logger.warning("WhisperProcessor unavailable offline (%s); continuing without a processor", type(e).__name__)
        return None


class WhisperLoRA(nn.Module):
    """Whisper with LoRA adapters on q_proj / v_proj."""

    def __init__(self, model_name: str, lora_r: int = 16, lora_alpha: int = 32, lora_dropout: float = 0.1,
                 target_modules: Optional[List[str]] = None, language: Optional[str] = None,
                 task: str = "transcribe", device: Optional[str] = None, dtype: Optional[torch.dtype] = None,
                 cache_dir: Optional[str] = None, use_gradient_checkpointing: bool = True,
                 random_init: Optional[bool] = None):
        super().__init__()
        self.model_name = get_model_name(model_name)
        self.language = language
        self.task = task
        self.device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        if target_modules is None:
            target_modules = ["q_proj", "v_proj"]
        offline = _random_init_requested(random_init)
        self.processor = _try_processor(self.model_name, language, task, cache_dir, offline=offline)
        self.model = load_base_model(self.model_name, device=self.device, dtype=dtype, cache_dir=cache_dir,
                                     random_init=random_init)
        if use_gradient_checkpointing:
            self.model.gradient_checkpointing_enable()
            self.model.config.use_cache = False
        # no task_type, as in the reference (:86-95): PEFT's seq2seq wrapper would inject input_ids handling
        self.lora_config = LoraConfig(r=lora_r, lora_alpha=lora_alpha, lora_dropout=lora_dropout,
                                      target_modules=list(target_modules), bias="none",
                                      base_model_name_or_path=self.model_name)
        self.model = get_peft_model(self.model, self.lora_config)
        self._log_trainable_params()

    def _log_trainable_params(self) -> None:
        trainable = sum(p.numel() for p in self.model.parameters() if p.requires_grad)
        total = sum(p.numel() for p in self.model.parameters())
        logger.info("Trainable params: %s / %s (%.2f%%)", f"{trainable:,}", f"{total:,}", 100 * trainable / total)

    def forward(self, input_features: torch.Tensor, labels: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, decoder_input_ids: Optional[torch.Tensor] = None,
                decoder_attention_mask: Optional[torch.Tensor] = None, **kwargs) -> Dict[str, torch.Tensor]:
        # only the five supported arguments are forwarded; other kwargs are swallowed like the reference (:136-143)
        extra = {}
        if self.training and torch.is_grad_enabled() and labels is not None:
            self._sync_lora_operands(input_features)
            # teacher-forced training never reads the KV cache HF would build alongside the loss (HF itself switches
            # it off under gradient checkpointing); without it the decoder layers take the fused forward+backward bodies
            extra["use_cache"] = False
        return self.model(input_features=input_features, labels=labels, attention_mask=attention_mask,
                          decoder_input_ids=decoder_input_ids, decoder_attention_mask=decoder_attention_mask, **extra)

    def _sync_lora_operands(self, input_features: torch.Tensor) -> None:
        """Training loop (src/training/trainer.py:251-268): after ``optimizer.step()`` the kernels' cached bf16 LoRA
        operands are stale.  From the third step on they are re-derived by ONE launch (operand_refresh.py) instead of the
        per-module host-side rebuild (~25 small kernels for each of the 72 LoRA'd projections and 48 fused calls)."""
        if not input_features.is_cuda or torch.cuda.is_current_stream_capturing():
            return                          # a captured step carries its own refresh node (train_graph.py)
        d = self.__dict__
        r = d.get("_sar_refresh")
        if r is not None:
            if r.maybe_refresh():
                return
            d["_sar_refresh"] = None
        steps = d.get("_sar_train_steps", 0)
        d["_sar_train_steps"] = steps + 1
        if steps >= 1:                      # the first step built every cache this object indexes
            from .operand_refresh import OperandRefresh

            d["_sar_refresh"] = OperandRefresh(self.model)

    def generate(self, input_features: torch.Tensor, max_new_tokens: int = 256, num_beams: int = 1,
                 language: Optional[str] = None, task: Optional[str] = None, **kwargs) -> torch.Tensor:
        encoder = self.model.base_model.model.model.encoder
        was_checkpointing = encoder.gradient_checkpointing
        if was_checkpointing:
            self.model.base_model.model.gradient_checkpointing_disable()
            self.model.config.use_cache = True
        try:
            # short-form greedy decoding runs as one CUDA graph per token step (decode.py); beam search, sampling,
            # timestamps ... keep HF's loop.  Like the reference, `language` / `task` are accepted and not forwarded.
            from .decode import greedy_decoder_for, plan_greedy

            hf = self.model.base_model.model
            native = greedy_decoder_for(hf)
            call = dict(kwargs, max_new_tokens=max_new_tokens, num_beams=num_beams)
            plan = plan_greedy(hf, input_features, call) if native.supported() else None
            if plan is not None:
                return native.generate(input_features, plan)
            return self.model.generate(input_features=input_features, max_new_tokens=max_new_tokens,
                                       num_beams=num_beams, **kwargs)
        finally:
            if was_checkpointing:
                self.model.base_model.model.gradient_checkpointing_enable()
                self.model.config.use_cache = False

    def decode(self, token_ids: torch.Tensor, skip_special_tokens: bool = True) -> List[str]:
        if self.processor is None:
            raise RuntimeError("no WhisperProcessor available (offline); cannot decode token ids to text")
        return self.processor.batch_decode(token_ids, skip_special_tokens=skip_special_tokens)

    def save_adapter(self, save_path: Union[str, Path]) -> None:
        save_path = Path(save_path)
        save_path.mkdir(parents=True, exist_ok=True)
        self.model.save_pretrained(save_path)
        logger.info("Saved adapter to %s", save_path)

    def load_adapter(self, adapter_path: Union[str, Path]) -> None:
        self.model = PeftModel.from_pretrained(self.model.base_model, Path(adapter_path))
        logger.info("Loaded adapter from %s", adapter_path)

    def merge_and_unload(self):
        return self.model.merge_and_unload()

    def train(self, mode: bool = True):
        self.model.train(mode)
        return self

    def eval(self):
        self.model.eval()
        return self


def create_whisper_lora(model_name: str, lora_config: Optional[Dict] = None, language: Optional[str] = None,
                        **kwargs) -> WhisperLoRA:
    lora_config = lora_config or {}
    return WhisperLoRA(model_name=model_name, lora_r=lora_config.get("r", 16),
                       lora_alpha=lora_config.get("lora_alpha", 32),
                       lora_dropout=lora_config.get("lora_dropout", 0.1),
                       target_modules=lora_config.get("target_modules", ["q_proj", "v_proj"]),
                       language=language, **kwargs)


def load_whisper_lora_from_checkpoint(checkpoint_path: Union[str, Path], model_name: str,
                                      language: Optional[str] = None, device: Optional[str] = None,
                                      **kwargs) -> WhisperLoRA:
    checkpoint_path = Path(checkpoint_path)
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    base_model = load_base_model(model_name, device=device, **kwargs)
    model = PeftModel.from_pretrained(base_model, checkpoint_path)
    model.to(device)
    wrapper = WhisperLoRA.__new__(WhisperLoRA)
    nn.Module.__init__(wrapper)          # deviation from the reference, see module docstring
    wrapper.model = model
    wrapper.processor = _try_processor(model_name, language, "transcribe",
                                       offline=_random_init_requested(kwargs.get("random_init")))
    wrapper.model_name = model_name
    wrapper.language = language
    wrapper.task = "transcribe"
    wrapper.device = device
    wrapper.lora_config = model.peft_config[model.active_adapter]
    logger.info("Loaded WhisperLoRA from %s", checkpoint_path)
    return wrapper
