"""CUDA-graph'd greedy decoding for routed Whisper.

The reference generates per utterance with ``adapter.generate(input_features[i:i+1])`` (src/models/adapter_router.py:744-750,
src/models/whisper_lora.py:145-186): one HF ``generate`` loop per clip, ~45 eager launches per decoder layer per token.
Here the whole mixed-language batch is decoded together; one token step of the 12/24/32-layer decoder is ONE CUDA
graph of libsar launches (fused q|k|v with the per-utterance adapter, epilogue-fused residual / GELU, LayerNorm, padded
lm head) over a static self-attention KV cache and cross-attention K/V computed once per batch.  Everything a step
needs to know about "where we are" (the position, the cache slot, the causal mask) is read from a device scalar, so
the same graph is replayed for every token without host synchronisation; finished rows are padded on the device and
the host looks at the stop condition every 16 tokens only.

What is reproduced from HF's Whisper ``generate`` (transformers 5.5.0, models/whisper/generation_whisper.py) for the
short-form greedy case — anything else is delegated to HF's own loop over the same kernels:
  * the prompt: ``decoder_start_token_id`` [+ language, task, <|notimestamps|>] (``_retrieve_init_tokens`` :1455-1607);
  * ``suppress_tokens`` at every step and ``begin_suppress_tokens`` at the first generated position (:1774-1806);
  * EOS: the EOS token is kept, later positions of that row are ``pad_token_id``, generation stops when every row has
    finished or after ``max_new_tokens``; the returned tensor holds the NEW tokens only, like HF's short-form path.
"""
from __future__ import annotations

from typing import Dict, List, NamedTuple, Optional

import torch

from . import ops
from ._lib import SAR_ACT_GELU
from .whisper_blocks import FUSED_BLOCKS_ENABLED, _dense, _ln  # noqa: F401  (same operand packs as the layer bodies)

_SUPPORTED_KWARGS = {"max_new_tokens", "num_beams", "do_sample", "language", "task", "return_timestamps",
                     "eos_token_id", "pad_token_id", "suppress_tokens", "begin_suppress_tokens", "use_cache"}


DETECT = -1   # placeholder in GreedyPlan.init_tokens: per-utterance language token chosen from the first step's logits


class GreedyPlan(NamedTuple):
    init_tokens: List[int]        # DETECT marks HF's automatic language detection (generation_whisper.py:1558-1565)
    lang_ids: List[int]           # candidate language token ids for DETECT
    max_new_tokens: int
    eos_ids: List[int]
    pad_id: int
    suppress: List[int]
    begin_suppress: List[int]


def _language_token(gc, language: str) -> Optional[int]:
    from transformers.models.whisper.tokenization_whisper import TO_LANGUAGE_CODE

    language = language.lower()
    if language in gc.lang_to_id:
        tok = language
    elif language in TO_LANGUAGE_CODE:
        tok = f"<|{TO_LANGUAGE_CODE[language]}|>"
    elif language in TO_LANGUAGE_CODE.values():
        tok = f"<|{language}|>"
    else:
        return None
    return gc.lang_to_id.get(tok)


LAST_FALLBACK_REASON = ""   # why the most recent plan_greedy() call declined (diagnostics)


def _no(reason: str) -> None:
    global LAST_FALLBACK_REASON
    LAST_FALLBACK_REASON = reason
    return None


def plan_greedy(model, input_features: torch.Tensor, kwargs: Dict) -> Optional[GreedyPlan]:
    """The native plan for this ``generate`` call, or None when it needs anything beyond short-form greedy decoding
    (beam search, sampling, timestamps, prompts, language detection, custom processors ...)."""
    if not FUSED_BLOCKS_ENABLED or not input_features.is_cuda or input_features.dim() != 3:
        return _no("fused blocks disabled or input not a CUDA [B, mel, frames] tensor")
    if any(k not in _SUPPORTED_KWARGS for k in kwargs):
        return _no("unsupported generate() keyword")
    gc = model.generation_config
    cfg = model.config
    def get(k, d=None):   # call kwarg > generation config > default (transformers 5 leaves unset fields at None)
        v = kwargs.get(k)
        if v is None:
            v = getattr(gc, k, None)
        return d if v is None else v

    if get("num_beams", 1) != 1 or get("do_sample", False) or get("return_timestamps", None):
        return _no("beam search / sampling / timestamps")
    if kwargs.get("max_new_tokens") is None:
        return _no("max_new_tokens not given")
    if input_features.shape[-1] != cfg.max_source_positions * 2:
        return _no("long-form input")
    for name in ("forced_decoder_ids",):
        if getattr(gc, name, None) is not None or getattr(cfg, name, None) is not None:
            return _no("forced_decoder_ids set")
    for name in ("no_speech_threshold", "logprob_threshold", "compression_ratio_threshold", "temperature",
                 "repetition_penalty", "no_repeat_ngram_size", "bad_words_ids", "min_length", "min_new_tokens",
                 "prompt_ids", "max_time", "stop_strings", "encoder_repetition_penalty", "sequence_bias",
                 "exponential_decay_length_penalty", "forced_bos_token_id", "forced_eos_token_id", "guidance_scale"):
        v = getattr(gc, name, None)
        if v is None or v is False or v == 0 or (name in ("temperature", "repetition_penalty",
                                                           "encoder_repetition_penalty") and v == 1.0):
            continue
        return _no(f"generation option outside greedy decoding: {name}={v!r}")
    start = getattr(gc, "decoder_start_token_id", None)
    if start is None:
        start = cfg.decoder_start_token_id
    init = [int(start)]
    language, task = kwargs.get("language"), kwargs.get("task")
    language = language if language is not None else getattr(gc, "language", None)
    task = task if task is not None else getattr(gc, "task", None)
    if language is not None:
        if isinstance(language, (list, tuple)) or not hasattr(gc, "lang_to_id"):
            return _no("list of languages or no lang_to_id")
        lid = _language_token(gc, language)
        if lid is None:
            return _no("unknown language")
        init.append(int(lid))
    elif hasattr(gc, "lang_to_id"):
        init.append(DETECT)   # HF detects the language of every utterance from the logits after <|startoftranscript|>
    if task is not None:
        if not hasattr(gc, "task_to_id") or task not in gc.task_to_id:
            return _no("task not in task_to_id")
        init.append(int(gc.task_to_id[task]))
    elif language is not None and hasattr(gc, "task_to_id"):
        init.append(int(gc.task_to_id["transcribe"]))
    nots = getattr(gc, "no_timestamps_token_id", None)
    if nots is not None and init[-1] != nots:
        init.append(int(nots))
    eos = get("eos_token_id", cfg.eos_token_id)
    eos_ids = [int(e) for e in (eos if isinstance(eos, (list, tuple)) else [eos])] if eos is not None else []
    pad = get("pad_token_id", cfg.pad_token_id)
    if pad is None:
        pad = eos_ids[0] if eos_ids else 0
    max_new = int(kwargs["max_new_tokens"])
    max_new = min(max_new, cfg.max_target_positions - len(init))
    if max_new <= 0:
        return _no("no room for new tokens")
    V = cfg.vocab_size
    sup = [int(t) for t in (get("suppress_tokens", None) or []) if 0 <= int(t) < V]
    bsup = [int(t) for t in (get("begin_suppress_tokens", None) or []) if 0 <= int(t) < V]
    lang_ids = sorted(int(v) for v in getattr(gc, "lang_to_id", {}).values()) if DETECT in init else []
    return GreedyPlan(init, lang_ids, max_new, eos_ids, int(pad), sup, bsup)


class _State:
    pass


class StaticGreedyDecoder:
    """One per WhisperForConditionalGeneration; caches one captured step graph per (batch, cache length, weights)."""

    def __init__(self, model):
        self.model = model
        self._states: Dict[tuple, _State] = {}

    # ------------------------------------------------------------------ support / keys
    def supported(self) -> bool:
        dec = self.model.model.decoder
        return (hasattr(dec, "_sar_pack") and all(hasattr(l, "_sar_pack") for l in dec.layers)
                and hasattr(self.model.proj_out, "_sar_hf_forward")
                and dec.embed_tokens.weight.dtype == torch.bfloat16 and dec.embed_tokens.weight.is_cuda)

    def _weights_key(self) -> int:
        dec = self.model.model.decoder
        from .routing import operand_epoch

        return hash(tuple((p.data_ptr(), p._version) for p in dec.parameters()) + (operand_epoch(),))

    # ------------------------------------------------------------------ one token step (graph-capturable)
    def _step(self, st: _State) -> None:
        model = self.model
        dec = model.model.decoder
        B = st.tok.shape[0]
        h = dec.embed_tokens(st.tok).unsqueeze(1) + dec.embed_positions.weight.index_select(0, st.pos).unsqueeze(0)
        idx = st.idx if st.use_idx else None
        for l, layer in enumerate(dec.layers):
            pk = layer._sar_pack
            qkv, cq = pk["self"].qkv.get(), pk["cross"].q.get()
            x = _ln(h, pk["ln1"], layer.self_attn_layer_norm.eps)
            q, k, v = qkv(x, idx if qkv.lora_mods else None)
            # cache write + masked softmax(q kᵀ) v in one launch, position read on the device
            o = ops.decode_self_attn(q, k, v, st.K[l], st.V[l], st.pos)
            h = _dense(o.view(B, 1, -1), pk["self"].out, residual=h)
            x = _ln(h, pk["ln2"], layer.encoder_attn_layer_norm.eps)
            (q,) = cq(x, idx if cq.lora_mods else None)
            o = ops.decode_cross_attn(q, st.CK[l], st.CV[l])      # streams the encoder K / V of every (b, h) once
            h = _dense(o.view(B, 1, -1), pk["cross"].out, residual=h, inplace=True)
            x = _ln(h, pk["ln3"], layer.final_layer_norm.eps)
            f = _dense(x, pk["fc1"], act=SAR_ACT_GELU)
            h = _dense(f, pk["fc2"], residual=h, inplace=True)
        ln = dec._sar_pack["ln"].get()
        h = ops.layernorm_fwd(h, ln.W, ln.b, dec.layer_norm.eps)
        st.logits = model.proj_out(h).view(B, -1)     # [B, V] view of the padded logits buffer (libsar lm head)
        st.pos.add_(1)

    def _get_state(self, B: int, Tmax: int, dev, use_idx: bool) -> _State:
        key = (B, Tmax, str(dev), use_idx, self._weights_key())
        st = self._states.get(key)
        if st is not None:
            return st
        self._states.clear()                           # one live graph (weights or shape changed): free the old pools
        cfg = self.model.config
        L, H, d = cfg.decoder_layers, cfg.decoder_attention_heads, cfg.d_model
        S = cfg.max_source_positions
        st = _State()
        st.use_idx = use_idx
        st.tok = torch.zeros(B, dtype=torch.long, device=dev)
        st.pos = torch.zeros(1, dtype=torch.long, device=dev)
        st.idx = torch.full((B,), -1, dtype=torch.int32, device=dev)
        st.arange = torch.arange(Tmax, device=dev)
        z = lambda T: torch.zeros(B, H, T, d // H, dtype=torch.bfloat16, device=dev)
        st.K = [z(Tmax) for _ in range(L)]
        st.V = [z(Tmax) for _ in range(L)]
        st.CK = [z(S) for _ in range(L)]
        st.CV = [z(S) for _ in range(L)]
        st.logits = None
        # warm up on a side stream (first launches set kernel attributes and build the operand packs), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._step(st)
            self._step(st)
        torch.cuda.current_stream(dev).wait_stream(side)
        st.pos.zero_()
        st.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st.graph):
            self._step(st)
        st.pos.zero_()
        self._states[key] = st
        return st

    # ------------------------------------------------------------------ public entry
    @torch.no_grad()
    def generate(self, input_features: torch.Tensor, plan: GreedyPlan,
                 utt_adapter: Optional[torch.Tensor] = None) -> torch.Tensor:
        """New tokens [B, <= max_new_tokens] (int64).  ``utt_adapter``: int32 [B] per-utterance adapter index, or None
        to use every LoRA module's own default (active adapter / routing context)."""
        from .routing import route

        model = self.model
        dec = model.model.decoder
        dev = input_features.device
        B = input_features.shape[0]
        x = input_features if input_features.dtype == torch.bfloat16 else input_features.to(torch.bfloat16)
        first_lora = next((m for l in dec.layers for m in l._sar_pack["self"].qkv.get().lora_mods), None)
        if utt_adapter is None and first_lora is not None:
            utt_adapter = first_lora.resolve_index(B, dev)
        use_idx = utt_adapter is not None
        n_init = len(plan.init_tokens)
        Tmax = ((n_init + plan.max_new_tokens + 63) // 64) * 64
        st = self._get_state(B, Tmax, dev, use_idx)
        if use_idx:
            st.idx.copy_(utt_adapter.to(torch.int32))
        with route(utt_adapter) if use_idx else _null():
            enc = model.model.encoder(x).last_hidden_state
        for l, layer in enumerate(dec.layers):          # cross-attention K/V: once per batch, into the static buffers
            ckv = layer._sar_pack["cross"].kv.get()
            k, v = ckv(enc, st.idx if (use_idx and ckv.lora_mods) else None)
            st.CK[l].copy_(k)
            st.CV[l].copy_(v)
        st.pos.zero_()
        V = model.config.vocab_size
        neg = float("-inf")
        sup = torch.tensor(plan.suppress, dtype=torch.long, device=dev) if plan.suppress else None
        bsup = torch.tensor(plan.begin_suppress, dtype=torch.long, device=dev) if plan.begin_suppress else None
        eos = torch.tensor(plan.eos_ids, dtype=torch.long, device=dev) if plan.eos_ids else None
        out = torch.full((B, plan.max_new_tokens), plan.pad_id, dtype=torch.long, device=dev)
        finished = torch.zeros(B, dtype=torch.bool, device=dev)
        init = torch.tensor(plan.init_tokens, dtype=torch.long, device=dev)
        langs = torch.tensor(plan.lang_ids, dtype=torch.long, device=dev) if plan.lang_ids else None
        st.tok.copy_(init[0].expand(B))
        n_done = 0
        for step in range(n_init + plan.max_new_tokens - 1):
            st.graph.replay()
            if step < n_init - 1:
                if plan.init_tokens[step + 1] == DETECT:   # argmax over the language tokens only (detect_language)
                    st.tok.copy_(langs[st.logits[:, :V].float().index_select(1, langs).argmax(dim=-1)])
                else:
                    st.tok.copy_(init[step + 1].expand(B))
                continue
            lg = st.logits[:, :V].float()
            if sup is not None:
                lg.index_fill_(1, sup, neg)
            if bsup is not None and n_done == 0:
                lg.index_fill_(1, bsup, neg)
            nxt = lg.argmax(dim=-1)
            nxt = torch.where(finished, torch.full_like(nxt, plan.pad_id), nxt)
            out[:, n_done] = nxt
            if eos is not None:
                finished |= torch.isin(nxt, eos)
            st.tok.copy_(nxt)
            n_done += 1
            if eos is not None and n_done % 16 == 0 and bool(finished.all()):
                break
        out = out[:, :n_done]
        if eos is not None and n_done > 0:
            # HF stops as soon as every row has finished: cut the columns decoded past that point
            is_eos = torch.isin(out, eos)
            first = torch.where(is_eos.any(1), is_eos.float().argmax(1) + 1, torch.full((B,), n_done, device=dev))
            out = out[:, : int(first.max().item())]
        return out


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def greedy_decoder_for(model) -> StaticGreedyDecoder:
    dec = getattr(model, "_sar_greedy_decoder", None)
    if dec is None:
        dec = StaticGreedyDecoder(model)
        object.__setattr__(model, "_sar_greedy_decoder", dec)
    return dec
