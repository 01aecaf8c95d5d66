"""CUDA-graph'd greedy decoding (decode.py) vs HF's own ``generate`` loop over the same model on the GPU: identical
token ids (integer work: bit-exact) wherever the two paths' bf16 logits do not tie — positions whose top-2 margin in
the HF path is below the bf16 noise floor are where argmax is legitimately ambiguous, and a row is compared up to the
first such position."""
import pytest
import torch

import speech_adapter_routing_b200 as sar
from speech_adapter_routing_b200 import decode, ops
from speech_adapter_routing_b200.routing import route

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny(cuda_dev):
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    cfg = WhisperConfig(vocab_size=51865, num_mel_bins=80, d_model=384, encoder_layers=2, decoder_layers=2,
                        encoder_attention_heads=6, decoder_attention_heads=6, encoder_ffn_dim=1536,
                        decoder_ffn_dim=1536, max_source_positions=1500, max_target_positions=448)
    torch.manual_seed(7)
    model = WhisperForConditionalGeneration(cfg).to(torch.bfloat16).to(cuda_dev).eval()
    model.config.forced_decoder_ids = None
    model.config.suppress_tokens = []
    model.generation_config.forced_decoder_ids = None
    model.generation_config.suppress_tokens = []
    langs = ["hindi", "italian", "punjabi"]
    for l in langs:
        sar.inject_lora(model, sar.LoraConfig(r=16, lora_alpha=32, target_modules=["q_proj", "v_proj"]), adapter_name=l)
    g = torch.Generator().manual_seed(3)
    for m in sar.lora_modules(model).values():
        for l in langs:
            m.lora_B[l].weight.data.copy_((torch.randn(m.out_features, 16, generator=g) * 0.05).to(cuda_dev))
    x = torch.randn(6, 80, 3000, generator=g).to(torch.bfloat16).to(cuda_dev)
    idx = torch.tensor([0, 2, -1, 1, 2, 0], dtype=torch.int32, device=cuda_dev)
    return model, x, idx


def hf_generate(model, x, idx, n_init, **kw):
    """HF's loop over the same model (K1 module slots, HF layer bodies with KV cache).  Returns the new tokens and the
    top-2 logit margin at every generated position (teacher-forced on HF's own tokens)."""
    with torch.no_grad(), route(idx):
        full = model.generate(input_features=x, return_dict_in_generate=True, **kw).sequences   # prompt + new tokens
        logits = model(input_features=x, decoder_input_ids=full[:, :-1], use_cache=False).logits.float()
    top2 = logits.topk(2, dim=-1).values
    margins = (top2[..., 0] - top2[..., 1])[:, n_init - 1:]
    return full[:, n_init:], margins


def compare(native, seqs, margins, floor=0.06):
    """rows equal up to the first position whose top-2 logit margin is below ``floor`` (bf16 logits of |v| ~ 4 carry
    ~0.03 of rounding noise; suppressed / forced positions are covered because both paths apply the same masks)."""
    B = seqs.shape[0]
    exact = 0
    for i in range(B):
        L = min(native.shape[1], seqs.shape[1])
        low = (margins[i, :L] < floor).nonzero()
        n_ok = int(low[0]) if len(low) else L
        assert torch.equal(native[i, :n_ok], seqs[i, :n_ok]), (i, native[i].tolist(), seqs[i].tolist())
        exact += int(native.shape[1] == seqs.shape[1] and torch.equal(native[i], seqs[i]))
    return exact


def test_native_greedy_equals_hf_generate_default_config(tiny):
    model, x, idx = tiny
    plan = decode.plan_greedy(model, x, {"max_new_tokens": 12, "num_beams": 1, "do_sample": False})
    assert plan is not None and plan.init_tokens == [model.config.decoder_start_token_id]
    assert plan.begin_suppress == [220, 50256]
    ops.reset_counters()
    native = decode.greedy_decoder_for(model).generate(x, plan, idx)
    seqs, scores = hf_generate(model, x, idx, 1, max_new_tokens=12, num_beams=1, do_sample=False)
    assert native.dtype == torch.long and native.shape[0] == 6
    assert compare(native, seqs, scores) >= 4
    # second call re-uses the captured graph (same shapes, same weights)
    again = decode.greedy_decoder_for(model).generate(x, plan, idx)
    assert torch.equal(native, again)


def test_native_greedy_language_detection_and_prompt_tokens(tiny):
    """A multilingual generation config: HF detects the language per utterance from the logits after
    <|startoftranscript|> and appends <|notimestamps|>; with an explicit language + task the prompt is fixed."""
    model, x, idx = tiny
    gc = model.generation_config
    gc.lang_to_id = {"<|en|>": 50259, "<|hi|>": 50276, "<|it|>": 50274, "<|pa|>": 50321, "<|te|>": 50299}
    gc.task_to_id = {"transcribe": 50359, "translate": 50358}
    gc.no_timestamps_token_id = 50363
    gc.is_multilingual = True
    try:
        plan = decode.plan_greedy(model, x, {"max_new_tokens": 8})
        assert plan.init_tokens == [model.config.decoder_start_token_id, decode.DETECT, 50363]
        native = decode.greedy_decoder_for(model).generate(x, plan, idx)
        seqs, scores = hf_generate(model, x, idx, 3, max_new_tokens=8)
        assert compare(native, seqs, scores) >= 4
        plan = decode.plan_greedy(model, x, {"max_new_tokens": 8, "language": "hindi", "task": "transcribe"})
        assert plan.init_tokens == [model.config.decoder_start_token_id, 50276, 50359, 50363]
        native = decode.greedy_decoder_for(model).generate(x, plan, idx)
        seqs, scores = hf_generate(model, x, idx, 4, max_new_tokens=8, language="hindi", task="transcribe")
        assert compare(native, seqs, scores) >= 4
    finally:
        for k in ("lang_to_id", "task_to_id", "no_timestamps_token_id", "is_multilingual"):
            delattr(gc, k)


def test_native_greedy_eos_padding_and_early_stop(tiny):
    """Make a frequently generated token the EOS: rows stop at different steps; later positions are pad, the EOS itself
    is kept, and the batch is cut where the last row finished — like HF."""
    model, x, idx = tiny
    plan0 = decode.plan_greedy(model, x, {"max_new_tokens": 24})
    free = decode.greedy_decoder_for(model).generate(x, plan0, idx)
    vals, counts = free[:, 2:].flatten().unique(return_counts=True)
    eos = int(vals[counts.argmax()])
    kw = {"max_new_tokens": 24, "eos_token_id": eos, "pad_token_id": 11}
    plan = decode.plan_greedy(model, x, kw)
    assert plan.eos_ids == [eos] and plan.pad_id == 11
    native = decode.greedy_decoder_for(model).generate(x, plan, idx)
    seqs, scores = hf_generate(model, x, idx, 1, **kw)
    compare(native, seqs, scores)
    for i in range(native.shape[0]):
        row = native[i].tolist()
        if eos in row:
            k = row.index(eos)
            assert all(t == 11 for t in row[k + 1:])
    assert native.shape[1] <= 24


def test_unsupported_requests_keep_hf_loop(tiny):
    model, x, idx = tiny
    assert decode.plan_greedy(model, x, {"max_new_tokens": 4, "num_beams": 2}) is None
    assert decode.plan_greedy(model, x, {"max_new_tokens": 4, "do_sample": True}) is None
    assert decode.plan_greedy(model, x, {"max_new_tokens": 4, "return_timestamps": True}) is None
    assert decode.plan_greedy(model, x, {"max_new_tokens": 4, "logits_processor": []}) is None
    assert decode.plan_greedy(model, x, {}) is None
    assert decode.plan_greedy(model, x.cpu(), {"max_new_tokens": 4}) is None
