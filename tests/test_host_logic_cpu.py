"""Host-side mirror of the reference API: module injection, PEFT checkpoint layout, routing context, router
checkpoint format, aggregation semantics.  No GPU and no kernel calls."""
import json
from pathlib import Path

import pytest
import torch
import torch.nn as nn

import speech_adapter_routing_b200 as sar
from oracle import fixtures, router as orouter, whisper as owhisper
from speech_adapter_routing_b200 import lid_router, routing
from speech_adapter_routing_b200.peft_compat import ADAPTER_CONFIG, ADAPTER_WEIGHTS


@pytest.fixture(scope="module")
def micro_model():
    return owhisper.build_whisper("micro")


def test_injection_targets_every_q_and_v_proj(micro_model):
    import copy
    m = copy.deepcopy(micro_model)
    paths = sar.inject_lora(m, sar.LoraConfig(r=16, lora_alpha=32, target_modules=["q_proj", "v_proj"]))
    L = m.config.encoder_layers
    assert len(paths) == 2 * L + 4 * L          # enc self-attn q,v + dec self-attn q,v + dec cross-attn q,v
    assert all(p.endswith(("q_proj", "v_proj")) for p in paths)
    assert not any("k_proj" in p or "out_proj" in p for p in paths)
    mod = m.get_submodule(paths[0])
    assert isinstance(mod, sar.RoutedLoRALinear)
    # PEFT defaults: A kaiming-uniform, B zero; adapters fp32; base frozen
    assert torch.count_nonzero(mod.lora_B["default"].weight) == 0
    assert mod.lora_A["default"].weight.dtype == torch.float32
    assert not mod.base_layer.weight.requires_grad and mod.lora_A["default"].weight.requires_grad


def test_peft_model_attribute_paths_and_trainable_set(micro_model):
    import copy
    pm = sar.get_peft_model(copy.deepcopy(micro_model), sar.LoraConfig(r=8, lora_alpha=16,
                                                                        target_modules=["q_proj", "v_proj"]))
    # paths the reference walks: whisper_lora.py:168-184 and adapter_router.py:425-437
    assert pm.base_model.model.model.encoder.gradient_checkpointing in (True, False)
    pm.base_model.model.gradient_checkpointing_enable()
    assert pm.base_model.model.model.encoder.gradient_checkpointing
    pm.base_model.model.gradient_checkpointing_disable()
    assert pm.config.use_cache in (True, False)
    names = [n for n, p in pm.named_parameters() if p.requires_grad]
    assert names and all(".lora_A." in n or ".lora_B." in n for n in names)
    # trainer's weight-decay grouping (trainer.py:112-118) relies on these substrings being absent
    assert not any(s in n for n in names for s in ("bias", "LayerNorm", "layer_norm"))
    d, r, L = micro_model.config.d_model, 8, micro_model.config.encoder_layers
    assert sum(p.numel() for p in pm.parameters() if p.requires_grad) == 6 * L * 2 * r * d


def test_save_pretrained_writes_peft_layout_and_round_trips(tmp_path, micro_model):
    import copy
    from safetensors.torch import load_file

    pm = sar.get_peft_model(copy.deepcopy(micro_model), sar.LoraConfig(r=16, lora_alpha=32,
                                                                        target_modules=["q_proj", "v_proj"]))
    with torch.no_grad():
        for m in sar.lora_modules(pm).values():
            m.lora_B["default"].weight.normal_(0, 0.02)
    pm.save_pretrained(tmp_path / "ad")
    assert sorted(p.name for p in (tmp_path / "ad").iterdir()) == [ADAPTER_CONFIG, ADAPTER_WEIGHTS]
    cfg = json.loads((tmp_path / "ad" / ADAPTER_CONFIG).read_text())
    for k, v in {"peft_type": "LORA", "task_type": None, "r": 16, "lora_alpha": 32, "bias": "none",
                 "fan_in_fan_out": False, "modules_to_save": None, "use_rslora": False, "use_dora": False}.items():
        assert cfg[k] == v
    assert sorted(cfg["target_modules"]) == ["q_proj", "v_proj"]
    sd = load_file(str(tmp_path / "ad" / ADAPTER_WEIGHTS))
    key = "base_model.model.model.encoder.layers.0.self_attn.q_proj.lora_A.weight"   # adapter name elided
    assert key in sd and sd[key].shape == (16, micro_model.config.d_model)
    assert all(k.startswith("base_model.model.model.") and k.endswith((".lora_A.weight", ".lora_B.weight")) for k in sd)
    # reload into a fresh base model
    pm2 = sar.PeftModel.from_pretrained(copy.deepcopy(micro_model), tmp_path / "ad")
    a, b = pm.state_dict(), pm2.state_dict()
    assert a.keys() == b.keys()
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_reader_accepts_an_independently_written_peft_adapter(tmp_path, micro_model):
    import copy
    weights = owhisper.make_adapter_weights(micro_model, 16, 3)
    owhisper.write_peft_adapter(tmp_path / "x", weights, k=2, r=16, lora_alpha=32)
    pm = sar.PeftModel.from_pretrained(copy.deepcopy(micro_model), tmp_path / "x")
    for path, (A, B) in weights.items():
        m = pm.base_model.model.get_submodule(path)
        assert torch.equal(m.lora_A["default"].weight, A[2]) and torch.equal(m.lora_B["default"].weight, B[2])
        assert m.scaling["default"] == 2.0


def test_merge_and_unload_folds_the_adapter(micro_model):
    import copy
    pm = sar.get_peft_model(copy.deepcopy(micro_model), sar.LoraConfig(r=4, lora_alpha=8, target_modules=["q_proj"]))
    m = next(iter(sar.lora_modules(pm).values()))
    with torch.no_grad():
        m.lora_B["default"].weight.normal_(0, 0.02)
    W = m.base_layer.weight + 2.0 * m.lora_B["default"].weight @ m.lora_A["default"].weight
    merged = pm.merge_and_unload()
    assert not sar.lora_modules(merged)
    q = merged.model.encoder.layers[0].self_attn.q_proj
    assert isinstance(q, nn.Linear) and torch.allclose(q.weight, W, atol=1e-6)


def test_routed_linear_has_no_cpu_fallback():
    lin = nn.Linear(64, 64)
    m = sar.RoutedLoRALinear(lin, "a", r=16, lora_alpha=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 64))


def test_routing_context_is_scoped_and_nested():
    assert routing.current_utt_adapter() is None
    a = torch.tensor([0, 1], dtype=torch.int64)
    with sar.route(a):
        cur = routing.current_utt_adapter()
        assert cur.dtype == torch.int32 and cur.tolist() == [0, 1]
        with sar.route(None):
            assert routing.current_utt_adapter() is None
        assert routing.current_utt_adapter().tolist() == [0, 1]
    assert routing.current_utt_adapter() is None
    assert sar.base_only(3, "cpu").tolist() == [-1, -1, -1]


def test_language_classifier_checkpoint_format_and_state_dict_keys(tmp_path):
    clf = sar.LanguageClassifier(input_dim=64, num_classes=4, languages=["hi", "it", "pa", "te"],
                                 class_weights=[1.0, 2.0, 1.0, 0.5])
    # with class weights the reference's state dict also carries the buffer and CrossEntropyLoss's copy of it
    assert set(clf.state_dict()) == set(fixtures.ROUTER_KEYS) | {"_class_weights", "loss_fn.weight"}
    assert set(sar.LanguageClassifier(input_dim=64, num_classes=4).state_dict()) == set(fixtures.ROUTER_KEYS)
    clf.save(tmp_path / "c" / "classifier.pt")
    ck = torch.load(tmp_path / "c" / "classifier.pt", weights_only=True)
    assert set(ck) == {"state_dict", "config"}
    assert set(ck["config"]) == {"input_dim", "num_classes", "pooling", "use_cnn", "label_smoothing", "languages",
                                 "class_weights"}
    clf2 = sar.LanguageClassifier.load(tmp_path / "c" / "classifier.pt")
    assert clf2.languages == ["hi", "it", "pa", "te"]
    assert all(torch.equal(v, clf2.state_dict()[k]) for k, v in clf.state_dict().items())


def test_language_classifier_torch_graph_matches_reference_golden():
    """The training/CPU graph of the mirror class reproduces the REFERENCE class's outputs on its golden vectors."""
    g = torch.load(Path(__file__).parent / "golden" / "router_golden.pt")
    for c in g["cases"]:
        clf = sar.LanguageClassifier(input_dim=c["d"], num_classes=c["C"]).eval()
        clf.load_state_dict(c["state_dict"])
        with torch.no_grad():
            out = clf(c["h"].float())
        assert torch.allclose(out["logits"], c["logits"], atol=1e-6)
        labels, probs = clf.predict(c["h"].float())
        assert torch.equal(labels, c["labels"]) and torch.allclose(probs, c["predict_probs"], atol=1e-7)
        names, _ = clf.predict_language(c["h"].float())
        assert names == [clf.idx_to_lang[i] for i in c["labels"].tolist()]


def test_class_weight_strategies():
    w = sar.LanguageClassifier.compute_class_weights_from_counts({"a": 100, "b": 10}, ["a", "b"])
    assert torch.allclose(w, torch.tensor([0.55 / 3.025, 5.5 / 3.025]), atol=1e-4)
    w = sar.LanguageClassifier.compute_class_weights_from_counts({"a": 100, "b": 1}, ["a", "b"], max_weight=1.5)
    # [0.505, 50.5] -> /mean -> [0.0198, 1.9802] -> clamp 1.5 -> /mean -> [0.02606, 1.97394]
    assert torch.allclose(w, torch.tensor([0.02606, 1.97394]), atol=1e-4)
    with pytest.raises(ValueError):
        sar.LanguageClassifier.compute_class_weights_from_counts({"a": 1}, ["a"], strategy="nope")


def test_per_utterance_loss_is_mean_of_per_sample_means():
    """Reference aggregation (adapter_router.py:707) — not HF's batch token-mean."""
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3, 5, 11, generator=g)
    labels = torch.randint(0, 11, (3, 5), generator=g)
    labels[1, 3:] = -100
    want = torch.stack([torch.nn.functional.cross_entropy(logits[i], labels[i], ignore_index=-100)
                        for i in range(3)]).mean()
    assert torch.allclose(lid_router._per_utterance_loss(logits, labels), want, atol=1e-6)


def test_cut_at_first_stop_matches_the_reference_runs_generate_padding():
    """Pinned by tests/golden/routed_forward_golden.pt ("generate_eos"): HF's Whisper generate drops a row's closing EOS,
    the reference's loop pads shorter rows with 0 (adapter_router.py:753-761)."""
    eos = 3
    ids = torch.tensor([[4, 9, 8, 3, 1, 1], [4, 7, 3, 1, 1, 1], [4, 5, 6, 7, 8, 9]])
    out = lid_router._cut_at_first_stop(ids, [eos], fill=0)
    assert out.tolist() == [[4, 9, 8, 0, 0, 0], [4, 7, 0, 0, 0, 0], [4, 5, 6, 7, 8, 9]]
    ids = torch.tensor([[4, 9, 3, 1], [4, 3, 1, 1]])
    assert lid_router._cut_at_first_stop(ids, eos).tolist() == [[4, 9], [4, 0]]        # width = longest row before EOS
    assert lid_router._cut_at_first_stop(ids, [eos], fill=1).tolist() == [[4, 9], [4, 1]]  # generate(language=): HF's pad
    assert lid_router._cut_at_first_stop(ids, None) is ids
    # the golden's own EOS case, replayed on the raw (EOS kept, then pad) rows a batched decoder produces
    from golden_cases import load_golden
    rec = load_golden()["cases"][0]
    e, free = rec["generate_eos"], rec["generate"]["ids"]
    raw = free.clone()
    hit = (raw == e["eos_token_id"]).cumsum(1) > 1          # after the first EOS: anything (here: pad id 1)
    raw[hit] = 1
    assert torch.equal(lid_router._cut_at_first_stop(raw, [e["eos_token_id"]], fill=0), e["ids"])


def test_whisper_lora_wrapper_api_offline(monkeypatch):
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    w = sar.WhisperLoRA("whisper-tiny", lora_r=8, lora_alpha=16, lora_dropout=0.0, language="hindi", device="cpu")
    assert w.model_name == "openai/whisper-tiny" and w.language == "hindi" and w.task == "transcribe"
    assert w.lora_config.r == 8 and w.lora_config.target_modules == ["q_proj", "v_proj"]
    assert w.model.base_model.model.model.encoder.gradient_checkpointing      # default use_gradient_checkpointing
    assert w.model.config.use_cache is False
    assert w.train() is w and w.eval() is w
    fx = sar.EncoderFeatureExtractor(w)
    assert fx.get_hidden_dim() == 384
    assert fx._get_encoder() is w.model.base_model.model.model.encoder
    w2 = sar.create_whisper_lora("whisper-tiny", {"r": 4, "lora_alpha": 8, "lora_dropout": 0.0}, device="cpu")
    assert w2.lora_config.r == 4
    assert sar.get_model_name("whisper-large") == "openai/whisper-large-v3"
    assert sar.get_model_info("whisper-small")["hidden_size"] == 768


def test_attention_dispatch_picks_the_own_kernel_where_it_measured_faster(monkeypatch):
    """whisper_blocks._sdpa: libsar's attention kernel for one query tile per head (<= 128 rows) and for causal
    self-attention up to 512 rows, only without an explicit mask; everything else — longer cross-attention, the
    1500 x 1500 encoder attention — goes to torch SDPA."""
    from speech_adapter_routing_b200 import ops, whisper_blocks as wb

    calls = []
    monkeypatch.setattr(ops, "attn_fwd", lambda q, k, v, causal: calls.append(("own", bool(causal))) or q)
    monkeypatch.setattr(wb.F, "scaled_dot_product_attention",
                        lambda q, k, v, attn_mask=None, is_causal=False, scale=None: calls.append(("torch", bool(is_causal))) or q)
    monkeypatch.setattr(wb, "OWN_ATTN_MAX_TQ", 128)
    monkeypatch.setattr(wb, "OWN_ATTN_MAX_TQ_CAUSAL", 512)
    t = lambda tq: torch.zeros(1, 2, tq, 64)
    wb._sdpa(t(128), t(1500), t(1500))                              # decoder cross-attention, T_dec = 128
    wb._sdpa(t(128), t(128), t(128), causal=True)                   # decoder self-attention
    wb._sdpa(t(448), t(448), t(448), causal=True)                   # ... at the maximum target length
    wb._sdpa(t(1), t(77), t(77), causal=True)                       # one query row: the causal mask is a no-op
    wb._sdpa(t(448), t(1500), t(1500))                              # long cross-attention
    wb._sdpa(t(1500), t(1500), t(1500))                             # encoder self-attention
    wb._sdpa(t(128), t(128), t(128), mask=torch.zeros(1, 1, 128, 128), causal=True)   # explicit mask
    assert calls == [("own", False), ("own", True), ("own", True), ("own", False), ("torch", False), ("torch", False),
                     ("torch", False)]
    monkeypatch.setattr(wb, "OWN_ATTN_MAX_TQ", 0)
    monkeypatch.setattr(wb, "OWN_ATTN_MAX_TQ_CAUSAL", 0)
    calls.clear()
    wb._sdpa(t(128), t(1500), t(1500))
    wb._sdpa(t(128), t(128), t(128), causal=True)
    assert calls == [("torch", False), ("torch", True)]


def test_logmel_host_tables_match_the_feature_extractor_and_padding_rules():
    """speech_adapter_routing_b200.logmel (host side of sar_logmel_fwd): its own Slaney filterbank equals
    WhisperFeatureExtractor.mel_filters, the window is the periodic Hann of torch.hann_window, and clips are zero-padded
    on the right / cut to 30 s like the feature extractor does."""
    import numpy as np
    from transformers import WhisperFeatureExtractor

    from speech_adapter_routing_b200 import logmel

    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        assert np.abs(logmel.mel_filterbank(n_mels).numpy() - fe.mel_filters).max() <= 1e-12
    window, cos_t, sin_t, filt = logmel.tables("cpu", 80)
    assert torch.allclose(window, torch.hann_window(400), atol=1e-7)
    j = torch.arange(400, dtype=torch.float64)
    assert torch.allclose(cos_t.double() ** 2 + sin_t.double() ** 2, torch.ones(400, dtype=torch.float64), atol=1e-6)
    assert filt.shape == (201, 80) and filt.dtype == torch.float32
    x = logmel.pad_or_trim([torch.ones(10), torch.ones(logmel.N_SAMPLES + 5)])
    assert x.shape == (2, logmel.N_SAMPLES) and x[0, :10].sum() == 10 and x[0, 10:].abs().sum() == 0 and x[1].min() == 1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        logmel.log_mel_spectrogram(torch.zeros(2, 16000))


def test_lora_dropout_correction_completes_pefts_formula():
    """lora_dropout > 0 in training mode (the reference's default 0.1): the fused kernels give base(x) + s·B(A(x)) and
    RoutedLoRALinear adds s·B(A(drop(x) - x)); the sum must be PEFT's base(x) + s·B(A(drop(x))) for the mask that was
    drawn, per utterance adapter, with gradients reaching A and B.  (Pure torch, so it is checked here on the CPU.)"""
    from speech_adapter_routing_b200.lora_linear import RoutedLoRALinear

    torch.manual_seed(0)
    base = nn.Linear(32, 48)
    m = RoutedLoRALinear(base, "hindi", r=4, lora_alpha=8, lora_dropout=0.25)
    m.add_adapter("telugu", 8, 16, 0.25)
    for name in ("hindi", "telugu"):
        nn.init.normal_(m.lora_B[name].weight, std=0.1)
    m.train()
    assert m._dropout_active()
    x = torch.randn(3, 5, 32)
    idx = torch.tensor([1, -1, 0], dtype=torch.int32)
    torch.manual_seed(123)
    corr = m._dropout_correction(x, idx)
    torch.manual_seed(123)
    xd = m.lora_dropout["hindi"](x)                      # the same mask
    assert (xd == 0).any() and not torch.equal(xd, x)
    for b, k in enumerate(idx.tolist()):
        if k < 0:
            assert corr[b].abs().max() == 0
            continue
        name = m.adapter_order[k]
        A, Bw, s_ = m.lora_A[name].weight, m.lora_B[name].weight, m.scaling[name]
        fused = base(x[b]) + s_ * (x[b] @ A.t()) @ Bw.t()            # what K1 computes
        peft = base(x[b]) + s_ * (xd[b] @ A.t()) @ Bw.t()            # PEFT lora.Linear.forward in training mode
        assert torch.allclose(fused + corr[b], peft, atol=1e-5)
    corr.sum().backward()
    assert m.lora_A["hindi"].weight.grad.abs().sum() > 0 and m.lora_B["telugu"].weight.grad.abs().sum() > 0
    m.eval()
    assert not (m.training and m._dropout_active())      # eval: no correction, the kernels' result is PEFT's


def test_route_batch_covers_the_non_default_lid_architectures_on_the_torch_graph():
    """LanguageClassifier.route_batch: the K2 kernel serves the default architecture on a GPU; max / attention pooling
    and the CNN front-end (reference src/models/adapter_router.py:210-249, 271-275) run the torch graph and still get
    idx / perm / seg_starts (stable sort by adapter index) without a host round trip."""
    torch.manual_seed(4)
    h = torch.randn(6, 40, 32)
    for kw in ({"pooling": "max"}, {"pooling": "attention"}, {"use_cnn": True, "cnn_channels": 16}):
        clf = sar.LanguageClassifier(input_dim=32, num_classes=3, languages=["a", "b", "c"], **kw).eval()
        out = clf.route_batch(h)
        labels, probs = clf.predict(h)
        assert torch.equal(out.idx.long(), labels) and torch.allclose(out.probs, probs.float())
        assert out.idx.dtype == torch.int32 and out.perm.dtype == torch.int32 and out.seg_starts.dtype == torch.int32
        exp_perm = sorted(range(6), key=lambda i: (int(labels[i]), i))
        assert out.perm.tolist() == exp_perm
        assert out.seg_starts.tolist() == [0] + torch.bincount(labels, minlength=3).cumsum(0).tolist()


def test_lora_on_fc1_fc2_keeps_hfs_layer_body(micro_model):
    """a LoRA target on fc1 / fc2 (scripts/train_lora.py:57 exposes --target_modules) must not bind the fused
    layer body, which reads fc1 / fc2 as plain dense operands and would drop (or crash on) the adapter term."""
    import copy
    from speech_adapter_routing_b200 import whisper_blocks as wb

    m = copy.deepcopy(micro_model)
    sar.inject_lora(m, sar.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj", "v_proj", "fc1", "fc2"]))
    wb.install_fused_blocks(m)
    for layer in list(m.model.encoder.layers) + list(m.model.decoder.layers):
        assert isinstance(layer.fc1, sar.RoutedLoRALinear) and isinstance(layer.fc2, sar.RoutedLoRALinear)
        assert not hasattr(layer, "_sar_pack")                       # HF's body stays
    m2 = copy.deepcopy(micro_model)
    sar.inject_lora(m2, sar.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj", "v_proj"]))
    wb.install_fused_blocks(m2)
    assert all(hasattr(l, "_sar_pack") for l in m2.model.encoder.layers)   # the default target set is bound


def test_refresh_operands_invalidates_every_cache_key():
    """writes through ``param.data`` do not bump ``_version``; ``refresh_operands()`` is the explicit hook and
    load_adapter / add_adapter call it."""
    from speech_adapter_routing_b200 import whisper_blocks as wb

    lin = nn.Linear(64, 64)
    m = sar.RoutedLoRALinear(lin, "a", r=8, lora_alpha=16)
    k0, p0 = m._key(), wb._pver(lin.weight)
    lin.weight.data.mul_(2.0)                                        # invisible to the version counter ...
    assert m._key() == k0 and wb._pver(lin.weight) == p0
    sar.refresh_operands()                                           # ... visible through the epoch
    assert m._key() != k0 and wb._pver(lin.weight) != p0
    k1 = m._key()
    with torch.no_grad():
        lin.weight.mul_(0.5)                                         # in-place op under no_grad: version bump, no hook needed
    assert m._key() != k1
    e = routing.operand_epoch()
    m.add_adapter("b", 8, 16)
    assert routing.operand_epoch() > e


def test_adapter_router_rejects_a_classifier_with_the_wrong_class_count(micro_model):
    import copy
    m = copy.deepcopy(micro_model)
    langs = ["hindi", "italian"]
    for l in langs:
        sar.inject_lora(m, sar.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj", "v_proj"]), adapter_name=l)
    clf = sar.LanguageClassifier(input_dim=m.config.d_model, num_classes=3, languages=["a", "b", "c"])
    with pytest.raises(ValueError, match="3 classes"):
        sar.AdapterRouter.from_stacked(m, clf, langs)
    ok = sar.LanguageClassifier(input_dim=m.config.d_model, num_classes=2, languages=langs)
    assert sar.AdapterRouter.from_stacked(m, ok, langs).languages == langs


def test_route_mix_context_and_merged_adapter_layout(micro_model):
    """soft_fused plumbing on the host: the routing context carries fp32 weights and one merged adapter index; the
    merged stacks put adapter g's rank rows / columns at [g*r, (g+1)*r)."""
    import copy
    from speech_adapter_routing_b200 import whisper_blocks as wb

    w = torch.tensor([[0.25, 0.75], [1.0, 0.0]])
    with sar.route_mix(w):
        assert torch.equal(routing.current_mix_weights(), w) and routing.current_utt_adapter().tolist() == [0, 0]
        with sar.route_base():
            assert routing.current_mix_weights() is None
        assert routing.current_mix_weights() is not None
    assert routing.current_mix_weights() is None and routing.current_utt_adapter() is None

    m = copy.deepcopy(micro_model)
    for name in ("a", "b"):
        sar.inject_lora(m, sar.LoraConfig(r=16, lora_alpha=32, target_modules=["q_proj", "v_proj"]), adapter_name=name)
    attn = m.model.encoder.layers[0].self_attn
    with torch.no_grad():
        for mod in (attn.q_proj, attn.v_proj):
            for name in ("a", "b"):
                mod.lora_B[name].weight.normal_(0, 0.02)
    pack = wb._ProjPack([attn.q_proj, attn.k_proj, attn.v_proj], [0.125, 1.0, 1.0])
    A_m, Bp_m, rp = pack.merged()
    assert rp == 16 and tuple(A_m.shape) == (2, 32, 256) and tuple(Bp_m.shape) == (2, 256, 64)
    assert torch.equal(A_m[0, 16:32].float(), attn.q_proj.lora_A["b"].weight.to(torch.bfloat16).float())
    assert torch.equal(A_m[1, :16].float(), attn.v_proj.lora_A["a"].weight.to(torch.bfloat16).float())
    # q's segment scale 2^-3 is folded into its B (exact), v's is 1
    assert torch.equal(Bp_m[0, :, 16:32].float(), (attn.q_proj.lora_B["b"].weight * 0.125).to(torch.bfloat16).float())
    assert torch.equal(Bp_m[1, :, :16].float(), attn.v_proj.lora_B["a"].weight.to(torch.bfloat16).float())
    assert torch.count_nonzero(Bp_m[:, :, 32:]) == 0                       # rank padding


def test_operand_refresh_descriptor_matches_the_header_and_cache_harvest_sees_every_stack():
    """sar_refresh_desc (include/sar.h) <-> the ctypes mirror used to build the device table; cached_tensors() returns
    every tensor a captured graph may address in a module's operand cache (GraphedTrainStep keeps them alive)."""
    import ctypes
    import re

    from speech_adapter_routing_b200.operand_refresh import _Desc, cached_tensors

    header = (Path(__file__).resolve().parents[1] / "include" / "sar.h").read_text()
    body = re.search(r"typedef struct sar_refresh_desc \{(.*?)\} sar_refresh_desc;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [n.strip() for decl in body.split(";") if decl.strip()
              for n in decl.replace("const ", "").replace("*", " ").split(None, 1)[1].split(",")]
    assert fields == [n for n, _ in _Desc._fields_]
    assert ctypes.sizeof(_Desc) == 64

    m = sar.RoutedLoRALinear(nn.Linear(128, 256), "a", r=16, lora_alpha=32)
    m.add_adapter("b", 16, 32)
    st = m._stacks(backward=True)
    held = {t.data_ptr() for t in cached_tensors(nn.Sequential(m))}
    for k in ("W", "A", "Bp", "Wt", "At", "Bt"):
        assert st[k].data_ptr() in held, k


def test_training_layer_inputs_follow_where_the_gradients_go():
    """whisper_train._fn_inputs: when K3 writes every LoRA gradient of a layer straight into the flat bucket the weights
    are not autograd inputs of the layer Function (no AccumulateGrad nodes, CUDA-graph capture does not depend on older
    autograd graphs); a zero-size leaf keeps a layer whose activation input needs no gradient differentiable; otherwise
    the weights are passed and _weight_grads returns one gradient per weight."""
    from speech_adapter_routing_b200 import whisper_train as wt
    from speech_adapter_routing_b200.dist import FlatGradBucket

    mods = [sar.RoutedLoRALinear(nn.Linear(128, 128), "default", r=16, lora_alpha=32) for _ in range(2)]
    ws = wt._lora_params(mods)
    assert [tuple(w.shape) for w in ws] == [(16, 128), (128, 16), (16, 128), (128, 16)]
    h = torch.zeros(1, 4, 128)
    assert wt._fn_inputs(ws, mods, h, False) is ws                    # no bucket: autograd accumulates
    bucket = FlatGradBucket(ws)
    assert wt._all_direct(mods, h.device)
    anchor = wt._fn_inputs(ws, mods, h, False)
    assert len(anchor) == 1 and anchor[0].numel() == 1 and anchor[0].requires_grad
    assert wt._fn_inputs(ws, mods, h.clone().requires_grad_(True), False) == []
    assert wt._fn_inputs(ws, mods, h, True) == []
    ws[0].grad = None                                                  # optimizer.zero_grad(set_to_none=True) after zero_()
    assert wt._fn_inputs(ws, mods, h, False) is ws
    bucket.attach()
    mods[1].add_adapter("second", 16, 32)                              # two adapters: K3's partials are scaled per adapter
    assert not wt._all_direct(mods, h.device)

    class Ctx:
        n_ws = 4
    assert wt._weight_grads(Ctx, [None, 1, None, 2]) == [None, 1, None, 2]
    Ctx.n_ws = 1
    assert wt._weight_grads(Ctx, [None] * 4) == [None]
    with pytest.raises(RuntimeError, match="left the flat bucket"):
        wt._weight_grads(Ctx, [None, torch.zeros(1), None, None])


def test_lora_dropout_terms_of_the_fused_training_layers_complete_pefts_formula():
    """whisper_train._lora_u_dropout / _lora_dropout_bwd (plain torch ops around the kernels): with the drop-free parts
    that the U pass and K3 compute, they give PEFT's y = base(x) + s·B(A(drop(x))) and its gradients — checked against
    fp32 autograd of that formula with the same mask."""
    import types

    from speech_adapter_routing_b200 import whisper_train as wt

    torch.manual_seed(0)
    d, r, s = 128, 16, 2.0
    m = sar.RoutedLoRALinear(nn.Linear(d, d).to(torch.bfloat16), "default", r=r, lora_alpha=32, lora_dropout=0.25)
    with torch.no_grad():
        m.lora_B["default"].weight.normal_(0, 0.05)
    m.train()
    x = torch.randn(2, 5, d).to(torch.bfloat16)
    dy = (torch.randn(2, 5, d) * 0.1).to(torch.bfloat16)
    A16 = m.lora_A["default"].weight.detach().to(torch.bfloat16).float()
    B16 = m.lora_B["default"].weight.detach().to(torch.bfloat16).float()

    torch.manual_seed(5)                                   # the mask the module's nn.Dropout will draw
    keep_scale = m.lora_dropout["default"](torch.ones_like(x)).float()
    torch.manual_seed(5)
    u = ((x.float().view(-1, d) @ A16.t()) * s).to(torch.bfloat16).view(1, 2, 5, r).clone()
    gs = wt._lora_u_dropout(types.SimpleNamespace(lora_mods=[m]), x, None, u)
    assert torch.equal(gs[0].float() + 1, keep_scale)
    xr = x.float().requires_grad_(True)
    Ar = A16.clone().requires_grad_(True)
    u_ref = ((xr * keep_scale).view(-1, d) @ Ar.t()) * s
    assert ((u.float().view(-1, r) - u_ref).abs().max() / u_ref.abs().max()).item() <= 2 ** -6
    ((u_ref @ B16.t()).view(2, 5, d) * dy.float()).sum().backward()

    v = (dy.float().view(-1, d) @ B16) * s                 # what K3 computes without dropout
    dA_base, dx_base = v.t() @ x.float().view(-1, d), (v @ A16).view(2, 5, d).to(torch.bfloat16)
    part, grads = wt._lora_dropout_bwd(m, dy, x, gs[0], dx_base, [dA_base, None], None)
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
    assert rel(grads[0], Ar.grad) <= 2 ** -6 and grads[1] is None
    assert rel(part, xr.grad) <= 2 ** -6
