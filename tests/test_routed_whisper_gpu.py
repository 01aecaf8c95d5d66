"""End-to-end GPU parity of the drop-in Python API (RoutedLoRALinear / WhisperLoRA / AdapterRouter) against the CPU
oracle of the reference's routed forward (oracle/whisper.py) on identical seeded synthetic inputs.

Gates (north star): router adapter indices BIT-EXACT; every LoRA'd module output and the final logits within a
stated bf16 tolerance of the fp32 oracle; identical greedy token sequences on the check set.

Tolerances: the GPU path stores activations in bf16 (2^-8 relative rounding per op) through 2·L layers, the oracle
is fp32 end to end.  Per LoRA'd module output and logits: max|err| <= 3e-2 * max|ref|.

Greedy tokens: a row must equal the fp32 reference tokens at every position before the first one whose reference
top-1 / top-2 logit margin is below MARGIN_EPS (there the argmax of a bf16 path is legitimately ambiguous, and every
later token depends on it); rows without such a position must be identical over their whole length.  No allowance
for "most rows".
"""
import contextlib
import copy
from pathlib import Path

import pytest
import torch

import speech_adapter_routing_b200 as sar
from oracle import fixtures, lora as olora, router as orouter, whisper as owhisper
from speech_adapter_routing_b200 import ops

pytestmark = pytest.mark.gpu

MODULE_TOL = 3e-2
LOGIT_TOL = 3e-2
MARGIN_EPS = 2e-2     # in units of max|logit| of the row's reference logits (bf16 logits carry ~2^-8 relative noise)


def rel_err(y, ref):
    y, ref = y.float().cpu(), ref.float().cpu()
    return ((y - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


@contextlib.contextmanager
def capture_lora_modules(whisper):
    """path -> [output of that q_proj / v_proj per call], whichever way the layer ran: forward hooks for HF's bodies
    over the module slots, whisper_blocks.PROJ_CAPTURE for the fused bodies (which call the kernels directly)."""
    from speech_adapter_routing_b200 import whisper_blocks as wb

    mods = sar.lora_modules(whisper)
    captured = {}
    hooks = [m.register_forward_hook(lambda mod, inp, out, p=p: captured.setdefault(p, []).append(out.detach()))
             for p, m in mods.items()]
    wb.PROJ_CAPTURE = {}
    try:
        yield captured
    finally:
        by_id = {id(m): p for p, m in mods.items()}
        for mid, outs in wb.PROJ_CAPTURE.items():
            if mid in by_id:
                captured.setdefault(by_id[mid], []).extend(outs)
        wb.PROJ_CAPTURE = None
        for hk in hooks:
            hk.remove()


class Setup:
    def __init__(self, geo, C, r, dev, tmp_path):
        self.geo, self.C, self.r, self.dev = geo, C, r, dev
        ref_model = owhisper.build_whisper(geo)
        self.cfg = ref_model.config
        self.weights = owhisper.make_adapter_weights(ref_model, r, C)
        sd = fixtures.make_router_state_dict(self.cfg.d_model, C)
        self.oracle = owhisper.RoutedWhisperOracle(ref_model, self.weights, r, 2 * r, sd)
        protos = owhisper.make_input_features(C, self.cfg.num_mel_bins, list(range(C)), C, seed=99)
        self.oracle.router_sd = owhisper.fit_router_head(sd, self.oracle.lid_features(protos))
        self.languages = [f"lang{k}" for k in range(C)]
        self.paths = {}
        for k, lang in enumerate(self.languages):     # adapters reach the product through PEFT's on-disk layout
            owhisper.write_peft_adapter(tmp_path / lang, self.weights, k, r, 2 * r)
            self.paths[lang] = tmp_path / lang
        gpu_model = owhisper.build_whisper(geo).to(torch.bfloat16).to(dev)
        clf = sar.LanguageClassifier(input_dim=self.cfg.d_model, num_classes=C, languages=self.languages)
        clf.load_state_dict(self.oracle.router_sd)
        self.router = sar.AdapterRouter(gpu_model, self.paths, clf.eval().to(dev), self.languages).eval()

    def batch(self, B, T_dec, kind="uniform", seed=0):
        langs = fixtures.language_mix(B, self.C, kind, seed=7 + seed)
        x = owhisper.make_input_features(B, self.cfg.num_mel_bins, langs, self.C, seed=2234 + seed)
        dec, labels = owhisper.make_decoder_inputs(B, T_dec, self.cfg.vocab_size, self.cfg.decoder_start_token_id)
        return x, dec, labels, langs


@pytest.fixture(scope="module")
def micro(cuda_dev, tmp_path_factory):
    return Setup("micro", 4, 16, cuda_dev, tmp_path_factory.mktemp("micro"))


@pytest.fixture(scope="module")
def tiny(cuda_dev, tmp_path_factory):
    return Setup("tiny", 4, 16, cuda_dev, tmp_path_factory.mktemp("tiny"))


@pytest.mark.parametrize("kind", ["uniform", "skewed", "single"])
def test_router_indices_bit_exact_and_logits_within_tolerance(micro, kind):
    s = micro
    x, dec, labels, langs = s.batch(8, 12, kind)
    ref = s.oracle.forward_hard(x, dec, labels, capture=True)
    xg = x.to(s.dev).to(torch.bfloat16)
    with torch.no_grad():
        h = s.router.extract_encoder_features(xg)
        routed = s.router.detect_indices(h)
        with capture_lora_modules(s.router.whisper) as captured:   # routed pass only (the LID pass ran on base weights)
            out = s.router._hard_routing(xg, routed.idx, labels.to(s.dev))
    assert torch.equal(routed.idx.cpu().long(), ref["idx"])                       # bit-exact routing
    assert ref["idx"].tolist() == langs                                           # and it is the intended mix
    for p, want in ref["captured"].items():                                       # every LoRA'd module, both stacks
        assert rel_err(captured[p][0], want) <= MODULE_TOL, p
    assert rel_err(out["logits"], ref["logits"]) <= LOGIT_TOL
    assert abs(out["loss"].item() - ref["loss"].item()) <= 2e-2 * abs(ref["loss"].item())
    assert out["logits"].shape == ref["logits"].shape


def test_forward_public_api_and_language_names(micro):
    s = micro
    x, dec, labels, langs = s.batch(6, 10, seed=1)
    xg = x.to(s.dev).to(torch.bfloat16)
    with torch.no_grad():
        out = s.router(xg, labels=labels.to(s.dev))
        names, probs = s.router.detect_language(s.router.extract_encoder_features(xg))
    assert set(out) == {"loss", "logits"}
    assert names == [s.languages[k] for k in langs]
    assert probs.shape == (6, s.C) and torch.allclose(probs.sum(-1).cpu(), torch.ones(6), atol=1e-5)
    ref = s.oracle.forward_hard(x, dec, labels)
    assert rel_err(out["logits"], ref["logits"]) <= LOGIT_TOL


def test_fused_lid_pass_equals_the_two_step_api(micro):
    """AdapterRouter.route_inputs (encoder final LayerNorm + LID LayerNorm + mean-pool inside K2, on the residual stream)
    gives what detect_indices(extract_encoder_features(x)) gives: same adapters, same probabilities."""
    s = micro
    x, *_ = s.batch(8, 4, "skewed", seed=6)
    xg = x.to(s.dev).to(torch.bfloat16)
    with torch.no_grad():
        fused = s.router.route_inputs(xg)
        two = s.router.detect_indices(s.router.extract_encoder_features(xg))
    assert torch.equal(fused.idx, two.idx) and torch.equal(fused.perm, two.perm)
    assert (fused.probs - two.probs).abs().max().item() <= 1e-3
    ref_idx, _, _ = s.oracle.detect(x)
    assert torch.equal(fused.idx.cpu().long(), ref_idx)


def test_lid_features_come_from_base_weights(micro):
    """The LID pass must not see any adapter: features equal those of the un-adapted oracle encoder."""
    s = micro
    x, *_ = s.batch(4, 4, seed=2)
    h = s.router.extract_encoder_features(x.to(s.dev).to(torch.bfloat16))
    assert rel_err(h, s.oracle.lid_features(x)) <= MODULE_TOL


def test_tiny_geometry_end_to_end(tiny):
    s = tiny
    x, dec, labels, langs = s.batch(5, 16, "uniform", seed=3)
    ref = s.oracle.forward_hard(x, dec, labels)
    with torch.no_grad():
        xg = x.to(s.dev).to(torch.bfloat16)
        routed = s.router.detect_indices(s.router.extract_encoder_features(xg))
        out = s.router(xg, labels=labels.to(s.dev))
    assert torch.equal(routed.idx.cpu().long(), ref["idx"])
    assert rel_err(out["logits"], ref["logits"]) <= LOGIT_TOL


def test_soft_routing_matches_reference_semantics(micro):
    """Soft strategy = every adapter on the whole batch, logits mixed by LID probability (reference :627-670)."""
    s = micro
    x, dec, labels, _ = s.batch(4, 8, seed=4)
    _, rout, _ = s.oracle.detect(x)
    probs = rout["probs"]
    want = None
    for k in range(s.C):
        r = s.oracle.forward_hard(x, dec, idx=torch.full((4,), k))
        term = probs[:, k, None, None] * r["logits"]
        want = term if want is None else want + term
    s.router.strategy = "soft"
    try:
        with torch.no_grad():
            out = s.router(x.to(s.dev).to(torch.bfloat16), decoder_input_ids=dec.to(s.dev))
    finally:
        s.router.strategy = "hard"
    assert set(out) == {"loss", "logits", "probs"}
    assert rel_err(out["logits"], want) <= LOGIT_TOL


def assert_greedy_rows(got, want, margins, absmax):
    """Row-by-row greedy gate (module docstring).  margins [B, L]: reference top-2 margins, teacher-forced on ``want``."""
    got, want = got.cpu(), want.cpu()
    assert got.shape[0] == want.shape[0]
    n_full = 0
    for i in range(want.shape[0]):
        L = want.shape[1]
        low = (margins[i, :L] < MARGIN_EPS * absmax).nonzero()
        n_ok = int(low[0]) if len(low) else L
        assert got.shape[1] >= n_ok, (i, got.shape, n_ok)
        assert torch.equal(got[i, :n_ok], want[i, :n_ok]), (i, n_ok, got[i].tolist(), want[i].tolist())
        if n_ok == L:
            assert got.shape[1] == L or bool((got[i, L:] == 0).all())
            n_full += 1
    return n_full


def oracle_margins(s, x, want, idx):
    """Reference top-2 margins at every generated position, teacher-forced on the reference's own tokens."""
    B, L = want.shape
    m = torch.zeros(B, L)
    absmax = 0.0
    start = torch.full((1, 1), s.cfg.decoder_start_token_id)
    for i in range(B):
        dec_in = torch.cat([start, want[i:i + 1, :-1]], 1)
        lg = s.oracle.forward_hard(x[i:i + 1], dec_in, idx=idx[i:i + 1])["logits"][0]
        m[i] = orouter.top2_margin(lg)
        absmax = max(absmax, lg.abs().max().item())
    return m, absmax


def test_soft_fused_strategy_weighted_mix_inside_the_projections(micro):
    """strategy="soft_fused" (opt-in): ONE forward whose every LoRA'd projection applies
    Σ_k p[b,k]·s·B_k A_k x inside the fused kernels — against its own oracle (per-utterance fp32 forward with the mixed
    LoRA term), module by module and at the logits; one-hot weights reduce to hard routing; top_k renormalises."""
    s = micro
    x, dec, labels, _ = s.batch(6, 8, seed=14)
    xg = x.to(s.dev).to(torch.bfloat16)
    g = torch.Generator().manual_seed(3)
    w = torch.softmax(torch.randn(6, s.C, generator=g) * 1.5, dim=-1)              # genuinely mixed weights
    ref = s.oracle.forward_soft_fused(x, dec, w, labels)
    for m in s.oracle.mods.values():
        m.capture = None
    with torch.no_grad():
        with capture_lora_modules(s.router.whisper) as captured:
            out = s.router._soft_fused_routing(xg, w.to(s.dev), labels.to(s.dev))
    assert set(out) == {"loss", "logits", "probs"}
    assert rel_err(out["logits"], ref["logits"]) <= LOGIT_TOL
    assert abs(out["loss"].item() - ref["loss"].item()) <= 2e-2 * abs(ref["loss"].item())
    assert len(captured) == len(s.oracle.mods)                                     # every q_proj / v_proj ran fused
    # the public strategy switch: LID probabilities as weights
    s.router.strategy = "soft_fused"
    try:
        with torch.no_grad():
            pub = s.router(xg, labels=labels.to(s.dev))
            probs = s.router.route_inputs(xg).probs
    finally:
        s.router.strategy = "hard"
    want = s.oracle.forward_soft_fused(x, dec, probs.cpu(), labels)
    assert rel_err(pub["logits"], want["logits"]) <= LOGIT_TOL
    # one-hot weights == hard routing (same adapters, U from a different kernel: equal to bf16 rounding noise)
    idx = torch.tensor([1, 3, 0, 2, 2, 1])
    onehot = torch.nn.functional.one_hot(idx, s.C).float()
    with torch.no_grad():
        a = s.router._soft_fused_routing(xg, onehot.to(s.dev), labels.to(s.dev))
        b = s.router._hard_routing(xg, idx.to(torch.int32).to(s.dev), labels.to(s.dev))
    assert rel_err(a["logits"], b["logits"]) <= 1e-2 and abs(a["loss"].item() - b["loss"].item()) <= 1e-2
    # top-1 of the mixed weights is hard routing on their argmax
    with torch.no_grad():
        t1 = s.router._soft_fused_routing(xg, w.to(s.dev), labels.to(s.dev), top_k=1)
        h1 = s.router._hard_routing(xg, w.argmax(-1).to(torch.int32).to(s.dev), labels.to(s.dev))
    assert rel_err(t1["logits"], h1["logits"]) <= 1e-2


def test_greedy_generation_matches_oracle_tokens(micro):
    s = micro
    B, steps = 8, 32
    x, *_ = s.batch(B, 4, seed=5)
    want = s.oracle.generate_hard(x, max_new_tokens=steps)
    with torch.no_grad():
        got = s.router.generate(x.to(s.dev).to(torch.bfloat16), max_new_tokens=steps, num_beams=1, do_sample=False)
    idx, _, _ = s.oracle.detect(x)
    margins, absmax = oracle_margins(s, x, want, idx)
    assert_greedy_rows(got, want, margins, absmax)


def test_module_default_and_beam_expanded_routing(cuda_dev):
    """RoutedLoRALinear outside any routing context uses its active adapter (plain WhisperLoRA behaviour); inside a
    context with fewer utterances than batch rows (beam search) the index is expanded."""
    dev = cuda_dev
    c = fixtures.make_lora_case(4, 50, 256, 256, 16, 3)
    lin = torch.nn.Linear(256, 256).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(c.W)
        lin.bias.copy_(c.bias)
    m = sar.RoutedLoRALinear(lin, None)
    for k in range(3):
        m.add_adapter(f"a{k}", 16, 32)
        with torch.no_grad():
            m.lora_A[f"a{k}"].weight.copy_(c.A_stack[k])
            m.lora_B[f"a{k}"].weight.copy_(c.B_stack[k])
    m.eval()
    m.set_adapter("a2")
    x = c.x.to(dev)
    with torch.no_grad():
        y = m(x)
        ref = olora.lora_linear_routed_k1_rounding(c.x, c.W, c.bias, c.A_stack, c.B_stack, 2.0,
                                                   torch.full((4,), 2, dtype=torch.int32))
        assert rel_err(y, ref) <= 2 ** -7
        with sar.route(torch.tensor([0, 1], device=dev)):          # 2 utterances x 2 beams
            yb = m(x)
        refb = olora.lora_linear_routed_k1_rounding(c.x, c.W, c.bias, c.A_stack, c.B_stack, 2.0,
                                                    torch.tensor([0, 0, 1, 1], dtype=torch.int32))
        assert rel_err(yb, refb) <= 2 ** -7
        m.disable_adapters = True
        assert rel_err(m(x), torch.nn.functional.linear(c.x.float(), c.W.float(), c.bias.float())) <= 2 ** -7
        m.disable_adapters = False
        assert m(x.float()).dtype == torch.float32                  # output follows the input dtype
        with sar.route(torch.tensor([0, 1, 2], device=dev)), pytest.raises(ValueError):
            m(x)


@pytest.mark.parametrize("fused,B_", [(True, 2), (False, 2), (True, 3), (False, 3)])
def test_whisper_lora_training_step_grads_match_oracle(cuda_dev, monkeypatch, fused, B_):
    """BASELINE config 5 shape family (single adapter, fwd + LoRA-only bwd): gradients of every lora_A / lora_B from
    K3 through autograd equal fp32 autograd of the oracle model on the same batch — through the fused training layers
    (whisper_train.py) and through HF's layer bodies over the K1 / K3 module slots; B_ = 3 puts the encoder's
    projections (4500 rows) on the split LoRA path of K1 and K3."""
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    from speech_adapter_routing_b200 import whisper_train

    monkeypatch.setattr(whisper_train, "ENABLED", fused)
    dev = cuda_dev
    w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda",
                        use_gradient_checkpointing=False)
    cfg = w.model.config
    hf = w.model.base_model.model
    # fp32 oracle copy with identical base weights (bf16 values) and identical adapter weights
    ref_model = owhisper.build_whisper("tiny")
    ref_model.load_state_dict({k: v.float().cpu() for k, v in hf.state_dict().items()
                               if ".lora_" not in k and ".base_layer." not in k}, strict=False)
    weights = owhisper.make_adapter_weights(ref_model, 16, 1)
    for p, m in sar.lora_modules(hf).items():
        ref_lin = ref_model.get_submodule(p)
        with torch.no_grad():
            ref_lin.weight.copy_(m.base_layer.weight.float().cpu())
            ref_lin.bias.copy_(m.base_layer.bias.float().cpu())
            m.lora_A["default"].weight.copy_(weights[p][0][0])
            m.lora_B["default"].weight.copy_(weights[p][1][0])
    params = {}
    for p, (A, B) in weights.items():
        params[p] = (A[0].clone().requires_grad_(True), B[0].clone().requires_grad_(True))
        lin = ref_model.get_submodule(p)

        def fwd(x, lin=lin, A=params[p][0], B=params[p][1]):
            return olora.lora_linear(x, lin.weight, lin.bias, A, B, 2.0)
        lin.forward = fwd
    T_dec = 6
    x = owhisper.make_input_features(B_, cfg.num_mel_bins, [0] * B_, 1, seed=8)
    dec, labels = owhisper.make_decoder_inputs(B_, T_dec, cfg.vocab_size, cfg.decoder_start_token_id)
    ref_model.train(False)
    ref_loss = ref_model(input_features=x, labels=labels).loss
    ref_loss.backward()

    w.train()
    calls0 = dict(whisper_train.CALLS)
    out = w(input_features=x.to(dev).to(torch.bfloat16), labels=labels.to(dev))
    out.loss.backward()
    # the fused training layers (forward + backward of a whole layer on libsar, whisper_train.py) are what ran
    n_fused = (cfg.encoder_layers, cfg.decoder_layers) if fused else (0, 0)
    assert whisper_train.CALLS["encoder_layers"] - calls0["encoder_layers"] == n_fused[0]
    assert whisper_train.CALLS["decoder_layers"] - calls0["decoder_layers"] == n_fused[1], whisper_train.REFUSED
    assert whisper_train.CALLS.get("lm_head", 0) - calls0.get("lm_head", 0) == (1 if fused else 0)
    assert abs(out.loss.item() - ref_loss.item()) <= 2e-2 * abs(ref_loss.item())
    checked = 0
    for p, m in sar.lora_modules(hf).items():
        gA, gB = m.lora_A["default"].weight.grad, m.lora_B["default"].weight.grad
        assert gA is not None and gB is not None and gA.dtype == torch.float32
        rA, rB = params[p][0].grad, params[p][1].grad
        # layer-wise gradients shrink with depth; compare against the largest gradient of the same kind
        assert rel_err(gA, rA) <= 8e-2, p
        assert rel_err(gB, rB) <= 8e-2, p
        checked += 1
    assert checked == 6 * cfg.encoder_layers
    assert all(p.grad is None for n, p in w.model.named_parameters() if ".lora_" not in n)


def test_save_and_reload_adapter_on_gpu(cuda_dev, tmp_path, monkeypatch):
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda")
    with torch.no_grad():
        for m in sar.lora_modules(w.model).values():
            m.lora_B["default"].weight.normal_(0, 0.02)
    cfg = w.model.config
    x = owhisper.make_input_features(2, cfg.num_mel_bins, [0, 0], 1, seed=9).to("cuda").to(torch.bfloat16)
    dec, _ = owhisper.make_decoder_inputs(2, 5, cfg.vocab_size, cfg.decoder_start_token_id)
    w.eval()
    with torch.no_grad():
        a = w(input_features=x, decoder_input_ids=dec.to("cuda")).logits
    w.save_adapter(tmp_path / "final")
    w2 = sar.load_whisper_lora_from_checkpoint(tmp_path / "final", "whisper-tiny", device="cuda")
    w2.eval()
    with torch.no_grad():
        b = w2(input_features=x, decoder_input_ids=dec.to("cuda")).logits
    # w was built for training (use_cache off -> fused decoder blocks); the reloaded model keeps HF's use_cache
    # default (KV cache -> HF decoder bodies over the K1 module slots): same weights, different rounding points
    assert (a.float() - b.float()).abs().max().item() <= 2e-2 * a.float().abs().max().item()
    w2.model.config.use_cache = w.model.config.use_cache
    with torch.no_grad():
        b = w2(input_features=x, decoder_input_ids=dec.to("cuda")).logits
    assert torch.equal(a, b)   # same code path -> bit-identical after the save / load round trip
    ids = w2.generate(x, max_new_tokens=4)
    assert ids.shape[0] == 2


def test_config1_whisper_small_single_adapter_one_clip(cuda_dev, tmp_path):
    """BASELINE config 1: whisper-small geometry, a single LoRA r16 on q_proj / v_proj, batch 1, one synthetic 30 s clip —
    the case the reference itself can run on the CPU.  fp32 CPU oracle vs the bf16 B200 path: logits within tolerance,
    identical greedy tokens up to the first position whose oracle top-2 margin is below the bf16 noise floor."""
    s = Setup("small", 1, 16, cuda_dev, tmp_path)
    x, dec, labels, _ = s.batch(1, 16, seed=11)
    ref = s.oracle.forward_hard(x, dec, labels)
    with torch.no_grad():
        out = s.router(x.to(s.dev).to(torch.bfloat16), labels=labels.to(s.dev))
    assert rel_err(out["logits"], ref["logits"]) <= LOGIT_TOL
    assert abs(out["loss"].item() - ref["loss"].item()) <= 5e-2 * abs(ref["loss"].item())
    steps = 12
    want = s.oracle.generate_hard(x, max_new_tokens=steps)
    with torch.no_grad():
        got = s.router.generate(x.to(s.dev).to(torch.bfloat16), max_new_tokens=steps, num_beams=1, do_sample=False).cpu()
    margins, absmax = oracle_margins(s, x, want, torch.zeros(1, dtype=torch.long))
    assert_greedy_rows(got, want, margins, absmax)
    assert got.shape[0] == 1 and got.dtype == torch.long


def test_training_forward_backward_with_lora_dropout_matches_pefts_formula(cuda_dev):
    """The reference trains with lora_dropout = 0.1 (scripts/train_lora.py:55).  In training mode the module output
    must be base(x) + s·B(A(drop(x))) for the mask drawn in that call (K1 + the autograd correction term), and the
    backward (K3 + autograd) must give the gradients of that expression."""
    import torch.nn as nn

    from speech_adapter_routing_b200.lora_linear import RoutedLoRALinear

    torch.manual_seed(3)
    d, r, p = 384, 16, 0.1
    base = nn.Linear(d, d).to(cuda_dev).to(torch.bfloat16)
    m = RoutedLoRALinear(base, "hindi", r=r, lora_alpha=2 * r, lora_dropout=p).to(cuda_dev)
    nn.init.normal_(m.lora_B["hindi"].weight, std=0.05)
    m.train()
    x = (torch.randn(4, 200, d, device=cuda_dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    torch.manual_seed(77)
    y = m(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    gA, gB, gx = m.lora_A["hindi"].weight.grad.float().clone(), m.lora_B["hindi"].weight.grad.float().clone(), x.grad.float().clone()

    xr = x.detach().float().requires_grad_(True)
    A = m.lora_A["hindi"].weight.detach().float().requires_grad_(True)
    Bw = m.lora_B["hindi"].weight.detach().float().requires_grad_(True)
    torch.manual_seed(77)
    mask = nn.functional.dropout(torch.ones_like(x), p=p, training=True).float()     # same generator state, same shape/dtype
    ref = nn.functional.linear(xr, base.weight.float(), base.bias.float()) + 2.0 * ((xr * mask) @ A.t()) @ Bw.t()
    ref.backward(dy.float())
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    assert rel(y.detach().float(), ref.detach()) <= 2.0 ** -6
    assert rel(gx, xr.grad) <= 2.0 ** -5
    assert rel(gA, A.grad) <= 2.0 ** -5 and rel(gB, Bw.grad) <= 2.0 ** -5


# ------------------------------------------------------------------------------------------------ reference-run golden
class GoldenSetup:
    """The product's AdapterRouter built on the regenerated inputs of one tests/golden/routed_forward_golden.pt case
    (outputs of the reference's unmodified AdapterRouter, tests/golden/make_routed_golden.py)."""

    def __init__(self, rec, dev, tmp):
        from golden_cases import Case

        self.rec, self.dev = rec, dev
        c = self.c = Case(rec)
        self.languages = rec["languages"]
        paths = {}
        for k, lang in enumerate(self.languages):
            owhisper.write_peft_adapter(tmp / lang, c.weights, k, rec["r"], 2 * rec["r"])
            paths[lang] = tmp / lang
        gpu_model = owhisper.build_whisper(rec["geometry"]).to(torch.bfloat16).to(dev)
        clf = sar.LanguageClassifier(input_dim=c.cfg.d_model, num_classes=rec["C"], languages=self.languages)
        clf.load_state_dict(c.router_sd)
        self.router = sar.AdapterRouter(gpu_model, paths, clf.eval().to(dev), self.languages).eval()
        self.x = c.x.to(dev).to(torch.bfloat16)
        self.labels = c.labels.to(dev)


@pytest.fixture(scope="module", params=[0, 1])
def golden(request, cuda_dev, tmp_path_factory):
    from golden_cases import load_golden

    rec = load_golden()["cases"][request.param]
    return GoldenSetup(rec, cuda_dev, tmp_path_factory.mktemp(f"golden{request.param}"))


def test_reference_run_golden_forward_hard_soft_threshold(golden):
    """B200 path vs the REFERENCE's own AdapterRouter outputs (fp32 CPU run, committed golden): detected languages
    bit-exact, logits of every strategy within LOGIT_TOL, losses within 2 %."""
    s, rec = golden, golden.rec
    r = s.router
    with torch.no_grad():
        h = r.extract_encoder_features(s.x)
        names, probs = r.detect_language(h)
        assert names == rec["names"]                                                   # :550-566
        assert torch.equal(r.detect_indices(h).idx.cpu().long(), rec["idx"])
        assert (probs.float().cpu() - rec["probs"]).abs().max().item() <= 2e-2
        assert rel_err(h[:, 0, :], rec["lid_features_row0"]) <= MODULE_TOL

        out = r(s.x, labels=s.labels)                                                  # hard :599-625, :695-713
        assert set(out) == {"loss", "logits"}
        assert rel_err(out["logits"], rec["hard_labels"]["logits"]) <= LOGIT_TOL
        assert abs(out["loss"].item() - rec["hard_labels"]["loss"].item()) <= 2e-2 * rec["hard_labels"]["loss"].item()
        try:
            r.strategy = "soft"                                                        # :627-670
            out = r(s.x, labels=s.labels)
            assert set(out) == {"loss", "logits", "probs"}
            assert rel_err(out["logits"], rec["soft_labels"]["logits"]) <= LOGIT_TOL
            assert abs(out["loss"].item() - rec["soft_labels"]["loss"].item()) <= 2e-2 * rec["soft_labels"]["loss"].item()
            for key in ("threshold_0p5", "threshold_1m"):                              # :672-693
                r.strategy, r.threshold = "threshold", rec[key]["threshold"]
                out = r(s.x, labels=s.labels)
                assert sorted(k for k, v in out.items() if v is not None) == rec[key]["keys"]
                assert rel_err(out["logits"], rec[rec[key]["same_as"]]["logits"]) <= LOGIT_TOL
                assert abs(out["loss"].item() - rec[key]["loss"].item()) <= 2e-2 * rec[key]["loss"].item()
        finally:
            r.strategy, r.threshold = "hard", 0.7


def test_reference_run_golden_generate(golden):
    """generate (:715-761) vs the reference run: same tokens (margin gate), EOS dropped and zeros after it, width = the
    longest row; generate(language=...) = the named adapter's batched HF generate, padded with pad_token_id."""
    s, rec = golden, golden.rec
    n = rec["gen_steps"]
    g = rec["generate"]
    kw = dict(max_new_tokens=n, num_beams=1, do_sample=False)
    with torch.no_grad():
        got = s.router.generate(s.x, **kw)
        full = assert_greedy_rows(got, g["ids"], g["margins"], g["logit_absmax"])
        assert got.dtype == torch.long
        if "generate_eos" in rec:
            e = rec["generate_eos"]
            got_e = s.router.generate(s.x, eos_token_id=e["eos_token_id"], **kw).cpu()
            if full == rec["B"] and torch.equal(got.cpu(), g["ids"]):
                assert torch.equal(got_e, e["ids"])                                    # incl. the zero right-padding
            assert (got_e == e["eos_token_id"]).sum() == 0
        gl = rec["generate_language"]
        got_l = s.router.generate(s.x, language=gl["language"], **kw).cpu()
        k = rec["languages"].index(gl["language"])
        m_l, absmax = oracle_margins(_OracleView(s.c), s.c.x, gl["ids"], torch.full((rec["B"],), k))
        assert_greedy_rows(got_l, gl["ids"], m_l, absmax)
        if "generate_language_eos" in rec:
            e = rec["generate_language_eos"]
            got_le = s.router.generate(s.x, language=gl["language"], eos_token_id=e["eos_token_id"], **kw).cpu()
            if torch.equal(got_l, gl["ids"]):
                assert torch.equal(got_le, e["ids"])


class _OracleView:
    def __init__(self, case):
        self.cfg = case.cfg
        self.oracle = case.oracle()


# ------------------------------------------------------------------------------------------------ benchmarked geometries
def _model_level_check(s, B, T_dec, kind, seed):
    x, dec, labels, langs = s.batch(B, T_dec, kind, seed=seed)
    ref = s.oracle.forward_hard(x, dec, labels, capture=True)
    xg = x.to(s.dev).to(torch.bfloat16)
    with torch.no_grad():
        routed = s.router.detect_indices(s.router.extract_encoder_features(xg))
        with capture_lora_modules(s.router.whisper) as captured:
            out = s.router._hard_routing(xg, routed.idx, labels.to(s.dev))
    assert torch.equal(routed.idx.cpu().long(), ref["idx"]) and ref["idx"].tolist() == langs
    assert rel_err(out["logits"], ref["logits"]) <= LOGIT_TOL
    assert abs(out["loss"].item() - ref["loss"].item()) <= 2e-2 * abs(ref["loss"].item())
    return captured, ref


def test_fused_blocks_are_what_the_micro_model_runs(micro):
    """The model-level fixtures must exercise the benchmarked path: every encoder / decoder layer of the head_dim-64
    micro geometry is bound to the fused body, and a routed forward launches libsar's projection / dense / LayerNorm /
    attention kernels for it (not HF's eager bodies over the K1 module slots)."""
    from speech_adapter_routing_b200 import ops

    w = micro.router.whisper
    layers = list(w.model.encoder.layers) + list(w.model.decoder.layers)
    assert all(hasattr(l, "_sar_pack") for l in layers)
    x, dec, labels, _ = micro.batch(4, 8, seed=21)
    ops.reset_counters()
    with torch.no_grad():
        micro.router(x.to(micro.dev).to(torch.bfloat16), labels=labels.to(micro.dev))
    n_enc, n_dec = len(w.model.encoder.layers), len(w.model.decoder.layers)
    assert ops.LAUNCHES["k1"] == 0                                    # no module-slot K1 calls: the layer bodies are fused
    # LID pass: 2 per encoder layer (its final LayerNorm runs inside K2); routed pass: 2 per encoder layer + final,
    # 3 per decoder layer + final
    assert ops.LAUNCHES["ln"] >= 2 * 2 * n_enc + 1 + 3 * n_dec + 1
    assert ops.LAUNCHES["linear"] >= 2 * 3 * n_enc + 4 * n_dec
    assert ops.LAUNCHES["proj"] >= 2 * n_enc + 3 * n_dec
    assert ops.LAUNCHES["k2"] == 2


def test_whisper_small_mixed_adapter_batch(cuda_dev, tmp_path):
    """BASELINE config 2's geometry and adapter set (whisper-small, 4 adapters r16), mixed-language batch of 8: routing
    bit-exact, logits / loss within tolerance of the fp32 oracle, every LoRA'd module output of one encoder and one
    decoder layer within tolerance."""
    s = Setup("small", 4, 16, cuda_dev, tmp_path)
    captured, ref = _model_level_check(s, 8, 16, "uniform", seed=31)
    for p, want in ref["captured"].items():
        if ".layers.0." in p or ".layers.11." in p:
            assert rel_err(captured[p][0], want) <= MODULE_TOL, p


@pytest.mark.parametrize("geo,C,r", [("medium-2l", 4, 32), ("large-v3-2l", 8, 64)])
def test_medium_and_large_v3_geometry_model_level(cuda_dev, tmp_path, geo, C, r):
    """BASELINE configs 3 / 4 kernel shapes (d = 1024 r32 x4 adapters; d = 1280, 128 mel bins, r64 x8 adapters) on
    two-layer models against the fp32 oracle: routing bit-exact, every LoRA'd module output and the logits within tolerance."""
    s = Setup(geo, C, r, cuda_dev, tmp_path)
    captured, ref = _model_level_check(s, 2 * C if C == 4 else C, 12, "uniform", seed=41)
    for p, want in ref["captured"].items():
        assert rel_err(captured[p][0], want) <= MODULE_TOL, p


def test_whisper_small_greedy_32_tokens_batch_8(cuda_dev, tmp_path):
    """SURVEY §8(d) greedy gate at its stated size: whisper-small, B = 8 mixed languages, 32 new tokens."""
    s = Setup("small", 4, 16, cuda_dev, tmp_path)
    x, *_ = s.batch(8, 4, "uniform", seed=51)
    want = s.oracle.generate_hard(x, max_new_tokens=32)
    with torch.no_grad():
        got = s.router.generate(x.to(s.dev).to(torch.bfloat16), max_new_tokens=32, num_beams=1, do_sample=False)
    idx, _, _ = s.oracle.detect(x)
    margins, absmax = oracle_margins(s, x, want, idx)
    assert_greedy_rows(got, want, margins, absmax)


def test_fused_training_layers_under_gradient_checkpointing_match_the_plain_run(cuda_dev, monkeypatch):
    """HF's gradient checkpointing (the reference default, src/models/whisper_lora.py:81-83) recomputes each layer's
    forward inside backward: with the fused training layers the first forward runs the inference body (no_grad), the
    recompute runs the autograd Function.  Gradients must equal the non-checkpointed run's up to the rounding difference
    between the two bodies (the inference body's LayerNorm kernel vs ATen's: 1e-2 of the largest gradient), and HF's
    eager bodies over the K1 / K3 module slots (SAR_FUSED_TRAIN=0 equivalent) within bf16 noise."""
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    from speech_adapter_routing_b200 import whisper_train

    dev = cuda_dev
    grads = {}
    for mode in ("plain", "ckpt", "hf_bodies"):
        torch.manual_seed(11)                       # lora_A's kaiming init draws from the global generator
        w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda",
                            use_gradient_checkpointing=(mode == "ckpt"))
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            for m in sar.lora_modules(w.model).values():
                m.lora_B["default"].weight.copy_((torch.randn(m.out_features, 16, generator=g) * 0.02).to(dev))
        cfg = w.model.config
        x = owhisper.make_input_features(3, cfg.num_mel_bins, [0, 0, 0], 1, seed=8).to(dev).to(torch.bfloat16)
        _, labels = owhisper.make_decoder_inputs(3, 9, cfg.vocab_size, cfg.decoder_start_token_id)
        w.train()
        whisper_train.ENABLED = mode != "hf_bodies"
        try:
            loss = w(input_features=x, labels=labels.to(dev)).loss
            loss.backward()
        finally:
            whisper_train.ENABLED = True
        grads[mode] = (loss.item(), {n: p.grad.float().clone() for n, p in w.model.named_parameters() if p.grad is not None})
    assert len(grads["plain"][1]) == 2 * 6 * cfg.encoder_layers
    assert abs(grads["plain"][0] - grads["ckpt"][0]) <= 1e-2 * abs(grads["plain"][0])
    for n, gp in grads["plain"][1].items():
        assert rel_err(gp, grads["ckpt"][1][n]) <= 1e-2, n
        gh = grads["hf_bodies"][1][n]
        assert rel_err(gp, gh) <= 1.2e-1, n          # two independent bf16 pipelines (each within ~5e-2 of fp32 autograd)
    assert abs(grads["plain"][0] - grads["hf_bodies"][0]) <= 1e-2 * abs(grads["hf_bodies"][0])


def test_graphed_train_step_matches_eager_and_tracks_parameter_updates(cuda_dev, monkeypatch):
    """GraphedTrainStep (forward + backward of the reference's training step, src/training/trainer.py:251-256, replayed
    as one CUDA graph): gradients equal the eager step's on a NEW batch, and still do after the LoRA parameters were
    updated in place (the graph's first node re-derives the cached bf16 operands, operand_refresh.py) — a replay
    that kept reading the operands of capture time would fail the second comparison."""
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    from speech_adapter_routing_b200 import whisper_train

    dev = cuda_dev
    torch.manual_seed(5)
    for ckpt in (False, True):
        w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda",
                            use_gradient_checkpointing=ckpt)
        w.train()
        cfg = w.model.config
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            for m in sar.lora_modules(w.model).values():
                m.lora_B["default"].weight.copy_((torch.randn(m.out_features, 16, generator=g) * 0.02).to(dev))
        params = [p for p in w.model.parameters() if p.requires_grad]
        bucket = sar.FlatGradBucket(params)

        def batch(seed):
            x = owhisper.make_input_features(3, cfg.num_mel_bins, [0, 0, 0], 1, seed=seed).to(dev).to(torch.bfloat16)
            _, labels = owhisper.make_decoder_inputs(3, 9, cfg.vocab_size, cfg.decoder_start_token_id, seed=seed)
            return x, labels.to(dev)

        def eager(x, labels):
            bucket.zero_()
            loss = w(input_features=x, labels=labels).loss
            loss.backward()
            return float(loss), bucket.buffer.clone()

        x0, l0 = batch(8)
        calls0 = dict(whisper_train.CALLS)
        step = sar.GraphedTrainStep(w, bucket, x0, l0, warmup=2)          # check=True: compares with eager on (x0, l0)
        assert whisper_train.CALLS["decoder_layers"] > calls0["decoder_layers"], whisper_train.REFUSED
        assert whisper_train.REFUSED == {"encoder": "", "decoder": ""}   # the capture took the fused layers
        x1, l1 = batch(9)
        loss_g = float(step(x1, l1))
        grad_g = bucket.buffer.clone()
        loss_e, grad_e = eager(x1, l1)
        assert rel_err(grad_g, grad_e) <= 2e-2 and abs(loss_g - loss_e) <= 2e-2 * abs(loss_e)
        assert rel_err(grad_g, eager(x0, l0)[1]) > 5e-2                   # the two batches do differ
        with torch.no_grad():                                             # an optimizer step's worth of change
            for p in params:
                p.add_(torch.randn(p.shape, generator=g).to(dev) * 0.05 * p.abs().max().clamp_min(0.02))
        loss_e2, grad_e2 = eager(x1, l1)
        assert rel_err(grad_e2, grad_e) > 5e-2                            # the update matters
        loss_g2 = float(step(x1, l1))
        assert rel_err(bucket.buffer, grad_e2) <= 2e-2 and abs(loss_g2 - loss_e2) <= 2e-2 * abs(loss_e2)
        del step


def test_eager_training_loop_refreshes_operands_in_one_launch_and_stays_exact(cuda_dev, monkeypatch):
    """The reference's loop (forward, backward, optimizer.step; src/training/trainer.py:251-268) through WhisperLoRA: from
    the third step on the stale bf16 LoRA operands are re-derived by one sar_operand_refresh launch.  The step after
    that must equal the same step with every cache rebuilt from scratch (refresh_operands() invalidates them) — same
    operand bits, so the only difference left is cuDNN's atomically accumulated dQ (5e-3) — and a cache rebuilt behind the
    refresher's back must make it stand down."""
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    dev = cuda_dev
    torch.manual_seed(3)
    w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda",
                        use_gradient_checkpointing=False)
    w.train()
    cfg = w.model.config
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for m in sar.lora_modules(w.model).values():
            m.lora_B["default"].weight.copy_((torch.randn(m.out_features, 16, generator=g) * 0.02).to(dev))
    params = [p for p in w.model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.5)
    x = owhisper.make_input_features(2, cfg.num_mel_bins, [0, 0], 1, seed=8).to(dev).to(torch.bfloat16)
    _, labels = owhisper.make_decoder_inputs(2, 6, cfg.vocab_size, cfg.decoder_start_token_id)
    labels = labels.to(dev)

    def grads():
        opt.zero_grad(set_to_none=True)
        loss = w(input_features=x, labels=labels).loss
        loss.backward()
        return loss.item(), [p.grad.clone() for p in params]

    for _ in range(3):
        grads()
        opt.step()
    assert w.__dict__["_sar_refresh"] is not None                     # built at the start of step 2, used at step 3
    launches0 = ops.LAUNCHES.get("refresh", 0)
    loss_a, ga = grads()                                              # operands refreshed in place by one launch
    assert ops.LAUNCHES.get("refresh", 0) == launches0 + 1
    sar.refresh_operands()                                            # epoch bump: every cache rebuilds from the parameters
    loss_b, gb = grads()
    assert w.__dict__["_sar_refresh"] is None or w.__dict__["_sar_refresh"]._frozen0 != ()   # stood down / rebuilt
    assert loss_a == loss_b
    for a, b in zip(ga, gb):
        assert rel_err(a, b) <= 5e-3


def test_training_step_under_autocast_as_the_reference_trainer_runs_it(cuda_dev, monkeypatch):
    """src/training/trainer.py:328-331 wraps the forward in torch.autocast(bf16) and backpropagates loss / accumulation
    steps: the fused training layers (and the LM head Function) take that unchanged — same layers taken, gradients equal
    to the plain call's scaled by 1 / 2 (cuDNN's atomically accumulated dQ aside)."""
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    from speech_adapter_routing_b200 import whisper_train

    dev = cuda_dev
    torch.manual_seed(7)
    w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda",
                        use_gradient_checkpointing=True)
    w.train()
    cfg = w.model.config
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for m in sar.lora_modules(w.model).values():
            m.lora_B["default"].weight.copy_((torch.randn(m.out_features, 16, generator=g) * 0.02).to(dev))
    x = owhisper.make_input_features(2, cfg.num_mel_bins, [0, 0], 1, seed=8).to(dev).to(torch.bfloat16)
    _, labels = owhisper.make_decoder_inputs(2, 6, cfg.vocab_size, cfg.decoder_start_token_id)
    labels = labels.to(dev)

    def grads(autocast):
        for p in w.model.parameters():
            p.grad = None
        calls0 = dict(whisper_train.CALLS)
        if autocast:
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                loss = w(input_features=x, labels=labels).loss
            (loss / 2).backward()
        else:
            loss = w(input_features=x, labels=labels).loss
            loss.backward()
        taken = {k: whisper_train.CALLS.get(k, 0) - calls0.get(k, 0) for k in ("encoder_layers", "decoder_layers", "lm_head")}
        return loss.item(), {n: p.grad.float().clone() for n, p in w.model.named_parameters() if p.grad is not None}, taken

    loss_a, ga, taken_a = grads(False)
    loss_b, gb, taken_b = grads(True)
    assert taken_a == taken_b and taken_b["decoder_layers"] >= cfg.decoder_layers and taken_b["lm_head"] >= 1, (
        taken_a, taken_b, whisper_train.REFUSED)
    assert abs(loss_a - loss_b) <= 1e-2 * abs(loss_a)
    assert len(ga) == len(gb) == 2 * 6 * cfg.encoder_layers
    for n, a in ga.items():
        assert rel_err(2 * gb[n], a) <= 1e-2, n


class _FixedMaskDropout(torch.nn.Dropout):
    """nn.Dropout whose mask depends only on (seed, shape): the fused layers and HF's bodies draw it in different orders."""

    def __init__(self, p, seed):
        super().__init__(p)
        self.seed = seed

    def forward(self, x):
        if not self.training or self.p == 0:
            return x
        g = torch.Generator().manual_seed(self.seed)
        keep = (torch.rand(x.shape, generator=g) >= self.p).to(x.device)
        return x * keep.to(x.dtype) / (1 - self.p)


def test_fused_training_layers_with_lora_dropout_equal_hf_bodies_with_the_same_masks(cuda_dev, monkeypatch):
    """The reference's default lora_dropout = 0.1 (src/models/whisper_lora.py:30): PEFT drops the LoRA branch's input only.
    The fused training layers add s·(x∘g)·Aᵀ to U (whisper_train._lora_u_dropout) and the matching dA / dx terms; HF's
    bodies over the K1 / K3 module slots add the same remainder with autograd ops (checked against PEFT's formula in
    test_training_forward_backward_with_lora_dropout_matches_pefts_formula).  With identical masks both give the same
    loss and gradients."""
    monkeypatch.setenv("SAR_RANDOM_INIT", "1")
    from speech_adapter_routing_b200 import whisper_train

    dev = cuda_dev
    torch.manual_seed(13)
    w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.1, device="cuda",
                        use_gradient_checkpointing=False)
    w.train()
    cfg = w.model.config
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for i, m in enumerate(sar.lora_modules(w.model).values()):
            m.lora_B["default"].weight.copy_((torch.randn(m.out_features, 16, generator=g) * 0.02).to(dev))
            m.lora_dropout["default"] = _FixedMaskDropout(0.3, 1000 + i)      # a rate that matters
    w.train()
    x = owhisper.make_input_features(2, cfg.num_mel_bins, [0, 0], 1, seed=8).to(dev).to(torch.bfloat16)
    _, labels = owhisper.make_decoder_inputs(2, 6, cfg.vocab_size, cfg.decoder_start_token_id)
    labels = labels.to(dev)

    def grads(fused):
        for p in w.model.parameters():
            p.grad = None
        monkeypatch.setattr(whisper_train, "ENABLED", fused)
        calls0 = dict(whisper_train.CALLS)
        loss = w(input_features=x, labels=labels).loss
        loss.backward()
        taken = whisper_train.CALLS["decoder_layers"] - calls0["decoder_layers"]
        return loss.item(), {n: p.grad.float().clone() for n, p in w.model.named_parameters() if p.grad is not None}, taken

    loss_f, gf, taken_f = grads(True)
    loss_h, gh, taken_h = grads(False)
    assert taken_f == cfg.decoder_layers and taken_h == 0, whisper_train.REFUSED
    assert abs(loss_f - loss_h) <= 1e-2 * abs(loss_h)
    for n, a in gh.items():
        assert rel_err(gf[n], a) <= 1.2e-1, n
    # and the masks do matter: without dropout the gradients are different
    for m in sar.lora_modules(w.model).values():
        m.lora_dropout["default"].p = 0.0
    _, g0, _ = grads(True)
    assert max(rel_err(g0[n], gf[n]) for n in gf) > 2e-1
