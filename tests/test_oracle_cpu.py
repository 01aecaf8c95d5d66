"""CPU tests of the oracle itself: the router restatement is pinned to golden vectors produced by the REFERENCE's own
LanguageClassifier (tests/golden/make_golden.py); the LoRA restatement is checked for internal consistency (the
reference holds no vectors for it — parity unpinned, see oracle/__init__.py)."""
from pathlib import Path

import torch

from oracle import fixtures, lora as olora, router as orouter

GOLDEN = Path(__file__).parent / "golden" / "router_golden.pt"


def test_router_oracle_matches_reference_golden_vectors():
    g = torch.load(GOLDEN)
    assert g["source"].endswith("src/models/adapter_router.py")
    assert len(g["cases"]) >= 4
    for c in g["cases"]:
        out = orouter.classifier_forward(c["h"], c["state_dict"])
        # same torch ops in the same order as the reference class: bit-identical on the same build
        assert torch.allclose(out["logits"], c["logits"], rtol=0, atol=1e-6)
        assert torch.allclose(out["probs"], c["probs"], rtol=0, atol=1e-7)
        labels, probs = orouter.predict(c["h"], c["state_dict"])
        assert torch.equal(labels, c["labels"])
        assert labels.dtype == torch.int64
        assert torch.allclose(probs, c["predict_probs"], rtol=0, atol=1e-7)


def test_router_golden_state_dict_keys_are_the_documented_ones():
    g = torch.load(GOLDEN)
    keys = set(g["cases"][0]["state_dict"])
    assert keys == set(fixtures.ROUTER_KEYS)


def test_segments_is_a_stable_counting_sort():
    idx = torch.tensor([2, 0, 2, 1, 0, 2, 3, 0])
    perm, seg = orouter.segments(idx, 5)
    assert perm.tolist() == [1, 4, 7, 3, 0, 2, 5, 6]
    assert seg.tolist() == [0, 3, 4, 7, 8, 8]
    perm, seg = orouter.segments(torch.tensor([1, 1, 1]), 3)   # single language
    assert perm.tolist() == [0, 1, 2] and seg.tolist() == [0, 0, 3, 3]


def test_lora_routed_equals_per_utterance_batch1_loop():
    c = fixtures.make_lora_case(5, 9, 32, 48, 8, 3, base_only_every=2)
    x, W, b = c.x.float(), c.W.float(), c.bias.float()
    A, B = c.A_stack.float(), c.B_stack.float()
    y = olora.lora_linear_routed(x, W, b, A, B, c.scaling, c.utt_adapter)
    for i in range(5):
        k = int(c.utt_adapter[i])
        base = x[i] @ W.t() + b
        want = base if k < 0 else base + c.scaling * (x[i] @ A[k].t()) @ B[k].t()
        assert torch.allclose(y[i], want, atol=1e-5)


def test_lora_zero_B_is_identity_to_base():
    """PEFT's default init (lora_B = 0) leaves the base projection unchanged."""
    c = fixtures.make_lora_case(2, 4, 16, 16, 4, 1, mix="single")
    y = olora.lora_linear(c.x.float(), c.W.float(), c.bias.float(), c.A_stack[0].float(),
                          torch.zeros_like(c.B_stack[0]).float(), c.scaling)
    assert torch.equal(y, torch.nn.functional.linear(c.x.float(), c.W.float(), c.bias.float()))


def test_k1_rounding_oracle_is_close_to_fp32_oracle():
    c = fixtures.make_lora_case(3, 20, 64, 64, 16, 2)
    a = olora.lora_linear_routed(c.x.float(), c.W.float(), c.bias.float(), c.A_stack.float(), c.B_stack.float(),
                                 c.scaling, c.utt_adapter)
    b = olora.lora_linear_routed_k1_rounding(c.x, c.W, c.bias, c.A_stack, c.B_stack, c.scaling, c.utt_adapter)
    assert (a - b.float()).abs().max() <= 2 ** -7 * a.abs().max()


def test_lora_backward_matches_closed_form():
    c = fixtures.make_lora_case(3, 7, 16, 24, 4, 2, base_only_every=3)
    dy = torch.randn(3, 7, 24, generator=torch.Generator().manual_seed(0))
    x, W, A, B = c.x.float(), c.W.float(), c.A_stack.float(), c.B_stack.float()
    dx, dA, dB = olora.lora_linear_backward(dy, x, W, A, B, c.scaling, c.utt_adapter)
    s = c.scaling
    dA_ref, dB_ref = torch.zeros_like(A), torch.zeros_like(B)
    for i in range(3):
        k = int(c.utt_adapter[i])
        want_dx = dy[i] @ W
        if k >= 0:
            want_dx = want_dx + s * (dy[i] @ B[k]) @ A[k]
            dA_ref[k] += s * (dy[i] @ B[k]).t() @ x[i]
            dB_ref[k] += s * dy[i].t() @ (x[i] @ A[k].t())
        assert torch.allclose(dx[i], want_dx, atol=1e-5)
    assert torch.allclose(dA, dA_ref, atol=1e-4) and torch.allclose(dB, dB_ref, atol=1e-4)


def test_language_mix_fixture_covers_every_class():
    for kind in ("uniform", "skewed"):
        ids = fixtures.language_mix(64, 4, kind)
        counts = [ids.count(k) for k in range(4)]
        assert min(counts) >= 0.09 * 64, counts     # no class under ~10 % in the mixed fixtures
    assert set(fixtures.language_mix(8, 4, "single")) == {3}
