"""Golden vectors of the ROUTED FORWARD from the reference's own ``AdapterRouter`` (run in the build container only).

    python tests/golden/make_routed_golden.py        # reads /root/reference, writes routed_forward_golden.pt

/root/reference/src/models/adapter_router.py imports only torch, so it loads by file path without ``peft`` (the
``src.models`` package itself does not import here).  This script instantiates the UNMODIFIED reference classes

    LanguageClassifier (:14)   EncoderFeatureExtractor (:392)   AdapterRouter (:488)

on the CPU in fp32 and records what ``forward`` (hard :599-625, soft :627-670, threshold :672-693, loss aggregation
:695-713), ``detect_language`` (:550-566) and ``generate`` (:715-761, with and without ``language=``) return.

What stands where the reference puts its third-party pieces:
  * ``base_model``  = installed HF ``WhisperForConditionalGeneration`` (transformers 5.5.0; the reference's own
    dependency), random-init from ``oracle.whisper.build_whisper`` (seeded; no hub access).
  * ``adapters[lang]`` = ``AdapterStandIn``: a separate full copy of that model per language — as the reference keeps
    one ``WhisperLoRA`` per language (:518) — whose q_proj / v_proj apply PEFT's published formula
    (``oracle.lora.lora_linear``; ``peft`` itself is not installable offline) and whose ``forward`` / ``generate``
    forward exactly the keyword arguments ``WhisperLoRA`` forwards (src/models/whisper_lora.py:137-143, :172-177).

Only seeds, integer outputs, small float outputs and checksums are stored: inputs and weights are regenerated from the
seeds by ``oracle.whisper`` (same image on the GPU box), and the checksums catch any drift.
"""
from __future__ import annotations

import copy
import importlib.util
import sys
from pathlib import Path

import torch
import torch.nn as nn

REF = Path("/root/reference/src/models/adapter_router.py")
OUT = Path(__file__).resolve().parent / "routed_forward_golden.pt"

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import fixtures, lora as olora, router as orouter, whisper as owhisper  # noqa: E402


def load_reference_module():
    spec = importlib.util.spec_from_file_location("ref_adapter_router", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _FixedLoRALinear(nn.Module):
    """PEFT ``lora.Linear`` for ONE adapter: base(x) + B(A(x)) * (alpha / r)."""

    def __init__(self, base: nn.Linear, A: torch.Tensor, B: torch.Tensor, scaling: float):
        super().__init__()
        self.base_layer = base
        self.register_buffer("A", A.clone())
        self.register_buffer("B", B.clone())
        self.scaling = scaling

    def forward(self, x):
        return olora.lora_linear(x, self.base_layer.weight, self.base_layer.bias, self.A, self.B, self.scaling)


class AdapterStandIn(nn.Module):
    """Where the reference holds a ``WhisperLoRA`` (one full model per language)."""

    def __init__(self, whisper, weights, k: int, r: int, lora_alpha: float):
        super().__init__()
        self.model = copy.deepcopy(whisper)
        for path, (A, B) in weights.items():
            parent = self.model.get_submodule(path.rsplit(".", 1)[0])
            leaf = path.rsplit(".", 1)[-1]
            setattr(parent, leaf, _FixedLoRALinear(getattr(parent, leaf), A[k], B[k], lora_alpha / r))

    def forward(self, input_features, labels=None, attention_mask=None, decoder_input_ids=None,
                decoder_attention_mask=None, **kwargs):            # whisper_lora.py:114-143: other kwargs are swallowed
        return self.model(input_features=input_features, labels=labels, attention_mask=attention_mask,
                          decoder_input_ids=decoder_input_ids, decoder_attention_mask=decoder_attention_mask)

    def generate(self, input_features, max_new_tokens=256, num_beams=1, language=None, task=None, **kwargs):
        return self.model.generate(input_features=input_features, max_new_tokens=max_new_tokens, num_beams=num_beams,
                                   **kwargs)                       # whisper_lora.py:172-177


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def build_case(ref, geometry: str, C: int, r: int, B: int, T_dec: int, mix: str, seed: int, gen_steps: int):
    whisper = owhisper.build_whisper(geometry)
    cfg = whisper.config
    weights = owhisper.make_adapter_weights(whisper, r, C)
    languages = [f"lang{k}" for k in range(C)]

    # router head: reference class, default architecture, output layer fitted on one prototype clip per language
    sd0 = fixtures.make_router_state_dict(cfg.d_model, C)
    protos = owhisper.make_input_features(C, cfg.num_mel_bins, list(range(C)), C, seed=99)
    with torch.no_grad():
        proto_feats = whisper.model.encoder(protos).last_hidden_state
    sd = owhisper.fit_router_head(sd0, proto_feats)
    clf = ref.LanguageClassifier(input_dim=cfg.d_model, num_classes=C, languages=languages)
    clf.load_state_dict(sd)
    clf.eval()

    adapters = {lang: AdapterStandIn(whisper, weights, k, r, 2 * r).eval() for k, lang in enumerate(languages)}
    router = ref.AdapterRouter(base_model=whisper, adapters=adapters, classifier=clf, languages=languages,
                               strategy="hard").eval()

    langs = fixtures.language_mix(B, C, mix, seed=7 + seed)
    x = owhisper.make_input_features(B, cfg.num_mel_bins, langs, C, seed=2234 + seed)
    dec, labels = owhisper.make_decoder_inputs(B, T_dec, cfg.vocab_size, cfg.decoder_start_token_id)
    labels_pad = labels.clone()
    labels_pad[1::2, -3:] = -100                      # ragged targets: per-utterance token means differ from the batch mean

    rec = {"geometry": geometry, "C": C, "r": r, "B": B, "T_dec": T_dec, "mix": mix, "seed": seed,
           "languages": languages, "langs": langs, "gen_steps": gen_steps,
           "x_checksum": checksum(x), "weights_checksum": sum(checksum(p) for p in whisper.parameters()),
           "adapter_checksum": sum(checksum(A) + checksum(Bm) for A, Bm in weights.values()),
           "router_sd_checksum": sum(checksum(v) for v in sd.values())}
    with torch.no_grad():
        h = router.extract_encoder_features(x)
        names, probs = router.detect_language(h)
        rec["lid_features_checksum"] = checksum(h)
        rec["lid_features_row0"] = h[:, 0, :].clone()
        rec["names"] = names
        rec["probs"] = probs.clone()
        rec["idx"] = torch.tensor([languages.index(n) for n in names])

        out = router(x, labels=labels_pad)                                   # hard, labels only
        rec["hard_labels"] = {"logits": out["logits"].clone(), "loss": out["loss"].clone()}
        # NOT recorded: hard routing with a batch-shaped ``decoder_input_ids`` keyword.  The reference does not slice
        # ``**kwargs`` per utterance (:618-622), so each batch-1 forward gets the whole [B, T] tensor; HF then reshapes the
        # ONE clip's encoder K/V into B chunks of 1500/B frames (``.view(bsz, -1, h, hd)``), giving B*B rows of logits that
        # depend on B (and an exception whenever B does not divide 1500).  Only ``labels`` are sliced per utterance, so
        # labels-only calls are the reference's well-defined hard-routing contract.
        rec["reference_kwargs_quirk"] = True

        router.strategy = "soft"
        out = router(x, labels=labels_pad)
        rec["soft_labels"] = {"logits": out["logits"].clone(), "loss": out["loss"].clone(), "probs": out["probs"].clone()}

        router.strategy = "threshold"
        router.threshold = 0.5
        out = router(x, labels=labels_pad)
        # the threshold strategy returns exactly what hard / soft return: store which, and that it was bit-identical
        rec["threshold_0p5"] = {"loss": out["loss"].clone(), "keys": sorted(out), "threshold": 0.5, "same_as": "hard_labels",
                                "bit_identical": bool(torch.equal(out["logits"], rec["hard_labels"]["logits"]))}
        router.threshold = 1.0 - 1e-12                                      # nobody is that confident -> soft branch
        out = router(x, labels=labels_pad)
        rec["threshold_1m"] = {"loss": out["loss"].clone(), "keys": sorted(out), "threshold": router.threshold,
                               "same_as": "soft_labels",
                               "bit_identical": bool(torch.equal(out["logits"], rec["soft_labels"]["logits"]))}
        router.strategy = "hard"

        ids = router.generate(x, max_new_tokens=gen_steps, num_beams=1, do_sample=False)
        rec["generate"] = {"ids": ids.clone()}
        # EOS handling + the reference's zero right-padding (:753-761): random-init models never emit the real EOS, so declare
        # a token that one row reaches after its first position to be EOS for a second run (HF ``generate(eos_token_id=)``)
        def first_new_token(t_ids):
            for t in range(1, t_ids.shape[1]):
                for i in range(t_ids.shape[0]):
                    if t_ids[i, t] != t_ids[i, t - 1] and not (t_ids[:, :t] == t_ids[i, t]).any():
                        return int(t_ids[i, t])
            return None

        eos_tok = first_new_token(ids)
        if eos_tok is not None:
            ids_e = router.generate(x, max_new_tokens=gen_steps, num_beams=1, do_sample=False, eos_token_id=eos_tok)
            rec["generate_eos"] = {"ids": ids_e.clone(), "eos_token_id": eos_tok}
        ids_l = router.generate(x, language=languages[1], max_new_tokens=gen_steps, num_beams=1, do_sample=False)
        rec["generate_language"] = {"ids": ids_l.clone(), "language": languages[1]}
        eos_l = first_new_token(ids_l)
        if eos_l is not None:
            ids_le = router.generate(x, language=languages[1], max_new_tokens=gen_steps, num_beams=1, do_sample=False,
                                     eos_token_id=eos_l)
            rec["generate_language_eos"] = {"ids": ids_le.clone(), "eos_token_id": eos_l, "language": languages[1],
                                            "pad_token_id": int(cfg.pad_token_id)}

        # top-1 / top-2 logit margin of every generated position (teacher-forced on the reference's own tokens): the GPU
        # test compares tokens at every position whose margin is above the bf16 noise floor
        margins = torch.zeros(ids.shape, dtype=torch.float32)
        absmax = 0.0
        start = torch.full((1, 1), cfg.decoder_start_token_id)
        for i in range(B):
            seq = ids[i:i + 1]
            dec_in = torch.cat([start, seq[:, :-1]], 1)
            lg = adapters[names[i]](input_features=x[i:i + 1], decoder_input_ids=dec_in).logits[0]
            margins[i] = orouter.top2_margin(lg)
            absmax = max(absmax, float(lg.abs().max()))
        rec["generate"]["margins"] = margins
        rec["generate"]["logit_absmax"] = absmax
    print(f"{geometry} B={B} mix={mix}: idx={rec['idx'].tolist()} min max-prob={probs.max(-1).values.min():.4f} "
          f"hard loss={rec['hard_labels']['loss']:.5f} soft loss={rec['soft_labels']['loss']:.5f} "
          f"gen shape={tuple(ids.shape)} min margin={margins.min():.4f}")
    return rec


def main():
    ref = load_reference_module()
    torch.manual_seed(0)
    cases = [build_case(ref, "micro", 4, 16, 6, 8, "uniform", 0, 12),
             build_case(ref, "micro", 4, 16, 5, 6, "skewed", 1, 8)]
    torch.save({"source": str(REF), "torch": str(torch.__version__), "threads": torch.get_num_threads(),
                "cpu_capability": torch.backends.cpu.get_cpu_capability(), "cases": cases}, OUT)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
