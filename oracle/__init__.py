"""oracle/ — CPU restatement of the reference's routed multi-adapter LoRA hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker / timed CPU baseline.  The product package
(``speech_adapter_routing_b200``) never imports it and has no CPU fallback.

Parity pinning
--------------
* Router (``oracle.router``): PINNED.  The restatement is checked against golden vectors produced by the
  reference's own ``LanguageClassifier`` (``/root/reference/src/models/adapter_router.py``, imported by file
  path) — see ``tests/golden/make_golden.py`` and ``tests/golden/router_golden.pt``.
* LoRA linear (``oracle.lora``): PARITY UNPINNED by the reference.  The arithmetic lives in the third-party
  ``peft`` package (reference pins only ``peft>=0.7.0``, requirements.txt:6; no lockfile; not installed here and
  not installable offline).  The restatement follows PEFT's published ``lora.Linear.forward``
  (``result = base_layer(x) + lora_B(lora_A(dropout(x))) * scaling``, ``scaling = lora_alpha / r``) and the
  reference's call sites (src/models/whisper_lora.py:88-98).  The reference ships no expected tensors for it
  (SURVEY.md §8c).
* Routed Whisper forward (``oracle.whisper``): HF ``transformers`` Whisper (installed 5.5.0) + ``oracle.lora``
  modules at every q_proj/v_proj + the per-utterance hard-routing loop of
  src/models/adapter_router.py:599-625.  Parity unpinned by the reference (no forward-pass numbers exist).
"""
