"""Rebuild the inputs of a tests/golden/routed_forward_golden.pt case from its seeds (the golden file stores outputs of
the REFERENCE's AdapterRouter run in the build container plus checksums of the regenerated inputs, not the inputs)."""
from pathlib import Path

import torch

from oracle import fixtures, whisper as owhisper

GOLDEN = Path(__file__).resolve().parent / "golden" / "routed_forward_golden.pt"


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def load_golden():
    return torch.load(GOLDEN, weights_only=True)


class Case:
    """Model, adapters, router head, clips and targets of one golden case — regenerated, then verified against the stored
    checksums so that a drift of the seeded generators cannot masquerade as a parity failure."""

    def __init__(self, rec):
        self.rec = rec
        C, r = rec["C"], rec["r"]
        self.whisper = owhisper.build_whisper(rec["geometry"])
        self.cfg = self.whisper.config
        self.weights = owhisper.make_adapter_weights(self.whisper, r, C)
        sd0 = fixtures.make_router_state_dict(self.cfg.d_model, C)
        protos = owhisper.make_input_features(C, self.cfg.num_mel_bins, list(range(C)), C, seed=99)
        with torch.no_grad():
            feats = self.whisper.model.encoder(protos).last_hidden_state
        self.router_sd = owhisper.fit_router_head(sd0, feats)
        seed, B, T_dec = rec["seed"], rec["B"], rec["T_dec"]
        langs = fixtures.language_mix(B, C, rec["mix"], seed=7 + seed)
        assert langs == rec["langs"]
        self.x = owhisper.make_input_features(B, self.cfg.num_mel_bins, langs, C, seed=2234 + seed)
        self.dec, labels = owhisper.make_decoder_inputs(B, T_dec, self.cfg.vocab_size, self.cfg.decoder_start_token_id)
        self.labels = labels.clone()
        self.labels[1::2, -3:] = -100
        rel = lambda a, b: abs(a - b) <= 1e-9 * abs(b)
        assert rel(checksum(self.x), rec["x_checksum"]), "input generator drifted"
        assert rel(sum(checksum(p) for p in self.whisper.parameters()), rec["weights_checksum"]), "weight init drifted"
        assert rel(sum(checksum(A) + checksum(Bm) for A, Bm in self.weights.values()), rec["adapter_checksum"])
        assert rel(sum(checksum(v) for v in self.router_sd.values()), rec["router_sd_checksum"])

    def oracle(self):
        return owhisper.RoutedWhisperOracle(self.whisper, self.weights, self.rec["r"], 2 * self.rec["r"], self.router_sd)
