"""Log-mel front-end on the GPU: waveform → Whisper ``input_features`` (libsar ``sar_logmel_fwd``).

Drop-in for the per-example CPU call of the reference's data path,
``processor.feature_extractor(audio_array, sampling_rate=16000, return_tensors="pt").input_features``
(src/data/dataset.py:124-128; transformers' WhisperFeatureExtractor), for a whole batch at once: pad or cut every
clip to 30 s, 400-sample periodic-Hann frames every 160 samples, |DFT|², Slaney mel filterbank, log10, clamp to the
clip maximum − 8, (x + 4) / 4.  No CPU fallback: the kernels need a B200.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple, Union

import torch

from . import ops

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_SAMPLES = 30 * SAMPLE_RATE

_TABLES: Dict[Tuple[str, int], Tuple[torch.Tensor, ...]] = {}


def _hz_to_mel(f: torch.Tensor) -> torch.Tensor:
    # Slaney: linear below 1 kHz (200/3 Hz per mel), logarithmic above (27 mels per factor 6.4)
    lin = 3.0 * f / 200.0
    log = 15.0 + torch.log(f.clamp_min(1e-30) / 1000.0) * (27.0 / math.log(6.4))
    return torch.where(f >= 1000.0, log, lin)


def _mel_to_hz(m: torch.Tensor) -> torch.Tensor:
    lin = 200.0 * m / 3.0
    log = 1000.0 * torch.exp((math.log(6.4) / 27.0) * (m - 15.0))
    return torch.where(m >= 15.0, log, lin)


def mel_filterbank(n_mels: int, n_fft: int = N_FFT, sample_rate: int = SAMPLE_RATE) -> torch.Tensor:
    """[n_fft/2 + 1, n_mels] float64: triangular filters on the Slaney mel scale with Slaney (area) normalisation — the
    matrix WhisperFeatureExtractor holds as ``mel_filters``."""
    bins = n_fft // 2 + 1
    freqs = torch.linspace(0.0, sample_rate / 2.0, bins, dtype=torch.float64)
    top = _hz_to_mel(torch.tensor([0.0, sample_rate / 2.0], dtype=torch.float64))
    edges = _mel_to_hz(torch.linspace(float(top[0]), float(top[1]), n_mels + 2, dtype=torch.float64))
    width = edges[1:] - edges[:-1]
    rel = edges[None, :] - freqs[:, None]
    rising = -rel[:, :-2] / width[:-1]
    falling = rel[:, 2:] / width[1:]
    tri = torch.minimum(rising, falling).clamp_min(0.0)
    return tri * (2.0 / (edges[2:] - edges[:-2]))[None, :]


def tables(device: Union[str, torch.device], n_mels: int) -> Tuple[torch.Tensor, ...]:
    """(window, cos, sin, mel_filters) as fp32 device tensors, computed once per (device, n_mels) in float64."""
    key = (str(device), n_mels)
    if key not in _TABLES:
        j = torch.arange(N_FFT, dtype=torch.float64)
        ang = 2.0 * math.pi * j / N_FFT
        host = (0.5 - 0.5 * torch.cos(ang), torch.cos(ang), torch.sin(ang), mel_filterbank(n_mels))
        _TABLES[key] = tuple(t.to(torch.float32).contiguous().to(device) for t in host)
    return _TABLES[key]


def pad_or_trim(waves: Union[torch.Tensor, Sequence[torch.Tensor]], n_samples: int = N_SAMPLES) -> torch.Tensor:
    """A [B, n] tensor or a list of 1-D clips of different lengths → fp32 [B, n_samples], zero-padded on the right / cut
    (WhisperFeatureExtractor pads to max_length = 30 s with zeros and truncates)."""
    if isinstance(waves, torch.Tensor):
        waves = [waves] if waves.dim() == 1 else list(waves)
    out = torch.zeros(len(waves), n_samples, dtype=torch.float32, device=waves[0].device)
    for i, w in enumerate(waves):
        n = min(int(w.shape[-1]), n_samples)
        out[i, :n] = w[:n].to(torch.float32)
    return out


def log_mel_spectrogram(waves: Union[torch.Tensor, Sequence[torch.Tensor]], n_mels: int = 80,
                        dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """16 kHz waveforms on a CUDA device → ``input_features`` [B, n_mels, 3000] (80 mel bins for whisper-small / medium,
    128 for large-v3), ready for ``WhisperLoRA`` / ``AdapterRouter``."""
    x = pad_or_trim(waves)
    if not x.is_cuda:
        raise RuntimeError("log_mel_spectrogram: CUDA tensors required (libsar has no CPU fallback)")
    window, cos_t, sin_t, filt = tables(x.device, n_mels)
    return ops.logmel_fwd(x, window, cos_t, sin_t, filt, out_dtype=dtype)
