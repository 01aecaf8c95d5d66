"""CPU oracle of Whisper's log-mel front-end (TEST INFRASTRUCTURE ONLY — imported by tests/, never by the product path).

Restates, in float64 numpy, what the reference's data path computes per clip before the model ever runs:
src/data/dataset.py:124-128 calls ``processor.feature_extractor(audio_array, sampling_rate=sr, return_tensors="pt")``,
i.e. transformers' WhisperFeatureExtractor ($HF/models/whisper/feature_extraction_whisper.py:105-135 → audio_utils
``spectrogram`` / ``mel_filter_bank``): pad or cut to 30 s, reflect-pad n_fft/2, 400-sample periodic-Hann frames every 160
samples, |rFFT|², Slaney-scale / Slaney-normalised mel filterbank, log10 with a 1e-10 floor, drop the last frame, clamp to
(clip maximum − 8), (x + 4) / 4.
PARITY: pinned against the installed WhisperFeatureExtractor itself in tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_SAMPLES = 480000          # 30 s
N_FRAMES = N_SAMPLES // HOP  # 3000


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    log_region = f >= 1000.0
    mel = np.where(log_region, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4)), mel)
    return mel


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    log_region = m >= 15.0
    return np.where(log_region, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)


def mel_filterbank(n_mels: int, n_fft: int = N_FFT, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """[n_fft/2 + 1, n_mels] triangular filters, Slaney mel scale, Slaney (area) normalisation, 0 .. sr/2
    (audio_utils.mel_filter_bank as WhisperFeatureExtractor.__init__ calls it)."""
    n_bins = n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, sample_rate / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(sample_rate / 2.0), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]                  # [bins, n_mels + 2]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    return fb * enorm[None, :]


def pad_or_trim(wave: np.ndarray, n_samples: int = N_SAMPLES) -> np.ndarray:
    wave = np.asarray(wave, dtype=np.float64)
    if wave.shape[-1] >= n_samples:
        return wave[..., :n_samples]
    out = np.zeros(wave.shape[:-1] + (n_samples,), dtype=np.float64)
    out[..., :wave.shape[-1]] = wave
    return out


def log_mel(wave: np.ndarray, n_mels: int = 80) -> np.ndarray:
    """One clip [n] (any length; padded / cut to 30 s) → [n_mels, 3000] float64."""
    x = pad_or_trim(wave)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n = np.arange(N_FFT)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)            # periodic Hann
    idx = np.arange(N_FRAMES)[:, None] * HOP + n[None, :]           # the 3001st frame is dropped by the reference
    frames = xp[idx] * window[None, :]
    k = np.arange(N_FFT // 2 + 1)
    ang = 2.0 * np.pi * (np.outer(n, k) % N_FFT) / N_FFT
    re = frames @ np.cos(ang)
    im = frames @ np.sin(ang)
    power = re * re + im * im                                        # [frames, bins]
    mel = power @ mel_filterbank(n_mels)                             # [frames, n_mels]
    log_spec = np.log10(np.maximum(mel, 1e-10)).T                    # [n_mels, frames]
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0
