"""Times the GPU log-mel front-end (sar_logmel_fwd) for a 64-clip batch of 30 s waveforms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_adapter_routing_b200 import logmel

x = torch.randn(64, logmel.N_SAMPLES, device="cuda") * 0.1
for n_mels in (80, 128):
    for _ in range(2):
        logmel.log_mel_spectrogram(x, n_mels=n_mels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        logmel.log_mel_spectrogram(x, n_mels=n_mels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"log-mel 64 clips x 30 s, {n_mels} mel bins: {ms:.2f} ms ({64 / ms * 1e3:.0f} clips/s, {123.5 / ms:.1f} TFLOP/s fp32 in the DFT)")
