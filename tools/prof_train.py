"""Where a training step goes: torch.profiler over one BASELINE-config-5 step (whisper-small r16, 16 clips, checkpointing)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("SAR_RANDOM_INIT", "1")
import speech_adapter_routing_b200 as sar  # noqa: E402
from speech_adapter_routing_b200.dist import FlatGradBucket  # noqa: E402

dev = torch.device("cuda")
w = sar.WhisperLoRA("whisper-small", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda", use_gradient_checkpointing=True)
w.train()
cfg = w.model.config
params = [p for p in w.model.parameters() if p.requires_grad]
bucket = FlatGradBucket(params)
g = torch.Generator().manual_seed(1)
x = torch.randn(16, cfg.num_mel_bins, 3000, generator=g).to(dev).to(torch.bfloat16)
labels = torch.randint(5, cfg.vocab_size, (16, 128), generator=g).to(dev)


def step():
    bucket.zero_()
    w(input_features=x, labels=labels).loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
