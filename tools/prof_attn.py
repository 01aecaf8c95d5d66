"""ncu target: encoder self-attention (B*h = 768, 1500 x 1500, head dim 64) on libsar's attention kernel."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_adapter_routing_b200 import ops  # noqa: E402

q = (torch.randn(64, 12, 1500, 64, device="cuda") * 0.35).to(torch.bfloat16)
k = torch.randn(64, 12, 1500, 64, device="cuda").to(torch.bfloat16)
v = torch.randn(64, 12, 1500, 64, device="cuda").to(torch.bfloat16)
for _ in range(3):
    ops.attn_fwd(q, k, v, False)
torch.cuda.synchronize()
