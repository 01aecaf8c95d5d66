"""ncu target: the fused LayerNorm + LoRA-down kernel (and the plain LayerNorm / U pass beside it) at the bench shape."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_adapter_routing_b200 import ops  # noqa: E402

B, T, d, r, n = 64, 1500, 768, 16, 4
dev = "cuda"
g = torch.Generator().manual_seed(1)
h = torch.randn(B, T, d, device=dev, dtype=torch.bfloat16)
gw = torch.ones(d, device=dev, dtype=torch.bfloat16)
gb = torch.zeros(d, device=dev, dtype=torch.bfloat16)
A = (torch.randn(2 * n, r, d, device=dev) * 0.03).to(torch.bfloat16)
ia = torch.randint(0, n, (B,), generator=g).to(torch.int32).to(dev)
none = torch.full((B,), -1, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.layernorm_fwd(h, gw, gb, 1e-5)
    x, u = ops.layernorm_lora_u_fwd(h, gw, gb, A, ia, 2, 2.0)
    ops.layernorm_lora_u_fwd(h, gw, gb, A, none, 2, 2.0)     # same kernel, every utterance base-only: LayerNorm part alone
    ops.lora_u_fwd(x, A, ia, 2, 2.0, d)
torch.cuda.synchronize()
