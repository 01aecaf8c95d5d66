"""ncu target: the router (K2) alone at the bench shape, plain and with the encoder's final LayerNorm folded in."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import fixtures  # noqa: E402
from speech_adapter_routing_b200 import ops  # noqa: E402

B, T, d, C = 64, 1500, int(sys.argv[1]) if len(sys.argv) > 1 else 768, 4
sd = fixtures.make_router_state_dict(d, C)
p = ops.RouterParams.from_state_dict(sd, "cuda")
hs = [torch.randn(B, T, d, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
g = torch.ones(d, device="cuda", dtype=torch.bfloat16)
b = torch.zeros(d, device="cuda", dtype=torch.bfloat16)
for i in range(4):
    ops.router_fwd(hs[i % 2], p)
    ops.router_fwd(hs[i % 2], p, pre_ln=(g, b, 1e-5))
torch.cuda.synchronize()
