import sys; sys.path.insert(0, "/root/repo")
import torch
from oracle import fixtures
from speech_adapter_routing_b200 import ops
for d in (768, 1280):
    sd = fixtures.make_router_state_dict(d, 4)
    p = ops.RouterParams.from_state_dict(sd, "cuda:0")
    hs = [torch.randn(64, 1500, d, device="cuda:0", dtype=torch.bfloat16) for _ in range(3)]
    for i in range(6):
        ops.router_fwd(hs[i % 3], p)
torch.cuda.synchronize()
print("ok")
