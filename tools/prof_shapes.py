"""ncu driver: one launch per kernel / shape of the routed step (after one warm-up launch of each) — see
tools/profile_all.sh, which captures the SECOND round of launches with `ncu --set full`.

Order of the launches in a round (NAMES below, one per launch):
"""
NAMES = ["plain 96000x768->768", "fc1 96000x768->3072", "fc1+GELU (8 epilogue warps)",
         "out_proj head-major x + residual (8 epilogue warps)",
         "LayerNorm + U fused (sar_layernorm_lora_u_fwd: x and U = scale·x·A_kT for q and v in one pass)",
         "q|k|v + routed LoRA r16: dense 256-wide tiles + extra K block, U ready (north-star kernel)",
         "fc2 96000x3072->768 + residual", "plain LayerNorm",
         "q|k|v + routed LoRA r16, two-launch form: U pass", "q|k|v + routed LoRA r16, two-launch form: dense launch",
         "q|k|v base only (LID pass)", "K2 pool (router, bf16 states)", "K2 head"]
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_adapter_routing_b200 import ops  # noqa: E402

DEV = "cuda:0"
B, T, d, ffn, r, n = 64, 1500, 768, 3072, 16, 4
g = torch.Generator().manual_seed(1)
x = torch.randn(B, T, d, device=DEV, dtype=torch.bfloat16)
x2 = torch.randn(B, T, d, device=DEV, dtype=torch.bfloat16)
f = torch.randn(1, B * T, ffn, device=DEV, dtype=torch.bfloat16)
W = (torch.randn(d, d, device=DEV) * 0.02).to(torch.bfloat16)
W1 = (torch.randn(ffn, d, device=DEV) * 0.02).to(torch.bfloat16)
W2 = (torch.randn(d, ffn, device=DEV) * 0.02).to(torch.bfloat16)
bd = torch.zeros(d, device=DEV, dtype=torch.bfloat16)
bf = torch.zeros(ffn, device=DEV, dtype=torch.bfloat16)
Wqkv = (torch.randn(3 * d, d, device=DEV) * 0.02).to(torch.bfloat16)
bqkv = torch.zeros(3 * d, device=DEV, dtype=torch.bfloat16)
A = (torch.randn(2 * n, r, d, device=DEV) * 0.03).to(torch.bfloat16)
Bp = ops.pack_lora_b((torch.randn(2 * n, d, r, device=DEV) * 0.02).to(torch.bfloat16))
ia = torch.randint(0, n, (B,), generator=g).to(torch.int32).to(DEV)
xf = x.view(1, B * T, d)
xh = x.view(B, d // 64, T, 64)


from oracle import fixtures  # noqa: E402  (router parameters with the reference's key names)

gam = torch.ones(d, device=DEV, dtype=torch.bfloat16)
rp = ops.RouterParams.from_state_dict(fixtures.make_router_state_dict(d, n), DEV)


def all_shapes():
    ops.linear_fwd(xf, W, bd)
    ops.linear_fwd(xf, W1, bf)
    ops.linear_fwd(xf, W1, bf, None, 1)
    ops.linear_fwd(xh, W, bd, x2, 0, x_head_major=True)
    xn, u = ops.layernorm_lora_u_fwd(x2, gam, bd, A, ia, 2, 2.0)
    ops.attn_proj_fwd(xn, Wqkv, bqkv, A, Bp, ia, [0, -1, 1], [1, 1, 1], 2, 2.0, u=u)
    ops.linear_fwd(f, W2, bd, x2.view(1, B * T, d))
    ops.layernorm_fwd(x2, gam, bd)
    ops.attn_proj_fwd(x, Wqkv, bqkv, A, Bp, ia, [0, -1, 1], [1, 1, 1], 2, 2.0, split=True)
    ops.attn_proj_fwd(x, Wqkv, bqkv, None, None, None, [-1, -1, -1], [1, 1, 1], 1, 2.0)
    ops.router_fwd(x, rp)


all_shapes()
torch.cuda.synchronize()
all_shapes()
torch.cuda.synchronize()
print("done")
