"""ncu driver: one launch of the pair kernel per Whisper-block GEMM shape (after one warm-up launch of each).

    ncu --set full --clock-control none --import-source on -k regex:k1v2 -s 8 -c 8 -o gpurun_out/prof python tools/prof_shapes.py
Order of the profiled launches: plain 768->768 | fc1 768->3072 | fc1+GELU | out_proj head-major+residual |
q|k|v + LoRA split path (U pass, then dense AUG kernel) | fc2 3072->768 + residual | q|k|v + LoRA single launch.
"""
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


def all_shapes():
    ops.linear_fwd(xf, W, bd)
    ops.linear_fwd(xf, W1, bf)
    ops.linear_fwd(xf, W1, bf, None, 1)
    ops.linear_fwd(xh, W, bd, x2, 0, x_head_major=True)
    ops.attn_proj_fwd(x, Wqkv, bqkv, A, Bp, ia, [0, -1, 1], [1, 1, 1], 2, 2.0)
    ops.linear_fwd(f, W2, bd, x2.view(1, B * T, d))
    ops.attn_proj_fwd(x, Wqkv, bqkv, A, Bp, ia, [0, -1, 1], [1, 1, 1], 2, 2.0, split=False)


all_shapes()
torch.cuda.synchronize()
all_shapes()
torch.cuda.synchronize()
print("done")
