"""K3 (sar_qv_lora_bwd) alone at the training shapes of BASELINE config 5 (whisper-small, 16 clips): the three launches of
one call — dX on the pair kernel (U pass + dense launch with the low-rank K block), the skinny dA/dB reduction, the
partial reduce — with CUDA-event time of the whole call and the algorithmic work of each piece.  Run under
`ncu --metrics gpu__time_duration.sum` for the per-launch split (tools/profile_all.sh, profiles/r02_k3_launches.csv)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_adapter_routing_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
for name, B, T, d, r, need_dx in (("encoder q/v (dx needed)", 16, 1500, 768, 16, True),
                                  ("encoder layer 0 (no dx)", 16, 1500, 768, 16, False),
                                  ("decoder self q/v", 16, 128, 768, 16, True)):
    M = B * T
    x = torch.randn(B, T, d, device=dev).to(torch.bfloat16)
    dy = (torch.randn(B, T, d, device=dev) * 0.1).to(torch.bfloat16)
    W = (torch.randn(d, d, device=dev) * 0.03).to(torch.bfloat16)
    A = (torch.randn(1, r, d, device=dev) * 0.05).to(torch.bfloat16)
    Bm = (torch.randn(1, d, r, device=dev) * 0.02).to(torch.bfloat16)
    idx = torch.zeros(B, dtype=torch.int32, device=dev)
    u = ((x.float().view(M, d) @ A[0].float().t()) * 2.0).to(torch.bfloat16)
    Wt = W.t().contiguous()
    At = ops.pack_lora_b(A.transpose(1, 2).contiguous())
    Bt = Bm.transpose(1, 2).contiguous()
    dA = torch.zeros(1, r, d, device=dev)
    dB = torch.zeros(1, d, r, device=dev)
    for _ in range(3):
        ops.qv_lora_bwd(dy, x, u, Wt, At, Bt, idx, dA, dB, 2.0, need_dx=need_dx)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.qv_lora_bwd(dy, x, u, Wt, At, Bt, idx, dA, dB, 2.0, need_dx=need_dx)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    flops_dx = (2.0 * M * d * d + 2.0 * M * r * 2 * d) if need_dx else 2.0 * M * r * d
    bytes_skinny = 2.0 * M * d * 2 + 2.0 * M * r * 2          # one read of x and of dy (+ v and u)
    print(f"{name}: M={M} d={d} r={r}: whole call {us:.1f} us | dX algorithmic {flops_dx / 1e9:.1f} GFLOP, "
          f"dA/dB algorithmic {bytes_skinny / 1e6:.1f} MB")
