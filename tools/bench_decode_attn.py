"""Times the one-token decode attention kernels against torch SDPA (B = 64, whisper-small heads)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from speech_adapter_routing_b200 import ops


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for B, H, Tk in [(64, 12, 1500), (64, 20, 1500), (16, 12, 1500)]:
    dev = "cuda"
    # 12 layers' worth of K / V so that successive calls do not hit L2
    Ks = [torch.randn(B, H, Tk, 64, device=dev, dtype=torch.bfloat16) for _ in range(6)]
    Vs = [torch.randn(B, H, Tk, 64, device=dev, dtype=torch.bfloat16) for _ in range(6)]
    q = (torch.randn(B, H, 1, 64, device=dev) * 0.4).to(torch.bfloat16)
    i = [0]

    def own():
        i[0] = (i[0] + 1) % 6
        return ops.decode_cross_attn(q, Ks[i[0]], Vs[i[0]])

    def lib():
        i[0] = (i[0] + 1) % 6
        return F.scaled_dot_product_attention(q, Ks[i[0]], Vs[i[0]], scale=1.0)

    t_own, t_lib = timeit(own), timeit(lib)
    gb = 2 * B * H * Tk * 64 * 2 / 1e9
    print(f"cross B={B} H={H} Tk={Tk}: libsar {t_own:.1f} us ({gb / t_own * 1e6:.0f} GB/s)   torch SDPA {t_lib:.1f} us "
          f"({gb / t_lib * 1e6:.0f} GB/s)")
