"""Debug tool: per-phase cycle stamps of the attention kernel (needs tools/ubench/libsar_trace.so, built with -DFA_TRACE).

  nvcc ... -DFA_TRACE -shared -o tools/ubench/libsar_trace.so speech_adapter_routing_b200/csrc/*.cu
  python tools/trace_attn.py
"""
import ctypes, os, sys
import numpy as np
import torch

here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "ubench", "libsar_trace.so"))
vp, i32 = ctypes.c_void_p, ctypes.c_int
lib.sar_attn_fwd.argtypes = [vp] * 4 + [i32] * 6 + [vp]
lib.sar_attn_fwd.restype = i32
lib.sar_last_error.restype = ctypes.c_char_p
lib.sar_debug_fa_trace.argtypes = [vp, i32]
lib.sar_debug_fa_trace.restype = i32

B, H, T = 64, 12, 1500
q = torch.randn(B * H, T, 64, device="cuda", dtype=torch.bfloat16) * 0.35
k = torch.randn(B * H, T, 64, device="cuda", dtype=torch.bfloat16)
v = torch.randn(B * H, T, 64, device="cuda", dtype=torch.bfloat16)
o = torch.empty_like(q)
st = torch.cuda.current_stream().cuda_stream


def run():
    rc = lib.sar_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, T, T, 64, 0, st)
    assert rc == 0, lib.sar_last_error()


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.sar_debug_fa_trace(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("traced launch: %.1f us" % (e0.elapsed_time(e1) * 1e3))
n = 16 * 24 * 12
buf = np.zeros(n, dtype=np.int64)
lib.sar_debug_fa_trace(buf.ctypes.data, 0)
tr = buf.reshape(16, 24, 12)
t0 = tr[tr > 0].min()
names = ["top", "max", "exp0", "ldS0", "sts0", "exp1", "sts1", "arrive", "mma:S_j", "mma:p_full", "mma:PV"]
for c in range(16):
    if tr[c].max() == 0:
        continue
    print("CTA slot", c)
    for j in range(24):
        row = tr[c, j]
        if row.max() == 0:
            continue
        base = row[0]
        print("  j=%2d top@%8d |" % (j, row[0] - t0), " ".join("%s+%d" % (names[e], row[e] - base) for e in range(1, 11) if row[e]))
