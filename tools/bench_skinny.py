import sys; sys.path.insert(0, "/root/repo")
import torch
from speech_adapter_routing_b200 import ops
dev="cuda:0"
def timeit(fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters//20): g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e)/iters*1e3
for (M,K,N,act,res) in [(64,768,3072,1,False),(64,3072,768,0,True),(64,768,768,0,True),(64,768,2304,0,False),(64,1280,5120,1,False),(64,5120,1280,0,True)]:
    x=torch.randn(1,M,K,device=dev,dtype=torch.bfloat16); W=(torch.randn(N,K,device=dev)*0.02).to(torch.bfloat16); b=torch.zeros(N,device=dev,dtype=torch.bfloat16)
    r=torch.randn(1,M,N,device=dev,dtype=torch.bfloat16) if res else None
    y=torch.empty(1,M,N,device=dev,dtype=torch.bfloat16)
    t1=timeit(lambda: ops.linear_fwd(x,W,b,r,act,out=y))
    t2=timeit(lambda: ops.linear_fwd(x,W,b,r,act,out=y,block_n=128))
    t3=timeit(lambda: torch.nn.functional.linear(x,W,b))
    print(f"M={M} {K}->{N} act={act} res={res}: skinny {t1:.1f} us  pair(bn128) {t2:.1f} us  cuBLAS {t3:.1f} us  weights {N*K*2/1e6:.1f} MB -> {N*K*2/t1/1e6:.0f} GB/s")
