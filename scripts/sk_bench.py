"""stream-K on/off for the wgrad shapes of the bench step (in-step-like data: A = sparse-ish gradient, B = activations)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pika_b200 import kernels as K
def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed); return (torch.randn(*s, generator=g, device="cuda") * scale).to(torch.bfloat16)
def timeit(fn, it=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
for (M, N, Kd) in [(6000, 1024, 1159680), (1024, 1024, 31808), (4096, 1024, 31808), (1024, 4096, 31808), (1024, 1024, 7680), (4096, 1024, 4832)]:
    a = rnd(Kd, M, seed=1, scale=0.01); b = rnd(Kd, N, seed=2)
    c = torch.empty(M, N, device="cuda")
    out = []
    for ks in (1, 0):
        ms = timeit(lambda: K.gemm(a, b, c, a_mn=True, b_mn=True, k_splits=ks))
        out.append("%s %.3f ms %.0f TF/s" % ("off" if ks == 1 else "streamK", ms, 2.0 * M * N * Kd / ms / 1e9))
    print((M, N, Kd), " | ".join(out), flush=True)
    del a, b, c
