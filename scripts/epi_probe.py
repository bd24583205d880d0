"""epilogue probe: fc2-forward-like GEMM (bf16 out, bias) with / without the fused row-LSE partials, plus short-K f32/bf16 shapes"""
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
M, N, Kd = 1159680 // 2, 6000, 1024
a, b = rnd(M, Kd, seed=1, scale=0.3), rnd(N, Kd, seed=2, scale=0.05)
bias = torch.randn(N, device="cuda")
c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
parts = torch.empty((N + 255) // 256, M, 2, device="cuda")
for lse in (None, parts):
    ms = timeit(lambda: K.gemm(a, b, c, bias=bias, block_n=256, row_lse=lse))
    print("fc2-like half-M %s: %.3f ms %.0f TF/s" % ("+row_lse" if lse is not None else "plain  ", ms, 2.0 * M * N * Kd / ms / 1e9), flush=True)
del a, b, c, parts
for (M, N, Kd, cdt) in [(31808, 4096, 1024, torch.bfloat16), (31808, 1024, 1024, torch.bfloat16), (15904, 994 * 4, 64, torch.float32), (31808, 4096, 128, torch.bfloat16)]:
    a, b = rnd(M, Kd, seed=1), rnd(N, Kd, seed=2)
    c = torch.empty(M, N, device="cuda", dtype=cdt)
    ms = timeit(lambda: K.gemm(a, b, c))
    print((M, N, Kd, str(cdt)), "%.3f ms %.0f TF/s  %.0f GB/s out" % (ms, 2.0 * M * N * Kd / ms / 1e9, c.numel() * c.element_size() / ms / 1e6), flush=True)
