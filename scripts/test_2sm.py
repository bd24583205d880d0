import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pika_b200 import kernels as K
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
def rnd(*s, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed); return torch.randn(*s, generator=g, device="cuda").to(torch.bfloat16)
for (M, N, Kd, a_mn, b_mn, cdt) in [(256, 256, 64, 0, 0, torch.float32), (300, 520, 200, 0, 0, torch.float32), (1000, 6000, 1024, 0, 0, torch.bfloat16),
                                    (512, 512, 256, 0, 1, torch.float32), (512, 512, 256, 1, 0, torch.float32), (776, 1032, 320, 1, 1, torch.float32),
                                    (128 * 37, 256 * 5, 192, 0, 0, torch.bfloat16)]:
    a = rnd(Kd, M, seed=1) if a_mn else rnd(M, Kd, seed=1)
    b = rnd(Kd, N, seed=2) if b_mn else rnd(N, Kd, seed=2)
    c = torch.full((M, N), float("nan"), device="cuda", dtype=cdt)
    K.gemm(a, b, c, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=256, two_sm=1, k_splits=1)
    torch.cuda.synchronize()
    ref = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
    print("2sm", (M, N, Kd, a_mn, b_mn, str(cdt)), "rel", rel(c, ref), flush=True)
# epilogue features + batched
M, N, Kd = 520, 768, 128
a, b = rnd(M, Kd, seed=3), rnd(N, Kd, seed=4)
bias = torch.randn(N, device="cuda"); res = rnd(M, N, seed=5)
c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
K.gemm(a, b, c, bias=bias, act=K.ACT_RELU, aux=res, aux_mode=K.AUX_ADD, two_sm=1)
print("2sm epilogue rel", rel(c, torch.relu(a.float() @ b.float().t() + bias) + res.float()), flush=True)
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
for name, (M, N, Kd, a_mn, b_mn, cdt) in {"sq8192": (8192, 8192, 8192, 0, 0, torch.bfloat16), "fc2_like": (144960, 6000, 1024, 0, 0, torch.bfloat16),
                                          "tdnn_like": (32000, 1024, 3072, 0, 0, torch.bfloat16), "dgrad_like": (32000, 3072, 1024, 0, 1, torch.bfloat16),
                                          "wgrad_like": (6000, 1024, 290000, 1, 1, torch.float32)}.items():
    a = rnd(Kd, M, seed=1) if a_mn else rnd(M, Kd, seed=1)
    b = rnd(Kd, N, seed=2) if b_mn else rnd(N, Kd, seed=2)
    c = torch.empty(M, N, device="cuda", dtype=cdt)
    for mode in (-1, 1):
        ms = timeit(lambda: K.gemm(a, b, c, a_mn=bool(a_mn), b_mn=bool(b_mn), two_sm=mode))
        print(name, "two_sm" if mode == 1 else "one_sm", "%.3f ms %.0f TFLOP/s" % (ms, 2.0 * M * N * Kd / ms / 1e9), flush=True)
