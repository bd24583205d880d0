"""One GEMM shape, a few launches (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pika_b200 import kernels as K
M, N, Kd = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (32000, 4096, 1024))]
cdt = torch.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else torch.bfloat16
a = torch.randn(M, Kd, device="cuda").to(torch.bfloat16)
b = torch.randn(N, Kd, device="cuda").to(torch.bfloat16)
c = torch.empty(M, N, device="cuda", dtype=cdt)
for _ in range(4):
    K.gemm(a, b, c)
torch.cuda.synchronize()
