"""One GEMM shape, a few launches (for ncu captures).  usage: one_gemm.py M N K [f32|bf16] [a_mn] [b_mn] [lse]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pika_b200 import kernels as K
a = sys.argv[1:]
M, N, Kd = [int(v) for v in (a[0:3] if len(a) > 2 else (32000, 4096, 1024))]
cdt = torch.float32 if (len(a) > 3 and a[3] == "f32") else torch.bfloat16
a_mn = len(a) > 4 and a[4] == "1"
b_mn = len(a) > 5 and a[5] == "1"
A = torch.randn((Kd, M) if a_mn else (M, Kd), device="cuda").to(torch.bfloat16)
B = torch.randn((Kd, N) if b_mn else (N, Kd), device="cuda").to(torch.bfloat16)
C = torch.empty(M, N, device="cuda", dtype=cdt)
lse = len(a) > 6 and a[6] == "1"
parts = torch.empty((N + 255) // 256, M, 2, device="cuda") if lse else None
bias = torch.zeros(N, device="cuda") if lse else None
for _ in range(4):
    K.gemm(A, B, C, a_mn=a_mn, b_mn=b_mn, k_splits=1, bias=bias, row_lse=parts, block_n=256 if lse else 0)
torch.cuda.synchronize()
