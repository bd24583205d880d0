"""fused vs materialised attention at the bench shape (B=32, heads=16, T=994): fwd+bwd ms"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pika_b200 import engine as E
E.set_precision("bf16")
B, T, heads = int(os.environ.get("B", 32)), int(os.environ.get("T", 994)), 16
D = heads * 64
g = torch.Generator(device="cuda").manual_seed(1)
qkv0 = (torch.randn(B, T, 3 * D, generator=g, device="cuda") * 0.5).bfloat16()
dy = torch.randn(B, T, D, generator=g, device="cuda").bfloat16()
def run(fused, p):
    E._FUSED_ATTN = fused
    qkv = qkv0.clone().requires_grad_(True)
    out = E.AttentionFn.apply(qkv, heads, p, 99)
    out.backward(dy)
    return out, qkv.grad
for p in (0.0, 0.1):
    for fused in (False, True):
        for _ in range(2): run(fused, p)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): run(fused, p)
        e.record(); torch.cuda.synchronize()
        print("drop %.1f fused=%d  %.3f ms fwd+bwd" % (p, fused, s.elapsed_time(e) / 5), flush=True)
    a, b = run(True, p), run(False, p)
    r = lambda x, y: ((x.float() - y.float()).norm() / y.float().norm()).item()
    print("  rel out %.2e  rel grad %.2e" % (r(a[0], b[0]), r(a[1], b[1])), flush=True)
