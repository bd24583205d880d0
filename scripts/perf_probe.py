"""Quick device-side timing of the two graded kernels (run on the B200 via gpurun)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pika_b200 import kernels as K

def timeit(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(it)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(it)]
    for i in range(it):
        st[i].record(); fn(); en[i].record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in zip(st, en))
    return ts[len(ts)//2], ts[0]

out = {}
def gemm_case(name, M, N, Kd, bn=0, cdt=torch.bfloat16, a_mn=False, b_mn=False):
    a = torch.randn((Kd, M) if a_mn else (M, Kd), device="cuda").to(torch.bfloat16)
    b = torch.randn((Kd, N) if b_mn else (N, Kd), device="cuda").to(torch.bfloat16)
    c = torch.empty(M, N, device="cuda", dtype=cdt)
    med, best = timeit(lambda: K.gemm(a, b, c, a_mn=a_mn, b_mn=b_mn, block_n=bn))
    fl = 2.0 * M * N * Kd
    out[name] = dict(ms=med, tflops=fl / med / 1e9, best_tflops=fl / best / 1e9)
    print(name, out[name], flush=True)
    ref, _ = timeit(lambda: torch.matmul(a.t() if a_mn else a, b if b_mn else b.t()))
    print("   cublas", fl / ref / 1e9, flush=True)

which = sys.argv[1:] or ["gemm", "rnnt"]
if "gemm" in which:
    gemm_case("sq8192", 8192, 8192, 8192)
    gemm_case("sq8192_bn128", 8192, 8192, 8192, bn=128)
    gemm_case("tdnn_like", 32000, 1024, 3072)
    gemm_case("fc2_like", 144960, 6000, 1024)
    gemm_case("ffn1", 32000, 4096, 1024)
    gemm_case("wgrad_like", 1024, 3072, 32000, cdt=torch.float32, a_mn=True, b_mn=True)
    gemm_case("dgrad_like", 32000, 3072, 1024, b_mn=True)
if "rnnt" in which:
    for (B, T, U1, V) in [(8, 240, 151, 6000), (32, 240, 151, 6000)]:
        z = torch.randn(B, T, U1, V, device="cuda").to(torch.bfloat16)
        lab = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
        fl = torch.full((B,), T, device="cuda", dtype=torch.int32)
        ll = torch.full((B,), U1 - 1, device="cuda", dtype=torch.int32)
        dz = torch.empty_like(z)
        med, best = timeit(lambda: K.rnnt_loss_fwd_bwd(z, lab, fl, ll, dlogits=dz), warm=2, it=5)
        byts = 3.0 * z.numel() * 2
        out["rnnt_B%d" % B] = dict(ms=med, gbs=byts / med / 1e6, best_gbs=byts / best / 1e6)
        print("rnnt", B, out["rnnt_B%d" % B], flush=True)
        med2, _ = timeit(lambda: K.rnnt_loss_fwd_bwd(z, lab, fl, ll, want_grad=False), warm=2, it=5)
        print("   loss-only ms", med2, "GB/s", z.numel() * 2 / med2 / 1e6, flush=True)
        del z, dz
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/perf_probe.json", "w"), indent=1)
