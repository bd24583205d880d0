"""Per-shape GEMM census of one bench train step (each pk_gemm_bf16 call timed with CUDA events), then
standalone sweeps: k_splits per wgrad shape, and operand-major combinations.  Exploration tool, not a bench."""
import sys, os, types, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pika_b200 import engine, kernels as K
from pika_b200.frontend import FbankOptions, Frontend
from pika_b200.model.transducer import Net
from pika_b200.trainer.bmuf import BmufTrainer
from pika_b200.trainer.flat import FlatParams, SgdNesterovClip
from pika_b200.trainer.step import TrainStep
from pika_b200.utils.spec_augment import SpecAugment

dev = torch.device("cuda", 0)
if os.environ.get("CENSUS", "1") == "1":
    a = types.SimpleNamespace(batch=32, T=1000, U=150, V=6000)
    ta = bench.train_args()
    torch.manual_seed(777)
    model = Net(bench.model_args(a.V), 240, a.V).to(dev); model.train()
    flat = FlatParams(model); bmuf = BmufTrainer(0, 0, 1, model, 0.9, 1.0, flat=flat)
    opt = SgdNesterovClip(flat, ta.initial_lr, 0.9, 3.0)
    fe = Frontend(FbankOptions(num_mel_bins=80, low_freq=40.0, high_freq=-200.0, dither=0.0, window_type="hamming"), 1, 1, dev)
    step = TrainStep(model, ta, fe, bmuf, opt, spec_augmentor=SpecAugment(15, 35))
    B = a.batch
    pcm = torch.from_numpy(bench.synth_pcm(B, a.T, 777)).to(dev)
    n = pcm.shape[1]
    new_len, frames = Frontend.lengths([n] * B, [1.0] * B)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    batch = dict(pcm=pcm, target=torch.randint(1, a.V, (B, a.U), device=dev), n_samples=i32([n] * B), new_len=i32(new_len),
                 n_frames=i32(frames), ali_lens=i32([a.U] * B), rate=torch.ones(B, device=dev), target_db=torch.full((B,), -25.0, device=dev),
                 t_max=max(frames))
    for _ in range(3):
        step(batch)
    torch.cuda.synchronize()
    rec = []
    orig = K.gemm

    def timed(a_, b_, c_, *args, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(a_, b_, c_, *args, **kw)
        e.record()
        al = a_ if isinstance(a_, (list, tuple)) else [a_]
        a_mn, b_mn = bool(kw.get("a_mn", False)), bool(kw.get("b_mn", False))
        a0 = al[0]
        Kd = a0.shape[-2] if a_mn else a0.shape[-1]
        M, N = c_.shape[-2], c_.shape[-1]
        zb = int(np.prod(c_.shape[:-2])) if c_.dim() > 2 else 1
        rec.append(((M, N, Kd, len(al), kw.get("kz_count", 1), zb, int(a_mn), int(b_mn), str(c_.dtype).replace("torch.", ""),
                     int(kw.get("k_splits", 0))), s, e))
        return r
    K.gemm = timed
    step(batch)
    torch.cuda.synchronize()
    K.gemm = orig
    agg = collections.defaultdict(lambda: [0, 0.0])
    for sig, s, e in rec:
        agg[sig][0] += 1
        agg[sig][1] += s.elapsed_time(e)
    tot = sum(v[1] for v in agg.values())
    print("GEMM census: %d calls, %.2f ms total (event-timed, includes launch gaps)" % (len(rec), tot))
    print("   ms    n   TF/s   (M, N, K, pairs, kz, zb, a_mn, b_mn, cdt, ksplit)")
    for sig, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        M, N, Kd, npair, kz, zb = sig[:6]
        fl = 2.0 * M * N * Kd * npair * max(kz, 1) * zb * c
        print("%7.3f %4d %6.0f   %s" % (t, c, fl / t / 1e9, sig), flush=True)
    del step, model, flat, bmuf, opt, batch, pcm
    torch.cuda.empty_cache()


def rnd(*s, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*s, generator=g, device="cuda").to(torch.bfloat16)


def timeit(fn, it=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it


print("--- operand-major matrix at 8192^3 (bf16 out)")
for a_mn in (0, 1):
    for b_mn in (0, 1):
        M = N = Kd = 8192
        a_ = rnd(Kd, M, seed=1) if a_mn else rnd(M, Kd, seed=1)
        b_ = rnd(Kd, N, seed=2) if b_mn else rnd(N, Kd, seed=2)
        c_ = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: K.gemm(a_, b_, c_, a_mn=bool(a_mn), b_mn=bool(b_mn)))
        print("a_mn=%d b_mn=%d  %.3f ms %.0f TFLOP/s" % (a_mn, b_mn, ms, 2.0 * M * N * Kd / ms / 1e9), flush=True)
print("--- dgrad: B MN-major vs pre-transposed K-major")
for (M, N, Kd) in [(32000, 3072, 1024), (289920, 1024, 6000), (32000, 1024, 4096), (32000, 1024, 1024)]:
    a_ = rnd(M, Kd, seed=1)
    bm = rnd(Kd, N, seed=2)
    bt = bm.t().contiguous()
    c_ = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for name, fn in (("b_mn", lambda: K.gemm(a_, bm, c_, b_mn=True)), ("b_kmajor", lambda: K.gemm(a_, bt, c_))):
        ms = timeit(fn)
        print((M, N, Kd), name, "%.3f ms %.0f TFLOP/s" % (ms, 2.0 * M * N * Kd / ms / 1e9), flush=True)
print("--- wgrad k_splits sweep (a_mn=b_mn=1, f32 out)")
for (M, N, Kd) in [(6000, 1024, 289920), (1024, 1024, 31808), (4096, 1024, 31808), (1024, 4096, 31808), (1024, 1024, 7680), (4096, 1024, 4832)]:
    a_ = rnd(Kd, M, seed=1)
    b_ = rnd(Kd, N, seed=2)
    c_ = torch.empty(M, N, device="cuda", dtype=torch.float32)
    out = []
    for ks in (0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 24, 32):
        if ks > 1 and Kd // 64 // ks < 8: continue
        ms = timeit(lambda: K.gemm(a_, b_, c_, a_mn=True, b_mn=True, k_splits=ks))
        out.append("%d:%.0f" % (ks, 2.0 * M * N * Kd / ms / 1e9))
    print((M, N, Kd), "TFLOP/s by k_splits (0=heuristic):", " ".join(out), flush=True)
