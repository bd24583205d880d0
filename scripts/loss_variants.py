import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from pika_b200 import kernels as K
    B, T, U1, V = 32, 240, 151, 6000
    z = (torch.randn(B, T, U1, V, device="cuda") * 2).to(torch.bfloat16)
    lab = torch.randint(1, V, (B, U1 - 1), device="cuda", dtype=torch.int32)
    fl = torch.full((B,), T, device="cuda", dtype=torch.int32); ll = torch.full((B,), U1 - 1, device="cuda", dtype=torch.int32)
    cs = torch.empty(V, device="cuda")
    for use_cs in (False, True):
        for _ in range(2): K.rnnt_loss_fwd_bwd(z, lab, fl, ll, dlogits=z, colsum=cs if use_cs else None)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): K.rnnt_loss_fwd_bwd(z, lab, fl, ll, dlogits=z, colsum=cs if use_cs else None)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        print("variant", os.environ.get("PK_RNNT_GRAD_VARIANT"), "colsum", use_cs, "ms %.3f GB/s %.0f" % (ms, 3 * z.numel() * 2 / ms / 1e6), flush=True)
else:
    for v in "0123":
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, PK_RNNT_GRAD_VARIANT=v))
