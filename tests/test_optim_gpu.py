"""CUDA optimiser / BMUF kernels (pika_b200/csrc/optim.cu) against the oracle (oracle/train.py), through the C ABI.

Reference: trainer/train_transducer_bmuf_otfaug.py:105-110 (clip_grad_norm_(inf) + SGD(nesterov).step()),
trainer/bmuf.py:76-100 (BmufTrainer.update_and_sync).  fp32 element-wise arithmetic: the kernels evaluate the same
expressions in the same order as torch's, so parameters agree to 1 ulp-class tolerances (stated per assert).
The last test runs the NCCL BmufTrainer on 2 GPUs against the parameter trajectory of the reference's own
BmufTrainer (tests/golden/bmuf_2rank.npz); it is skipped on a single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rand(n, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(n) * scale).astype(np.float32)


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096 + 5, 1 << 20, (1 << 20) + 3])
def test_absmax_matches_numpy(n):
    from pika_b200 import kernels as K
    x = _rand(n, n, 3.0)
    x[n // 2] = -17.5 if n > 2 else x[n // 2]
    out = torch.full((1,), 123.0, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    K.absmax(torch.from_numpy(x).cuda(), out, flag)
    assert out.item() == float(np.abs(x).max())            # exact: a max of fp32 values
    assert flag.item() == 0


@pytest.mark.parametrize("n,pos", [(5, 4), (4096, 17), ((1 << 18) + 3, (1 << 18) + 2), (1 << 18, 0)])
def test_absmax_nan_flag(n, pos):
    """a NaN anywhere (vector body or the n % 4 tail) raises the flag; torch's clip_grad_norm_(inf) would see a NaN norm"""
    from pika_b200 import kernels as K
    x = _rand(n, 7)
    x[pos] = np.nan
    out = torch.zeros(1, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    K.absmax(torch.from_numpy(x).cuda(), out, flag)
    assert flag.item() == 1


@pytest.mark.parametrize("n", [7, 4096, (1 << 20) + 1])
@pytest.mark.parametrize("clip", [-1.0, 3.0, 0.25])
def test_sgd_nesterov_clip_matches_oracle_and_torch(n, clip):
    """three consecutive steps (first=True, then momentum), clip off / inactive / active, n not a multiple of 4"""
    from oracle import train as ot
    from pika_b200 import kernels as K
    lr, mom = 4e-4, 0.9
    p0 = _rand(n, 1)
    p_dev = torch.from_numpy(p0.copy()).cuda()
    buf_dev = torch.zeros(n, device="cuda")
    absmax_t = torch.zeros(1, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    p_ref, buf_ref = p0.astype(np.float32), np.zeros(n, np.float32)
    # independent check: torch's own SGD + clip_grad_norm_(inf) on the CPU
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    topt = torch.optim.SGD([tp], lr, momentum=mom, nesterov=True)
    for it in range(3):
        g = _rand(n, 10 + it, 0.5 if it != 1 else 2.0)      # step 1 has the larger gradient (clip 3.0 becomes active there)
        g_dev = torch.from_numpy(g).cuda()
        if clip > 0:
            flag.zero_()
            K.absmax(g_dev, absmax_t, flag)
        K.sgd_nesterov_clip(p_dev, g_dev, buf_dev, lr, mom, clip, absmax_t if clip > 0 else None, it == 0,
                            nan_flag=flag if clip > 0 else None)
        coef = np.float32(1.0)
        if clip > 0:
            tot, c = ot.clip_coef_inf([g], clip)
            assert absmax_t.item() == np.float32(tot)
            coef = np.float32(min(np.float32(clip) / (np.float32(tot) + np.float32(1e-6)), np.float32(1.0)))
        gc = (g * coef).astype(np.float32)
        p_ref, buf_ref = ot.sgd_nesterov_step(p_ref, gc, buf_ref, np.float32(lr), np.float32(mom), it == 0)
        p_ref, buf_ref = p_ref.astype(np.float32), buf_ref.astype(np.float32)
        tp.grad = torch.from_numpy(g.copy())
        if clip > 0:
            torch.nn.utils.clip_grad_norm_([tp], clip, norm_type=float("inf"))
        topt.step()
        # same fp32 expressions; fused multiply-adds on the device differ from numpy's separate roundings by <= 1 ulp of |p|
        np.testing.assert_allclose(p_dev.cpu().numpy(), p_ref, rtol=0, atol=2e-7 * max(1.0, float(np.abs(p_ref).max())))
        np.testing.assert_allclose(buf_dev.cpu().numpy(), buf_ref, rtol=1e-6, atol=3e-7)     # fused multiply-add vs two roundings near cancellation
        np.testing.assert_allclose(p_dev.cpu().numpy(), tp.detach().numpy(), rtol=0, atol=2e-7 * max(1.0, float(np.abs(p_ref).max())))


def test_sgd_clip_propagates_nan_like_torch():
    """clip_grad_norm_(inf) of a gradient holding one NaN: total = NaN, coef = NaN, every parameter turns NaN (reference behaviour)"""
    from pika_b200 import kernels as K
    n = 1000
    g = _rand(n, 3)
    g[123] = np.nan
    p = torch.from_numpy(_rand(n, 4)).cuda()
    buf = torch.zeros(n, device="cuda")
    am = torch.zeros(1, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    gd = torch.from_numpy(g).cuda()
    K.absmax(gd, am, flag)
    K.sgd_nesterov_clip(p, gd, buf, 1e-3, 0.9, 3.0, am, True, nan_flag=flag)
    tp = torch.nn.Parameter(torch.from_numpy(_rand(n, 4)))
    tp.grad = torch.from_numpy(g.copy())
    torch.nn.utils.clip_grad_norm_([tp], 3.0, norm_type=float("inf"))
    torch.optim.SGD([tp], 1e-3, momentum=0.9, nesterov=True).step()
    assert bool(torch.isnan(tp).all()) and bool(torch.isnan(p).all())


@pytest.mark.parametrize("n", [5, 4099, (1 << 20) + 2])
@pytest.mark.parametrize("world", [1, 2, 8])
def test_bmuf_delta_update_match_oracle(n, world):
    """two consecutive block syncs: delta = global - local; the sum over ranks is formed on the host here (NCCL in production)"""
    from oracle import train as ot
    from pika_b200 import kernels as K
    bm, blr = 0.9, 1.0
    glob = _rand(n, 1)
    g_dev = torch.from_numpy(glob.copy()).cuda()
    dprev_dev = torch.zeros(n, device="cuda")
    g_ref, dprev_ref = glob.astype(np.float64), np.zeros(n, np.float64)
    for it in range(2):
        locals_ = [(g_ref + 0.01 * _rand(n, 100 * it + r)).astype(np.float32) for r in range(world)]
        dsum = torch.zeros(n, device="cuda")
        last_local = None
        for r in range(world):
            l_dev = torch.from_numpy(locals_[r]).cuda()
            d = torch.empty(n, device="cuda")
            K.bmuf_delta(g_dev, l_dev, d)
            np.testing.assert_array_equal(d.cpu().numpy(), g_dev.cpu().numpy() - locals_[r])   # exact fp32 subtraction
            dsum += d                                         # stands in for the all-reduce
            last_local = l_dev
        K.bmuf_update(g_dev, last_local, dprev_dev, dsum, world, bm, blr)
        g_ref, dprev_ref = ot.bmuf_update(g_ref, [l.astype(np.float64) for l in locals_], dprev_ref, bm, blr)
        np.testing.assert_allclose(g_dev.cpu().numpy(), g_ref, rtol=0, atol=1e-6)
        np.testing.assert_allclose(dprev_dev.cpu().numpy(), dprev_ref, rtol=0, atol=1e-6)
        assert torch.equal(last_local, g_dev)                 # local := global (the broadcast of trainer/bmuf.py:98)


# ------------------------------------------------------------------------------------------------ 2 GPUs, NCCL
def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from pika_b200.trainer.bmuf import BmufTrainer, SUCCESS, STOP
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    torch.manual_seed(100 + rank)                      # same construction as make_golden.py:golden_bmuf
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).to(dev)
    tr = BmufTrainer(0, rank, world, model, 0.9, 1.0)  # CUDA kernels + NCCL all-reduce
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bmuf_2rank.npz"))["params"]
    pvec = lambda: torch.nn.utils.parameters_to_vector(model.parameters()).detach().cpu().numpy()   # noqa: E731
    ok = np.array_equal(pvec(), gold[0])
    errs = []
    for it in range(3):
        g = torch.Generator().manual_seed(7 * it + rank)
        noise = (0.01 * torch.randn(gold.shape[1], generator=g)).to(dev)
        with torch.no_grad():
            off = 0
            for prm in model.parameters():
                prm.add_(noise[off:off + prm.numel()].view_as(prm))
                off += prm.numel()
        assert tr.update_and_sync() == SUCCESS
        errs.append(float(np.abs(pvec() - gold[it + 1]).max()))
        ok &= errs[-1] <= 2e-7 and torch.equal(tr.param, tr.flat.data)
    # every rank holds bit-identical parameters after a sync
    mine = tr.flat.data.clone()
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    same = all(torch.equal(g_, gathered[0]) for g_ in gathered)
    if rank == 1:
        with torch.no_grad():
            tr.flat.data[3] = float("nan")
    stop = tr.update_and_sync()
    q.put((rank, bool(ok), same, stop == STOP, errs))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_bmuf_nccl_two_gpus_matches_reference_trajectory():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(30)
    for r in res:
        assert all(r[1:4]), r
