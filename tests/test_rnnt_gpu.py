"""GPU parity of the fused log-softmax + RNN-T loss + gradient kernels against the oracle
(oracle/rnnt.py, oracle/rnnt_c.c) and the committed torchaudio goldens, through the C ABI.
Tolerance: costs 1e-4 relative (fp32 log-space DP vs float64 oracle), gradients 2e-5 absolute
(+5e-4 relative: fp32 log-space DP, well inside the 1e-3 the north star states)
for f32 logits; bf16 logits are compared after rounding the oracle's result to bf16."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_gpu(logits_np, labels, fl, ll, dtype, ldv=None, inplace=False, grad_scale=None):
    from pika_b200 import kernels
    B, T, U1, V = logits_np.shape
    ldv = ldv or V
    z = torch.zeros(B, T, U1, ldv, dtype=dtype, device="cuda")
    z[..., :V] = torch.from_numpy(logits_np).to("cuda").to(dtype)
    lab = torch.from_numpy(np.ascontiguousarray(labels, np.int32)).reshape(B, -1).cuda()
    if lab.shape[1] == 0:
        lab = torch.zeros(B, 1, dtype=torch.int32, device="cuda")
    gs = None if grad_scale is None else torch.tensor(grad_scale, dtype=torch.float32, device="cuda")
    costs, dz = kernels.rnnt_loss_fwd_bwd(z, lab, torch.from_numpy(np.asarray(fl, np.int32)).cuda(),
                                          torch.from_numpy(np.asarray(ll, np.int32)).cuda(), V=V,
                                          grad_scale=gs, dlogits=z if inplace else None)
    torch.cuda.synchronize()
    return costs.cpu().numpy(), dz.float().cpu().numpy(), z.float().cpu().numpy()


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_golden_f32(golden_dir, i):
    from oracle import rnnt
    d = np.load(os.path.join(golden_dir, "rnnt_loss.npz"))
    logits, labels, fl, ll = d["logits_%d" % i], d["labels_%d" % i], d["fl_%d" % i], d["ll_%d" % i]
    costs, dz, _ = run_gpu(logits, labels, fl, ll, torch.float32, ldv=((logits.shape[-1] + 3) // 4) * 4)
    np.testing.assert_allclose(costs, d["costs_%d" % i], rtol=1e-4, atol=1e-4)
    _, dz_ref = rnnt.rnnt_loss_from_logits(logits, labels, fl, ll)
    V = logits.shape[-1]
    np.testing.assert_allclose(dz[..., :V], dz_ref, atol=2e-5, rtol=5e-4)
    assert np.all(dz[..., V:] == 0)


@pytest.mark.parametrize("shape", [(3, 33, 12, 200), (2, 60, 40, 1000), (5, 17, 1, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ragged_vs_oracle(shape, dtype):
    from oracle import rnnt
    B, T, U, V = shape
    rng = np.random.default_rng(B * 1000 + T)
    logits = (3 * rng.standard_normal((B, T, U + 1, V))).astype(np.float32)
    labels = rng.integers(1, V, (B, U)).astype(np.int32)
    fl = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
    ll = rng.integers(0, U + 1, B).astype(np.int32)
    fl[0], ll[0] = T, U
    if dtype == torch.bfloat16:
        logits = torch.from_numpy(logits).to(torch.bfloat16).float().numpy()    # same rounded inputs for both
    gs = rng.uniform(0.5, 2.0, B).astype(np.float32)
    costs, dz, _ = run_gpu(logits, labels, fl, ll, dtype, grad_scale=gs)
    c_ref, dz_ref = rnnt.rnnt_loss_from_logits(logits, labels, fl, ll)
    dz_ref = dz_ref * gs[:, None, None, None]
    np.testing.assert_allclose(costs, c_ref, rtol=1e-4, atol=1e-4)
    if dtype == torch.float32:
        np.testing.assert_allclose(dz, dz_ref, atol=3e-5, rtol=1e-3)   # fp32 log-space DP over T+U = 180 steps
    else:
        ref_b = torch.from_numpy(dz_ref).to(torch.bfloat16).float().numpy()
        np.testing.assert_allclose(dz, ref_b, atol=2e-5, rtol=1.6e-2)   # <= 2 bf16 ulps
    # padded nodes get exact zeros
    for n in range(B):
        assert np.all(dz[n, fl[n]:] == 0) and np.all(dz[n, :, ll[n] + 1:] == 0)


def test_inplace_aliasing_matches_out_of_place():
    rng = np.random.default_rng(9)
    logits = rng.standard_normal((2, 20, 9, 128)).astype(np.float32)
    labels = rng.integers(1, 128, (2, 8)).astype(np.int32)
    c1, dz1, _ = run_gpu(logits, labels, [20, 15], [8, 5], torch.bfloat16)
    c2, dz2, zbuf = run_gpu(logits, labels, [20, 15], [8, 5], torch.bfloat16, inplace=True)
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_array_equal(dz1, dz2)
    np.testing.assert_array_equal(zbuf, dz2)


def test_gradient_sums_to_zero_per_node_and_flow_conservation_large():
    """Size-independent properties at a larger shape: d/dlogits sums to ~0 over V at every node
    (softmax Jacobian), and sum_n costs equals the C oracle's."""
    import ctypes, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "_build", "liboracle_rnnt.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(root, "oracle")])
    lib = ctypes.CDLL(so)
    B, T, U, V = 2, 120, 60, 2048
    rng = np.random.default_rng(1)
    logits = (2 * rng.standard_normal((B, T, U + 1, V))).astype(np.float32)
    labels = rng.integers(1, V, (B, U)).astype(np.int32)
    fl = np.array([T, T - 13], np.int32); ll = np.array([U, U - 7], np.int32)
    costs, dz, _ = run_gpu(logits, labels, fl, ll, torch.float32)
    assert np.abs(dz.sum(-1)).max() < 1e-4
    c_ref = np.zeros(B); dz_ref = np.zeros_like(logits)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.oracle_rnnt_loss(p(logits), p(labels), p(fl), p(ll), B, T, U + 1, V, V, U, p(c_ref), p(dz_ref)) == 0
    np.testing.assert_allclose(costs, c_ref, rtol=1e-4)
    np.testing.assert_allclose(dz, dz_ref, atol=3e-5, rtol=1e-3)   # fp32 log-space DP over T+U = 180 steps
