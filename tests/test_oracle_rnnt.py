"""Pins oracle/rnnt.py + oracle/rnnt_c.c: golden vectors from torchaudio's independent
implementation (tests/golden/rnnt_loss.npz), brute-force path enumeration, and numpy-vs-C."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import rnnt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "liboracle_rnnt.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return ctypes.CDLL(so)


def run_c(logits, labels, fl, ll):
    lib = c_oracle()
    B, T, U1, V = logits.shape
    logits = np.ascontiguousarray(logits, np.float32)
    labels = np.ascontiguousarray(labels, np.int32).reshape(B, -1)
    costs = np.zeros(B)
    dz = np.zeros_like(logits)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.oracle_rnnt_loss(p(logits), p(labels), p(np.ascontiguousarray(fl, np.int32)),
                              p(np.ascontiguousarray(ll, np.int32)), B, T, U1, V, V,
                              max(labels.shape[1], 0), p(costs), p(dz))
    assert rc == 0
    return costs, dz


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_numpy_oracle_matches_torchaudio_golden(golden_dir, i):
    d = np.load(os.path.join(golden_dir, "rnnt_loss.npz"))
    logits, labels, fl, ll = d["logits_%d" % i], d["labels_%d" % i], d["fl_%d" % i], d["ll_%d" % i]
    costs, grads = rnnt.rnnt_loss(rnnt.log_softmax(logits), labels, fl, ll)
    np.testing.assert_allclose(costs, d["costs_%d" % i], rtol=2e-6, atol=1e-5)
    np.testing.assert_allclose(grads, d["grads_%d" % i], rtol=1e-4, atol=1e-5)  # golden is fp32


def test_brute_force_small():
    rng = np.random.default_rng(3)
    for T, U in [(1, 0), (1, 2), (3, 0), (4, 3), (5, 4)]:
        V = 6
        lp = rnnt.log_softmax(rng.standard_normal((1, T, U + 1, V)))
        labels = rng.integers(1, V, (1, max(U, 1)))[:, :U]
        c, _ = rnnt.rnnt_loss(lp, labels, [T], [U])
        bf = rnnt.rnnt_brute_force(lp[0], labels[0], T, U)
        assert abs(c[0] - bf) < 1e-9


def test_gradient_is_probability_flow():
    """sum of blank+label gradient mass leaving node (0,0) is -1; grads vanish outside (T_n,U_n)."""
    rng = np.random.default_rng(4)
    lp = rnnt.log_softmax(rng.standard_normal((2, 6, 5, 9)))
    labels = rng.integers(1, 9, (2, 4))
    fl, ll = np.array([6, 4]), np.array([4, 2])
    _, g = rnnt.rnnt_loss(lp, labels, fl, ll)
    for n in range(2):
        assert abs(g[n, 0, 0].sum() + 1.0) < 1e-9
        assert np.all(g[n, fl[n]:] == 0) and np.all(g[n, :, ll[n] + 1:] == 0)


@pytest.mark.parametrize("shape", [(2, 9, 4, 17), (3, 20, 7, 50)])
def test_c_oracle_matches_numpy(shape):
    B, T, U, V = shape
    rng = np.random.default_rng(11)
    logits = (2 * rng.standard_normal((B, T, U + 1, V))).astype(np.float32)
    labels = rng.integers(1, V, (B, U)).astype(np.int32)
    fl = rng.integers(T // 2, T + 1, B).astype(np.int32)
    ll = rng.integers(U // 2, U + 1, B).astype(np.int32)
    c0, dz0 = rnnt.rnnt_loss_from_logits(logits, labels, fl, ll)
    c1, dz1 = run_c(logits, labels, fl, ll)
    np.testing.assert_allclose(c1, c0, rtol=1e-10)
    # padded region: numpy oracle yields exact zeros; the C oracle too
    np.testing.assert_allclose(dz1, dz0, atol=2e-7)


def test_fused_gradient_matches_finite_difference():
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((1, 4, 3, 5))
    labels = np.array([[2, 4]])
    c, dz = rnnt.rnnt_loss_from_logits(logits, labels, [4], [2])
    eps = 1e-6
    for idx in [(0, 0, 0, 0), (0, 1, 1, 4), (0, 3, 2, 0), (0, 2, 0, 2), (0, 2, 1, 3)]:
        zp = logits.copy(); zp[idx] += eps
        zm = logits.copy(); zm[idx] -= eps
        fd = (rnnt.rnnt_loss_from_logits(zp, labels, [4], [2])[0][0] -
              rnnt.rnnt_loss_from_logits(zm, labels, [4], [2])[0][0]) / (2 * eps)
        assert abs(fd - dz[idx]) < 1e-6
