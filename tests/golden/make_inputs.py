"""Seeded inputs shared by make_golden.py (build container, reference side) and the GPU tests: numpy Generator streams
are platform independent, so large feature tensors are regenerated instead of committed."""
import numpy as np


def full_shape_inputs(seed=2025, B=2, T=1000, U=150, V=6000):
    """BASELINE (T=1000, U=150, V=6000) shape, ragged lengths"""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, 240)).astype(np.float32)
    y = rng.integers(1, V, (B, U)).astype(np.int64)
    lens = np.array([T, T - 131][:B], np.int32)
    ulens = np.array([U, U - 30][:B], np.int32)
    return x, y, lens, ulens


def decode_big_inputs(seed=606, B=6, Tp=72, H=1024):
    """seeded ENCODER OUTPUTS [B, T', H] of the beam-16 / V=6000 decode fixture (see make_golden.py:golden_decode_big)"""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, Tp, H)).astype(np.float32)


# decode_fixture_reinit arguments of the beam-16 / V=6000 fixture (shared by the generator and the GPU test)
DECODE_BIG_REINIT = dict(blank_bias=34.0, s_o=0.5, pred_scale=1.0, enc_scale=1.0)
