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


def decode_big_inputs(seed=606, B=6, T=330):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, T, 240)).astype(np.float32)
