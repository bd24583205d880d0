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


def toy_backoff_lm(V=40, n_hist=12, seed=31, backoff_id=1):
    """A small back-off "bigram" LM as a sorted arc table, for the FST shallow-fusion fixture.  Labels follow the reference's
    convention ilabel = token + 1 (decoder/beam_transducer.py:139), so ilabel 1 (= blank + 1) is never queried and serves as the
    back-off label.  State 0 = unigram state (an arc for every token, final); states 1..n_hist = history states with arcs for a
    random third of the tokens plus a back-off arc to state 0; half of them are final.
    Returns (arcs, finals): arcs[state] = list of (ilabel, weight, nextstate) sorted by ilabel; finals[state] = cost or inf."""
    rng = np.random.default_rng(seed)
    arcs, finals = [], []
    uni = [(y + 1, float(np.round(rng.uniform(0.1, 1.5), 3)), int(rng.integers(0, n_hist + 1))) for y in range(1, V)]
    arcs.append(sorted(uni))
    finals.append(float(np.round(rng.uniform(0.5, 2.0), 3)))
    for s in range(1, n_hist + 1):
        toks = sorted(rng.choice(np.arange(1, V), size=V // 3, replace=False).tolist())
        a = [(backoff_id, float(np.round(rng.uniform(0.2, 1.5), 3)), 0)]
        a += [(int(y) + 1, float(np.round(rng.uniform(0.05, 1.0), 3)), int(rng.integers(0, n_hist + 1))) for y in toks]
        arcs.append(sorted(a))
        finals.append(float(np.round(rng.uniform(0.5, 2.0), 3)) if s % 2 == 0 else float("inf"))
    return arcs, finals
