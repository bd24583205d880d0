"""Pins oracle/frontend.py: AudioSegment hot methods + splice against the reference's own code,
Kaldi fbank against torchaudio.compliance.kaldi (tests/golden/frontend.npz)."""
import os

import numpy as np
import pytest

from oracle import frontend as fe


@pytest.mark.parametrize("key,rate,db", [("r09", 0.9, -23.5), ("r10", 1.0, -41.0), ("r11", 1.1, -12.25)])
def test_augment_and_fbank(golden_dir, key, rate, db):
    d = np.load(os.path.join(golden_dir, "frontend.npz"))
    aug = fe.augment(d["pcm"], rate, db)
    assert aug.dtype == np.int16
    np.testing.assert_array_equal(aug, d["aug_" + key])          # integer path: bit-exact
    fb = fe.kaldi_fbank(aug.astype(np.float32))
    assert fb.shape == d["fbank_" + key].shape
    np.testing.assert_allclose(fb, d["fbank_" + key], rtol=0, atol=2e-3)
    assert np.abs(fb - d["fbank_" + key]).mean() < 1e-4
    np.testing.assert_array_equal(fe.splice(d["fbank_" + key], 1, 1)[::5], d["splice_" + key])


def test_batch_assembly_padding_and_filter():
    rng = np.random.default_rng(0)
    feats = [rng.standard_normal((t, 80)).astype(np.float32) for t in (30, 50, 41)]
    labels = [[3, 4, 5], [7] * 10, [1, 2]]
    data, tgt, lens, ali = fe.assemble_batch(feats, labels, tu_limit=15000, padding_tgt=99)
    assert data.shape == (3, 50, 240) and tgt.shape == (3, 10)
    assert lens.tolist() == [30, 50, 41] and ali.tolist() == [3, 10, 2]
    np.testing.assert_array_equal(data[0, 30:], np.broadcast_to(data[0, 29], (20, 240)))   # last-frame pad
    assert tgt[0, 3:].tolist() == [99] * 7
    # TU filter: U*T//3 > limit drops the utterance (loader/otf_utt_loader.py:247)
    data2, _, lens2, _ = fe.assemble_batch(feats, labels, tu_limit=100, padding_tgt=99)
    assert lens2.tolist() == [30, 41]
    none = fe.assemble_batch(feats, labels, tu_limit=0)
    assert none[0] is None and none[2].tolist() == [0]


def test_cmvn_and_specaug():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 20, 240)).astype(np.float32)
    stats = np.zeros((2, 81))
    mean, var, n = rng.standard_normal(80), np.abs(rng.standard_normal(80)) + 0.5, 100.0
    stats[0, :80], stats[0, 80] = mean * n, n
    stats[1, :80] = (var + mean * mean) * n
    off, sc = fe.cmvn_from_stats(stats)
    y = fe.apply_cmvn(x, off, sc, cmn=True)
    ref = (x - x.mean(1, keepdims=True) + off[None, None].astype(np.float32)) * sc[None, None].astype(np.float32)
    np.testing.assert_allclose(y, ref, atol=1e-5)
    z = fe.spec_augment(y, 10, 5, 3, 4)
    assert np.all(z[:, :, 10:15] == 0) and np.all(z[:, 3:7, :] == 0)
    assert np.array_equal(z[:, 8:, 20:], y[:, 8:, 20:])
