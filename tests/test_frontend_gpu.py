"""GPU parity of the on-the-fly front end against the fixtures produced by the reference's own
AudioSegment / splice code and torchaudio's Kaldi fbank (tests/golden/frontend.npz), and against the
numpy oracle for CMN/CMVN/SpecAugment and last-frame padding.  Integer path (augmented int16 samples):
bit-exact for the speed-perturbed (float64) branch; <= 1 LSB on the rate == 1.0 (float32) branch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_frontend():
    from pika_b200.frontend import FbankOptions, Frontend
    opts = FbankOptions(num_mel_bins=80, low_freq=40.0, high_freq=-200.0, dither=0.0, window_type="hamming")
    return Frontend(opts, 1, 1, "cuda")


def run(fe, pcms, rates, dbs, **kw):
    from pika_b200.frontend import Frontend
    B = len(pcms)
    n = [len(p) for p in pcms]
    new_len, frames = Frontend.lengths(n, rates)
    n_max = max(max(n), max(new_len))
    pcm = torch.zeros(B, n_max, dtype=torch.int16)
    for i, p in enumerate(pcms):
        pcm[i, :len(p)] = torch.from_numpy(p)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device="cuda")
    f32 = lambda v: torch.tensor(v, dtype=torch.float32, device="cuda")
    t_max = kw.pop("t_max", max(frames))
    out, wave = fe(pcm.cuda(), i32(n), f32(rates), f32(dbs), i32(new_len), i32(frames), t_max, want_wave=True, **kw)
    torch.cuda.synchronize()
    return out.float().cpu().numpy(), wave.cpu().numpy(), new_len, frames


def test_augment_fbank_splice_vs_reference_golden(golden_dir):
    d = np.load(os.path.join(golden_dir, "frontend.npz"))
    fe = make_frontend()
    keys, rates, dbs = ["r09", "r10", "r11"], [0.9, 1.0, 1.1], [-23.5, -41.0, -12.25]
    out, wave, new_len, frames = run(fe, [d["pcm"]] * 3, rates, dbs, cmn=False)
    for i, k in enumerate(keys):
        aug = d["aug_" + k]
        assert new_len[i] == len(aug) and frames[i] == d["fbank_" + k].shape[0]
        diff = np.abs(wave[i, :len(aug)].astype(np.int32) - aug.astype(np.int32))
        if rates[i] == 1.0:
            assert diff.max() <= 1 and (diff != 0).mean() < 5e-3     # float32 branch: see module docstring
        else:
            assert diff.max() == 0                                   # float64 branch: bit-exact
        ref = d["splice_" + k]                                       # every 5th spliced frame
        got = out[i, :frames[i]][::5]
        np.testing.assert_allclose(got, ref, atol=5e-3)
        assert np.abs(got - ref).mean() < 2e-4
        # padding beyond the utterance replicates the last valid spliced frame (loader/otf_utt_loader.py:262-266)
        if frames[i] < out.shape[1]:
            np.testing.assert_array_equal(out[i, frames[i]:], np.broadcast_to(out[i, frames[i] - 1], out[i, frames[i]:].shape))


def test_cmn_cmvn_specaug_vs_oracle(golden_dir):
    from oracle import frontend as ofe
    d = np.load(os.path.join(golden_dir, "frontend.npz"))
    fe = make_frontend()
    rng = np.random.default_rng(3)
    pcm2 = np.clip(np.round(rng.normal(0, 2500, 400 + 160 * 39)), -32768, 32767).astype(np.int16)
    pcms, rates, dbs = [d["pcm"], pcm2], [1.1, 0.9], [-20.0, -30.0]
    raw, wave, new_len, frames = run(fe, pcms, rates, dbs, cmn=False)
    stats = np.zeros((2, 81))
    mean, var, n = rng.standard_normal(80) * 3 + 8, np.abs(rng.standard_normal(80)) + 1.0, 1000.0
    stats[0, :80], stats[0, 80], stats[1, :80] = mean * n, n, (var + mean * mean) * n
    off, sc = ofe.cmvn_from_stats(stats)
    sa = (100, 9, 11, 17)
    out, _, _, _ = run(fe, pcms, rates, dbs, cmn=True, offset=torch.tensor(off, dtype=torch.float32, device="cuda"),
                       scale=torch.tensor(sc, dtype=torch.float32, device="cuda"), specaug=sa)
    ref = ofe.spec_augment(ofe.apply_cmvn(raw, off, sc, cmn=True), *sa)
    np.testing.assert_allclose(out, ref, atol=2e-4)
    assert np.all(out[:, :, 100:109] == 0) and np.all(out[:, 11:28, :] == 0)
    # oracle chain from the augmented samples (numpy Kaldi restatement)
    fb = ofe.kaldi_fbank(wave[1, :new_len[1]].astype(np.float32))
    np.testing.assert_allclose(raw[1, :frames[1]], ofe.splice(fb, 1, 1), atol=5e-3)


def test_bf16_output_and_batch_of_full_length():
    fe = make_frontend()
    rng = np.random.default_rng(4)
    n = 400 + 160 * 99
    pcms = [np.clip(np.round(rng.normal(0, 3000, n)), -32768, 32767).astype(np.int16) for _ in range(4)]
    o32, _, _, frames = run(fe, pcms, [1.0] * 4, [-25.0] * 4, cmn=True)
    o16, _, _, _ = run(fe, pcms, [1.0] * 4, [-25.0] * 4, cmn=True, out_dtype=torch.bfloat16)
    assert frames == [100] * 4 and o32.shape == (4, 100, 240)
    np.testing.assert_allclose(o16, o32, atol=2e-2, rtol=1e-2)
    assert abs(o32.mean(axis=1)).max() < 1e-4            # CMN: zero mean over time


def test_loader_feature_mode_matches_oracle(tmp_path):
    """drop-in loader in the reference's feature-yielding mode: (data [B,T,240] f32 CPU, target, lens, ali_lens)
    with the features computed by the GPU front end, against the numpy oracle chain on the same draws."""
    import random
    from oracle import frontend as ofe
    from test_loader_cpu import loader_args, make_dataset
    from pika_b200.loader import otf_utt_loader as L
    lst, utts = make_dataset(tmp_path, n_utts=4, shards=1)
    cfg = tmp_path / "fbank.conf"
    cfg.write_text("--window-type=hamming\n--sample-frequency=16000\n--dither=1\n--low-freq=40\n--high-freq=-200\n--num-mel-bins=80\n")
    a = loader_args(raw_batches=False, batch_first=True, feat_config=str(cfg))
    a.no_dither = True              # the recipe config asks for dither = 1; parity against the (deterministic) oracle chain opts out
    random.seed(3); np.random.seed(3)
    (data, target, lens, ali_lens), = list(L.dataloader(lst, [], [], a))
    assert data.dtype == torch.float32 and not data.is_cuda and tuple(data.shape) == (4, int(lens.max()), 240)
    random.seed(3); np.random.seed(3)
    for i, (pcm, lab) in enumerate(utts):
        spr = [0.9, 1.0, 1.1][random.randint(0, 2)]
        gain = np.random.uniform(-50.0, -10.0)
        fb = ofe.kaldi_fbank(ofe.augment(pcm, spr, np.float32(gain)).astype(np.float32))
        ref = ofe.splice(fb, 1, 1)
        assert int(lens[i]) == ref.shape[0]
        np.testing.assert_allclose(data[i, :ref.shape[0]].numpy(), ref, atol=1e-2)
        assert np.abs(data[i, :ref.shape[0]].numpy() - ref).mean() < 5e-4


def test_dither_is_gaussian_counter_based_and_off_by_default_in_parity_runs():
    """FbankOptions.dither (egs/fbank.conf: dither=1): dither * N(0,1) per sample of every window.  Checked on silence, where the
    features are the dither alone: the log-mel energies must match white noise of variance dither^2 through the same window / mel
    filters; same seed -> identical features, new seed -> new noise, dither = 0 -> exactly the undithered features."""
    from pika_b200.frontend import FbankOptions, Frontend
    dev = torch.device("cuda", 0)
    fe = Frontend(FbankOptions(num_mel_bins=80, low_freq=40.0, high_freq=-200.0, dither=1.0, window_type="hamming"), 1, 1, dev)
    B, T = 2, 300
    n = 400 + (T - 1) * 160
    nf = torch.full((B,), T, dtype=torch.int32, device=dev)
    silence = torch.zeros(B, n, device=dev)
    f1 = fe.fbank(silence, nf, T, dither=2.0, seed=123)
    f2 = fe.fbank(silence, nf, T, dither=2.0, seed=123)
    f3 = fe.fbank(silence, nf, T, dither=2.0, seed=124)
    assert torch.equal(f1, f2) and not torch.equal(f1, f3)
    # expected mel energy of white noise of variance sigma^2 = 4: every FFT bin k carries sigma^2 * sum(window^2) * |1 - c e^{-jw_k}|^2
    # (pre-emphasis x[i] - c x[i-1] shapes the flat spectrum; DC removal only touches bin 0), summed through the mel weights
    win = fe.window.double().cpu().numpy()
    c = fe.opts.preemphasis_coefficient
    mel = fe.mel_w.double().cpu().numpy()
    wk = 2.0 * np.pi * np.arange(mel.shape[1]) / 512.0
    gain = 1.0 + c * c - 2.0 * c * np.cos(wk)
    expect = np.log(4.0 * (win ** 2).sum() * (mel * gain[None, :]).sum(1))
    got = np.log(np.exp(f1.double().cpu().numpy()).mean((0, 1)))            # log of the MEAN mel energy over the 600 frames
    # (the mean of the logs sits below it by up to Euler's 0.577 for the one-bin low filters: a chi-square with 2 degrees of freedom)
    assert np.abs(got - expect).max() < 0.15, (got[:8], expect[:8])
    loud = (torch.randn(B, n, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 3000).round()
    g0 = fe.fbank(loud, nf, T, dither=0.0, seed=1)
    g1 = fe.fbank(loud, nf, T, dither=1.0, seed=1)
    diff = (g1 - g0).abs().max().item()
    assert 0.0 < diff < 0.05                         # 1 LSB-class noise against a 3000-amplitude signal moves log energies by < 5e-2


def test_compute_global_cmvn_cli_matches_oracle(tmp_path):
    """utils/compute_global_cmvn.py drop-in: the stats file (sums, sums of squares, frame count) over a synthetic .lst corpus against the
    numpy oracle (oracle/frontend.py: augmentation + Kaldi fbank) fed with the same augmentation draws; the file must read back through
    the trainer's CMVN reader"""
    import random
    from oracle import frontend as ofe
    from test_loader_cpu import make_dataset
    from pika_b200.loader import kaldi_io
    from pika_b200.utils import compute_global_cmvn as C
    lst, utts = make_dataset(tmp_path, n_utts=6, shards=2, n_lo=6000, n_hi=12000)
    cfg = tmp_path / "fbank.conf"
    cfg.write_text("--window-type=hamming\n--sample-frequency=16000\n--dither=0\n--low-freq=40\n--high-freq=-200\n--num-mel-bins=80\n")
    out = tmp_path / "cmvn.stats"
    random.seed(11); np.random.seed(11)
    C.main([lst, str(out), "--feat_config", str(cfg), "--cmn", "--batch_size", "4"])
    stats = kaldi_io.read_kaldi_text_matrix(str(out))
    assert stats.shape == (2, 81) and stats[1, 80] == 0.0
    random.seed(11); np.random.seed(11)
    s1, s2, cnt = np.zeros(80), np.zeros(80), 0
    for mrk_fn, seq_fn in [l.split()[:2] for l in open(lst)]:
        for _, audio in kaldi_io.iter_mrk_seq(mrk_fn, seq_fn):
            rate = [0.9, 1.0, 1.1][random.randint(0, 2)]
            gain = float(np.random.uniform(-55, -10))
            f = ofe.kaldi_fbank(ofe.augment(np.asarray(audio, np.int16), rate, gain).astype(np.float32)).astype(np.float64)
            f = f - f.mean(axis=0, keepdims=True)
            s1 += f.sum(0); s2 += (f * f).sum(0); cnt += f.shape[0]
    assert stats[0, 80] == cnt
    np.testing.assert_allclose(stats[1, :80], s2, rtol=2e-3)                    # second moments of the mean-removed log-mel energies
    assert np.abs(stats[0, :80] - s1).max() < 1e-2 * cnt                        # per-utterance CMN: the sums are ~0 on both sides
    off, scale = kaldi_io.cmvn_offset_scale(str(out), 3)
    assert off.shape == (240,) and np.isfinite(scale).all()
