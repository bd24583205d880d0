"""Oracle: on-the-fly front end (audio augmentation, Kaldi fbank, splice, batch assembly,
CMN/CMVN, SpecAugment), numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Paths relative to /root/reference.

fbank: the reference calls PyKaldi ``Fbank.compute_features`` (loader/otf_utt_loader.py:195-201,
231-234) with egs/fbank.conf; PyKaldi/Kaldi are un-vendored and un-pinned (README.md:30-32).
``kaldi_fbank`` restates Kaldi's published algorithm (feature-window.cc ProcessWindow,
feature-fbank.cc Compute, mel-computations.cc MelBanks) and is pinned against
``torchaudio.compliance.kaldi.fbank`` in tests/test_oracle_frontend.py.  dither = 0 for parity.
"""
import numpy as np


# ---------------------------------------------------------------- loader/audio.py
def to_float32(samples_i16):
    """AudioSegment._convert_samples_to_float32 (loader/audio.py:562-576): int16 -> f32 * 2^-15."""
    return samples_i16.astype(np.float32) * np.float32(1.0 / 2 ** 15)


def change_speed(samples, rate):
    """AudioSegment.change_speed (loader/audio.py:217-238): linear-interp resample; result is
    float64 (np.interp promotes) unless rate == 1.0."""
    if rate <= 0:
        raise ValueError("speed_rate should be greater than zero.")
    if rate == 1.0:
        return samples
    old_length = samples.shape[0]
    new_length = int(old_length / rate)
    old_indices = np.arange(old_length)
    new_indices = np.linspace(start=0, stop=old_length, num=new_length)
    return np.interp(new_indices, old_indices, samples)


def rms_db(samples):
    """AudioSegment.rms_db (loader/audio.py:551-560)."""
    mean_square = max(1e-20, np.mean(samples ** 2))
    return 10 * np.log10(mean_square)


def normalize(samples, target_db, max_gain_db=300.0):
    """AudioSegment.normalize + gain_db (loader/audio.py:240-262, 207-215)."""
    gain = target_db - rms_db(samples)
    if gain > max_gain_db:
        raise ValueError("Unable to normalize segment to %f dB" % target_db)
    return samples * 10. ** (min(max_gain_db, gain) / 20.)


def to_int16(samples):
    """AudioSegment._convert_samples_from_float32(.., 'int16') (loader/audio.py:578-603):
    scale by 2^15, clip to [-32768, 32767], C-cast truncation toward zero."""
    out = samples.copy() * (2 ** 15)
    out[out > 32767] = 32767
    out[out < -32768] = -32768
    return out.astype(np.int16)


def augment(pcm_i16, rate, target_db):
    """loader/otf_utt_loader.py:218-230: int16 -> float -> speed -> gain -> int16."""
    s = to_float32(pcm_i16)
    s = change_speed(s, rate)
    s = normalize(s, target_db)
    return to_int16(s)


# ---------------------------------------------------------------- Kaldi fbank
def _mel(f):
    return 1127.0 * np.log(1.0 + f / 700.0)


def mel_banks(num_bins=80, sample_freq=16000.0, low_freq=40.0, high_freq=-200.0, n_fft=512):
    """Kaldi MelBanks (no VTLN): triangular filters, equally spaced on the mel scale.
    Returns [num_bins, n_fft/2] float32 weights (the Nyquist bin is never used)."""
    nyquist = 0.5 * sample_freq
    if high_freq <= 0:
        high_freq += nyquist
    fft_bin_width = sample_freq / n_fft
    mel_lo, mel_hi = _mel(low_freq), _mel(high_freq)
    delta = (mel_hi - mel_lo) / (num_bins + 1)
    nb = n_fft // 2
    w = np.zeros((num_bins, nb), dtype=np.float64)
    mel_f = _mel(fft_bin_width * np.arange(nb))
    for j in range(num_bins):
        left, center, right = mel_lo + j * delta, mel_lo + (j + 1) * delta, mel_lo + (j + 2) * delta
        up = (mel_f - left) / (center - left)
        down = (right - mel_f) / (right - center)
        inside = (mel_f > left) & (mel_f < right)
        w[j] = np.where(inside, np.where(mel_f <= center, up, down), 0.0)
    return w.astype(np.float32)


def hamming_window(n=400):
    i = np.arange(n, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2.0 * np.pi * i / (n - 1))).astype(np.float32)


def num_frames(n_samples, frame_len=400, frame_shift=160):
    """snip-edges frame count."""
    return 0 if n_samples < frame_len else 1 + (n_samples - frame_len) // frame_shift


def kaldi_fbank(wave, num_bins=80, frame_len=400, frame_shift=160, preemph=0.97, n_fft=512):
    """wave: 1-D array of int16-scaled samples -> [T, num_bins] float32 log-mel energies.
    Options = egs/fbank.conf (hamming, 16 kHz, low 40, high -200, 80 bins) + Kaldi defaults
    (25 ms / 10 ms, snip-edges, remove-dc-offset, preemphasis 0.97, power spectrum, log, no energy)."""
    wave = np.asarray(wave, dtype=np.float32)
    T = num_frames(wave.shape[0], frame_len, frame_shift)
    if T == 0:
        return np.zeros((0, num_bins), np.float32)
    idx = np.arange(T)[:, None] * frame_shift + np.arange(frame_len)[None, :]
    fr = wave[idx].astype(np.float32)                                # [T, 400]
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=np.float32)       # remove_dc_offset
    pre = np.empty_like(fr)
    pre[:, 1:] = fr[:, 1:] - np.float32(preemph) * fr[:, :-1]
    pre[:, 0] = fr[:, 0] - np.float32(preemph) * fr[:, 0]
    pre = pre * hamming_window(frame_len)[None, :]
    spec = np.fft.rfft(pre.astype(np.float64), n=n_fft, axis=1)
    power = (spec.real ** 2 + spec.imag ** 2).astype(np.float32)[:, : n_fft // 2]
    mel = power @ mel_banks(num_bins, n_fft=n_fft).T
    return np.log(np.maximum(mel, np.finfo(np.float32).eps)).astype(np.float32)


# ---------------------------------------------------------------- loader/otf_utt_loader.py
def splice(feats, lctx, rctx):
    """splice (loader/otf_utt_loader.py:28-46): +-ctx frame stacking with edge replication."""
    length, dim = feats.shape
    padding = np.zeros((length + lctx + rctx, dim), dtype=np.float32)
    padding[:lctx] = feats[0]
    padding[lctx:lctx + length] = feats
    padding[lctx + length:] = feats[-1]
    spliced = np.zeros((length, dim * (lctx + 1 + rctx)), dtype=np.float32)
    for i in range(lctx + 1 + rctx):
        spliced[:, i * dim:(i + 1) * dim] = padding[i:i + length, :]
    return spliced


def assemble_batch(feats_list, labels_list, lctx=1, rctx=1, stride=1, tu_limit=15000, padding_tgt=-1):
    """otf_utt_generator batch assembly (loader/otf_utt_loader.py:239-289): TU filter (:247),
    splice + fill (:248-257), pad data with the LAST VALID FRAME and labels with padding_tgt
    (:262-270).  Returns (data [B,Tmax,D] f32, target [B,Umax] i32, lens [B] i32, ali_lens [B] i32)
    or (None, None, [0], [0]) when every utterance was filtered (:283-287)."""
    kept = []
    for f, a in zip(feats_list, labels_list):
        utt_len = f.shape[0] // stride + int(f.shape[0] % stride != 0)
        if len(a) * utt_len // 3 <= tu_limit:
            kept.append((splice(f, lctx, rctx)[::stride], np.asarray(a, np.int32), utt_len))
    if not kept:
        return None, None, np.zeros(1, np.int32), np.zeros(1, np.int32)
    tmax = max(k[2] for k in kept)
    umax = max(len(k[1]) for k in kept)
    D = kept[0][0].shape[1]
    data = np.zeros((len(kept), tmax, D), np.float32)
    target = np.full((len(kept), umax), padding_tgt, np.int32)
    lens = np.zeros(len(kept), np.int32)
    ali = np.zeros(len(kept), np.int32)
    for b, (s, a, L) in enumerate(kept):
        data[b, :L] = s
        data[b, L:] = s[L - 1]
        target[b, :len(a)] = a
        lens[b], ali[b] = L, len(a)
    return data, target, lens, ali


# ---------------------------------------------------------------- trainer
def cmvn_from_stats(stats, splice_width=3):
    """trainer/train_transducer_bmuf_otfaug.py:341-355: Kaldi 2x(D+1) double stats ->
    (offset, scale) float64, tiled lctx+1+rctx times."""
    stats = np.asarray(stats, np.float64)
    mean = stats[0][:-1] / stats[0][-1]
    var = stats[1][:-1] / stats[0][-1] - mean * mean
    return np.tile(-mean, splice_width), np.tile(1.0 / np.sqrt(var), splice_width)


def apply_cmvn(data, offset, scale, cmn=True):
    """trainer/train_transducer_bmuf_otfaug.py:86-91 (CMN mean is over the PADDED time axis)."""
    x = data.astype(np.float32).copy()
    if cmn:
        x -= x.mean(axis=1, dtype=np.float32, keepdims=True)
    x += offset.astype(np.float32)[None, None, :]
    x *= scale.astype(np.float32)[None, None, :]
    return x


def spec_augment(x, freq_start, freq_span, time_start, time_span):
    """SpecAugment.apply (utils/spec_augment.py:10-20) with the random draws made by the caller:
    one freq mask over the spliced feature axis and one time mask, shared by the whole batch."""
    x = x.copy()
    if freq_span > 0:
        x[:, :, freq_start:freq_start + freq_span] = 0.0
    if time_span > 0:
        x[:, time_start:time_start + time_span, :] = 0.0
    return x
