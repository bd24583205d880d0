"""Host-side loader logic (no GPU): native readers for the recipe's on-disk formats, TU filter, padding,
end-of-stream protocol and the augmentation draws coming from the reference's RNG streams."""
import argparse
import os
import random
import struct

import numpy as np
import pytest
import torch


def make_dataset(tmp_path, n_utts=7, shards=2, n_lo=3000, n_hi=9000):
    rng = np.random.default_rng(0)
    lines = []
    utts = []
    k = 0
    for s in range(shards):
        mrk, seq, ark = tmp_path / ("d%d.mrk" % s), tmp_path / ("d%d.seq" % s), tmp_path / ("d%d.ark" % s)
        off = 0
        with open(mrk, "w") as fm, open(seq, "wb") as fs, open(ark, "w") as fa:
            for _ in range(n_utts):
                n = int(rng.integers(n_lo, n_hi))
                pcm = rng.integers(-3000, 3000, n).astype(np.int16)
                lab = rng.integers(1, 50, int(rng.integers(1, 9))).tolist()
                fs.write(pcm.tobytes())
                fm.write("utt%03d %d %d\n" % (k, off, n * 2))
                fa.write("utt%03d %s\n" % (k, " ".join(str(v) for v in lab)))
                off += n * 2
                utts.append((pcm, lab))
                k += 1
        lines.append("%s %s ark:%s" % (mrk, seq, ark))
    lst = tmp_path / "data.lst"
    lst.write_text("\n".join(lines) + "\n")
    return str(lst), utts


def loader_args(**kw):
    from pika_b200.loader import otf_utt_loader as L
    p = argparse.ArgumentParser()
    L.register(p)
    a = p.parse_args([])
    a.lctx, a.rctx, a.feats_dim, a.batch_size, a.num_workers, a.padding_tgt = 1, 1, 80, 4, 1, 99
    a.max_len, a.TU_limit, a.gain_range, a.raw_batches = 1600, 15000, "50,10", True
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_register_and_inputdim_match_reference_flags():
    from pika_b200.loader import otf_utt_loader as L
    a = loader_args()
    assert L.get_inputdim(a) == 240
    for flag in ("lctx", "rctx", "max_len", "num_workers", "sample_rate", "buffer_size", "batch_first", "reverse_labels",
                 "feat_config", "stride", "batch_size", "SOS", "EOS", "queue_size", "TU_limit", "padding_tgt", "feats_dim",
                 "snr_range", "gain_range", "speed_rate", "verbose"):
        assert hasattr(a, flag), flag


def test_raw_batches_follow_reference_protocol(tmp_path):
    from pika_b200.frontend import Frontend
    from pika_b200.loader import otf_utt_loader as L
    lst, utts = make_dataset(tmp_path)
    a = loader_args()
    random.seed(5); np.random.seed(5)
    batches = list(L.dataloader(lst, [], [], a))
    # 14 utterances, batch 4 -> 3 full batches (the trailing partial batch is dropped, as in the reference)
    assert len(batches) == 3
    # replay the reference's draw order: per utterance random.randint then np.random.uniform
    random.seed(5); np.random.seed(5)
    k = 0
    for raw, target, lens, ali_lens in batches:
        B = raw["pcm"].shape[0]
        assert target.dtype == torch.int32 and lens.dtype == torch.int32
        for i in range(B):
            pcm, lab = utts[k]
            spr = [0.9, 1.0, 1.1][random.randint(0, 2)]
            gain = np.random.uniform(-50.0, -10.0)
            n = int(raw["n_samples"][i])
            assert n == len(pcm) and np.array_equal(raw["pcm"][i, :n].numpy(), pcm)
            assert abs(float(raw["rate"][i]) - spr) < 1e-7 and abs(float(raw["target_db"][i]) - gain) < 1e-5
            new_len, frames = Frontend.lengths([n], [spr])
            assert int(raw["new_len"][i]) == new_len[0] and int(lens[i]) == frames[0]
            assert target[i, :len(lab)].tolist() == lab and all(v == 99 for v in target[i, len(lab):].tolist())
            assert int(ali_lens[i]) == len(lab)
            k += 1


def test_tu_filter_and_empty_batch(tmp_path):
    from pika_b200.loader import otf_utt_loader as L
    lst, utts = make_dataset(tmp_path, n_utts=4, shards=1)
    a = loader_args(TU_limit=0)
    out = list(L.dataloader(lst, [], [], a))
    assert len(out) == 1 and out[0][0] is None and out[0][2].tolist() == [0] and out[0][3].tolist() == [0]
    a = loader_args(num_workers=2)
    lst2, _ = make_dataset(tmp_path, n_utts=4, shards=2)
    assert len(list(L.dataloader(lst2, [], [], a))) == 2          # one batch per worker, two None sentinels consumed


def test_kaldi_readers(tmp_path):
    from pika_b200.loader import kaldi_io
    t = tmp_path / "t.ark"
    t.write_text("a 1 2 3\nb \nc 7\n")
    assert list(kaldi_io.read_int_vector_ark("ark,t:%s" % t)) == [("a", [1, 2, 3]), ("b", []), ("c", [7])]
    b = tmp_path / "b.ark"
    with open(b, "wb") as f:
        for key, vals in (("k1", [5, 6]), ("k2", [])):
            f.write(key.encode() + b" \0B\x04" + struct.pack("<i", len(vals)))
            for v in vals:
                f.write(b"\x04" + struct.pack("<i", v))
    assert list(kaldi_io.read_int_vector_ark("ark:%s" % b)) == [("k1", [5, 6]), ("k2", [])]
    m = tmp_path / "cmvn"
    m.write_text(" [\n  10 20 5 \n  30 100 0 ]\n")
    off, sc = kaldi_io.cmvn_offset_scale(str(m), 3)
    mean, var = np.array([2.0, 4.0]), np.array([30 / 5 - 4.0, 100 / 5 - 16.0])
    np.testing.assert_allclose(off, np.tile(-mean, 3))
    np.testing.assert_allclose(sc, np.tile(1 / np.sqrt(var), 3))


# ------------------------------------------------------------------------------------------------ offline feature loader (decode CLI)
def _feature_tables(tmp_path, n_utts=5, dim=6, seed=3):
    from pika_b200.loader.kaldi_io import write_float_matrix_ark
    rng = np.random.default_rng(seed)
    feats = [("utt%02d" % i, rng.standard_normal((int(rng.integers(9, 23)), dim)).astype(np.float32)) for i in range(n_utts)]
    labels = [[int(v) for v in rng.integers(1, 30, int(rng.integers(1, 6)))] for _ in range(n_utts)]
    ark = tmp_path / "feats.ark"
    offs = write_float_matrix_ark(str(ark), feats)
    (tmp_path / "feats.scp").write_text("".join("%s %s:%d\n" % (k, ark, offs[k]) for k, _ in feats))
    write_float_matrix_ark(str(tmp_path / "feats.txt.ark"), feats, text=True)
    (tmp_path / "labels.ark").write_text("".join("%s %s\n" % (k, " ".join(str(v) for v in l)) for (k, _), l in zip(feats, labels)))
    return feats, labels


def test_float_matrix_tables_roundtrip(tmp_path):
    """binary 'FM ' archive, text archive and scp (key file:offset) all give back the matrices, in file order"""
    from pika_b200.loader.kaldi_io import read_float_matrix_table
    feats, _ = _feature_tables(tmp_path)
    for rspec, tol in (("ark:%s" % (tmp_path / "feats.ark"), 0.0), ("scp:%s" % (tmp_path / "feats.scp"), 0.0),
                       ("ark,t:%s" % (tmp_path / "feats.txt.ark"), 1e-5)):
        got = list(read_float_matrix_table(rspec))
        assert [k for k, _ in got] == [k for k, _ in feats]
        for (_, a), (_, b) in zip(feats, got):
            assert a.shape == b.shape and np.abs(a - b).max() <= tol


def test_utt_loader_batches_follow_reference_protocol(tmp_path):
    """loader/utt_loader.py:154-232: splice with edge repetition, stride, data padded with each utterance's last frame, labels with
    padding_tgt, only FULL batches are emitted, iteration ends after the generator's None"""
    from pika_b200.loader import utt_loader as UL
    feats, labels = _feature_tables(tmp_path)
    p = argparse.ArgumentParser()
    UL.register(p)
    a = p.parse_args(["--lctx", "1", "--rctx", "2", "--max_len", "40", "--batch_size", "2", "--padding_tgt", "33", "--feats_dim", "6",
                      "--batch_first", "--stride", "2"])
    a.cuda, a.local_rank = False, 0
    assert UL.get_inputdim(a) == 6 * 4
    batches = list(UL.dataloader("ark,t:%s" % (tmp_path / "labels.ark"), "scp:%s" % (tmp_path / "feats.scp"), False, a))
    assert len(batches) == 2                                   # 5 utterances, batch 2: the fifth never fills a batch
    for bi, (data, target, lens, ali_lens) in enumerate(batches):
        assert data.dtype == torch.float32 and target.dtype == torch.int64
        for r in range(2):
            f, l = feats[2 * bi + r][1], labels[2 * bi + r]
            n = f.shape[0]
            ref = np.stack([np.concatenate([f[min(max(t + o, 0), n - 1)] for o in (-1, 0, 1, 2)]) for t in range(n)])[::2]
            assert lens[r] == ref.shape[0] and ali_lens[r] == len(l)
            np.testing.assert_array_equal(data[r, :lens[r]].numpy(), ref)
            assert bool((data[r, lens[r]:] == torch.from_numpy(ref[-1])).all())          # last valid frame repeated
            assert target[r, :len(l)].tolist() == l and bool((target[r, len(l):] == 33).all())
        assert data.shape[1] == int(max(lens)) and target.shape[1] == int(max(ali_lens))


def test_read_fst_text(tmp_path):
    from pika_b200.decoder.sorted_matcher import SortedMatcher, read_fst_text
    (tmp_path / "g.txt").write_text("0 1 5 5 0.5\n0 2 3 3 1.25\n1 0 1 1 0.75\n1 2 4 4\n2 1.5\n0\n")
    arcs, finals = read_fst_text(str(tmp_path / "g.txt"))
    assert arcs[0] == [(3, 1.25, 2), (5, 0.5, 1)] and arcs[1] == [(1, 0.75, 0), (4, 0.0, 2)] and arcs[2] == []
    assert finals[0] == 0.0 and finals[2] == 1.5 and finals[1] == float("inf")
    m = SortedMatcher((arcs, finals), 2, 6, 1, [])
    sc, st = m.get_scores(1, 3)                                # no arc 3 in state 1 -> back off (label 1) to state 0, then arc 3
    assert st == [2] and abs(sc[0] - 2.0) < 1e-12


def _kaldi_compress_format1(mat):
    """CompressedMatrix format 1 as Kaldi writes it (matrix/compressed-matrix.cc: ComputeColHeader / CompressColumn), restated for the
    test: global (min, range), per-column 0/25/75/100 percentiles as uint16, bytes piecewise linear between them"""
    rows, cols = mat.shape
    vmin, vmax = float(mat.min()), float(mat.max())
    vrange = max(vmax - vmin, 1e-30)
    to_u16 = lambda v: int(np.clip(np.floor((v - vmin) / vrange * 65535.0 + 0.499), 0, 65535))      # noqa: E731
    hdr, data = [], np.zeros((cols, rows), np.uint8)
    for c in range(cols):
        col = np.sort(mat[:, c])
        q = [col[0], col[rows // 4], col[3 * rows // 4], col[-1]]
        u = [to_u16(v) for v in q]
        u[0] = min(u[0], 65532)                                   # strictly increasing percentiles, as Kaldi enforces them
        u[1] = min(max(u[1], u[0] + 1), 65533); u[2] = min(max(u[2], u[1] + 1), 65534); u[3] = min(max(u[3], u[2] + 1), 65535)
        p = [vmin + vrange * x / 65535.0 for x in u]
        hdr.append(u)
        for r in range(rows):
            v = float(mat[r, c])
            if v < p[1]:
                b = int(np.clip(np.floor((v - p[0]) / (p[1] - p[0]) * 64 + 0.5), 0, 64))
            elif v < p[2]:
                b = int(np.clip(np.floor(64 + (v - p[1]) / (p[2] - p[1]) * 128 + 0.5), 64, 192))
            else:
                b = int(np.clip(np.floor(192 + (v - p[2]) / (p[3] - p[2]) * 63 + 0.5), 192, 255))
            data[c, r] = b
    return struct.pack("<ffii", vmin, vrange, rows, cols) + np.asarray(hdr, np.uint16).tobytes() + data.tobytes()


def test_compressed_feature_archives(tmp_path):
    """Kaldi's feature dumps are compressed by default (make_fbank.sh, copy-feats --compress=true): 'CM ' (one byte per element with
    per-column percentile headers), 'CM2' (uint16) and 'CM3' (uint8) decode to within their quantisation step"""
    from pika_b200.loader.kaldi_io import read_float_matrix_table
    rng = np.random.default_rng(4)
    mat = (rng.standard_normal((37, 8)) * 3.0 + 5.0).astype(np.float32)
    vmin, vrange = float(mat.min()), float(mat.max() - mat.min())
    ark = tmp_path / "cm.ark"
    with open(ark, "wb") as f:
        f.write(b"u1 \0BCM " + _kaldi_compress_format1(mat))
        f.write(b"u2 \0BCM2 " + struct.pack("<ffii", vmin, vrange, 37, 8) + np.round((mat - vmin) / vrange * 65535).astype(np.uint16).tobytes())
        f.write(b"u3 \0BCM3 " + struct.pack("<ffii", vmin, vrange, 37, 8) + np.round((mat - vmin) / vrange * 255).astype(np.uint8).tobytes())
    got = dict(read_float_matrix_table("ark:%s" % ark))
    assert list(got) == ["u1", "u2", "u3"] and all(v.shape == (37, 8) and v.dtype == np.float32 for v in got.values())
    col_rng = mat.max(0) - mat.min(0)
    assert (np.abs(got["u1"] - mat).max(0) <= col_rng / 64.0 + 1e-3).all()       # coarsest segment: a quarter of the column range / 64 steps
    assert np.abs(got["u2"] - mat).max() <= vrange / 65535.0 + 1e-5
    assert np.abs(got["u3"] - mat).max() <= vrange / 255.0 * 0.51 + 1e-5
