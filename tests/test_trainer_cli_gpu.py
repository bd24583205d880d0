"""The drop-in training entry point end to end on a tiny synthetic corpus: recipe-style flags, raw .seq/.mrk/.lst
shards + label ark + Kaldi CMVN file -> loader threads -> GPU front end -> two epochs with BMUF syncs -> per-rank
log and pickled model files with the reference's names; the loss must go down and the pickles must reload."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("decoder_type", ["rnn", "transformer"])
def test_train_cli_two_epochs(tmp_path, decoder_type):
    """``--decoder_type transformer`` trains the convolutional-transformer prediction net (trainer/model/transducer.py:62-68)"""
    from test_loader_cpu import make_dataset
    from pika_b200.trainer import train_transducer_bmuf_otfaug as T
    lst, utts = make_dataset(tmp_path, n_utts=8, shards=1, n_lo=14000, n_hi=22000)
    cfg = tmp_path / "fbank.conf"
    cfg.write_text("--window-type=hamming\n--sample-frequency=16000\n--dither=1\n--low-freq=40\n--high-freq=-200\n--num-mel-bins=80\n")
    cmvn = tmp_path / "cmvn.stats"
    rng = np.random.default_rng(0)
    mean, var, n = rng.standard_normal(80) + 8.0, np.abs(rng.standard_normal(80)) + 4.0, 1000.0
    cmvn.write_text(" [\n  %s %g\n  %s 0 ]\n" % (" ".join("%g" % v for v in mean * n), n, " ".join("%g" % v for v in (var + mean * mean) * n)))
    out = tmp_path / "out"
    out.mkdir()
    log = tmp_path / "log.WORKER-ID"
    argv = ["transducer", lst.replace("data.lst", "data.lst"), str(log), str(out), "--cuda", "--local_rank", "0", "--encoder_type", "transformer",
            "--decoder_type", decoder_type, "--rnn_size", "1024", "--embd_dim", "100", "--output_dim", "60", "--padding_idx", "60", "--padding_tgt", "60",
            "--dec_layers", "2", "--dropout", "0.0", "--brnn", "--model_lctx", "21", "--model_rctx", "21", "--model_stride", "4",
            "--lctx", "1", "--rctx", "1", "--feats_dim", "80", "--feat_config", str(cfg), "--cmn", "--cmvn_stats", str(cmvn), "--batch_size", "4",
            "--num_workers", "1", "--batch_first", "--max_len", "1600", "--TU_limit", "50000", "--gain_range", "25,25", "--speed_rate", "1.0", "--grad_clip", "3.0",
            "--initial_lr", "0.002", "--final_lr", "0.001", "--momentum", "0.9", "--num_epochs", "5", "--num_batches_per_epoch", "2",
            "--sync_period", "1", "--block_momentum", "0.9", "--block_lr", "1.0", "--spec_augment", "--seed", "777"]
    os.environ.setdefault("WORLD_SIZE", "1")
    T.main(argv)
    text = open(str(log).replace("WORKER-ID", "0")).read()
    assert "===> Epoch 0 <===" in text and "===> Epoch 1 <===" in text and "Training Finished" in text
    losses = [float(l.split("Loss:")[1].split()[0]) for l in text.splitlines() if "Overall Avg Loss" in l]
    # same 8 utterances every epoch, augmentation draws fixed, dropout off (SpecAugment stays on): ten SGD steps must lower the loss
    assert len(losses) == 5 and np.isfinite(losses).all() and min(losses[-2:]) < losses[0]
    for e in (0, 1, 2, 3, 4):
        path = out / ("model.epoch.%d.0" % e)
        assert path.exists()
    m = torch.load(str(out / "model.epoch.4.0"), weights_only=False)
    assert m.fc2.weight.shape == (60, 1024) and bool(torch.isfinite(m.fc2.weight).all())
    assert m.decoder_type == decoder_type and (decoder_type == "rnn" or hasattr(m.decoder, "linear_out"))


def test_mbr_train_cli_one_epoch(tmp_path):
    """the MBR drop-in entry point (trainer/train_transducer_mbr_bmuf_otfaug.py): N-best generation + RNN-T and MBR branches + BMUF,
    starting from a model pickled by the RNN-T trainer, log lines with both losses, per-rank model files"""
    from test_loader_cpu import make_dataset
    from pika_b200.model.transducer import Net
    from pika_b200.trainer import train_transducer_mbr_bmuf_otfaug as T
    import types
    lst, utts = make_dataset(tmp_path, n_utts=4, shards=1, n_lo=14000, n_hi=18000)
    cfg = tmp_path / "fbank.conf"
    cfg.write_text("--window-type=hamming\n--sample-frequency=16000\n--dither=0\n--low-freq=40\n--high-freq=-200\n--num-mel-bins=80\n")
    out = tmp_path / "out"
    out.mkdir()
    torch.manual_seed(777)
    margs = types.SimpleNamespace(rnn_size=1024, local_rank=0, decoder_type="rnn", brnn=True, encoder_type="transformer", embd_dim=100,
                                  padding_idx=60, dropout=0.0, dec_layers=2, enc_layers=9)
    m0 = Net(margs, 240, 60)
    with torch.no_grad():
        m0.fc2.bias[0] += 3.0
    init = tmp_path / "init.model"
    torch.save(m0, str(init))
    log = tmp_path / "log.WORKER-ID"
    argv = ["transducer", lst, str(log), str(out), "--cuda", "--local_rank", "0", "--init_model", str(init), "--encoder_type", "transformer",
            "--decoder_type", "rnn", "--rnn_size", "1024", "--embd_dim", "100", "--output_dim", "60", "--padding_idx", "60", "--padding_tgt", "60",
            "--dec_layers", "2", "--dropout", "0.0", "--brnn", "--model_lctx", "21", "--model_rctx", "21", "--model_stride", "4",
            "--lctx", "1", "--rctx", "1", "--feats_dim", "80", "--feat_config", str(cfg), "--batch_size", "2", "--num_workers", "1", "--batch_first",
            "--max_len", "1600", "--TU_limit", "50000", "--gain_range", "25,25", "--speed_rate", "1.0", "--grad_clip", "3.0",
            "--initial_lr", "1e-4", "--final_lr", "1e-4", "--num_epochs", "1", "--num_batches_per_epoch", "2", "--sync_period", "1",
            "--beam_size", "4", "--rnnt_scale", "0.5", "--sm_scale", "0.8", "--seed", "777"]
    os.environ.setdefault("WORLD_SIZE", "1")
    T.main(argv)
    text = open(str(log).replace("WORKER-ID", "0")).read()
    assert "===> Epoch 0 <===" in text and "Overall Avg MBR Loss" in text and "Overall Avg RNNT Loss" in text and "Training Finished" in text
    vals = [float(v) for l in text.splitlines() if "Overall Avg" in l for v in [l.split("MBR Loss:")[1].split()[0], l.split("RNNT Loss:")[1].split()[0]]]
    assert np.isfinite(vals).all() and vals[1] > 0
    m = torch.load(str(out / "model.epoch.0.0"), weights_only=False)
    assert bool(torch.isfinite(m.fc2.weight).all()) and not torch.equal(m.fc2.weight.cpu(), m0.fc2.weight)
