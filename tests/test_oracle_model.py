"""Pins oracle/model.py against the reference's own nn.Modules (tests/golden/model_small.npz,
encoder_eval_T200.npz were produced by executing /root/reference -- see make_golden.py).

The drop-in model (pika_b200.model.transducer.Net) reproduces the reference's parameter
creation order, so ``torch.manual_seed(777); Net(...)`` yields bit-identical weights; the weight
fingerprints in the golden file check exactly that."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import model as om


def build_state_dict(V=40):
    from pika_b200.model.transducer import Net
    torch.manual_seed(777)
    args = types.SimpleNamespace(rnn_size=1024, local_rank=0, decoder_type="rnn", brnn=True,
                                 encoder_type="transformer", embd_dim=100, padding_idx=V,
                                 dropout=0.2, dec_layers=2, enc_layers=9)
    net = Net(args, 240, V)
    return net, {k: v.detach().clone() for k, v in net.state_dict().items()}


@pytest.fixture(scope="module")
def net_sd():
    return build_state_dict()


def test_drop_in_init_matches_reference_weights(golden_dir, net_sd):
    d = np.load(os.path.join(golden_dir, "model_small.npz"))
    _, sd = net_sd
    keys = [k[2:] for k in d.files if k.startswith("w_")]
    assert len(keys) > 100
    for k in keys:
        v = sd[k]
        fp = np.array([v.double().sum().item(), v.double().abs().sum().item(),
                       float(v.flatten()[0]), float(v.flatten()[-1])])
        np.testing.assert_allclose(fp, d["w_" + k], rtol=1e-12, atol=0, err_msg=k)


def test_oracle_encoder_eval_matches_reference(golden_dir, net_sd):
    d = np.load(os.path.join(golden_dir, "encoder_eval_T200.npz"))
    _, sd = net_sd
    with torch.no_grad():
        enc = om.encoder_forward(sd, torch.from_numpy(d["x"]), train=False)
    np.testing.assert_allclose(enc.numpy(), d["enc"], rtol=0, atol=2e-5)


def test_oracle_train_forward_matches_reference(golden_dir, net_sd):
    d = np.load(os.path.join(golden_dir, "model_small.npz"))
    _, sd = net_sd
    taps = {}
    with torch.no_grad():
        enc = om.encoder_forward(sd, torch.from_numpy(d["x"]), train=True, taps=taps)
        pred = om.prednet_forward(sd, torch.from_numpy(d["y"]).long())
        logits = om.joint_forward(sd, enc, pred, softmax=False)
    for k in d.files:
        if k.startswith("tap_"):
            np.testing.assert_allclose(taps[k[4:]][:, ::7, ::13].numpy(), d[k], atol=2e-4, err_msg=k)   # fp32 re-association across 12 layers
    np.testing.assert_allclose(enc.numpy(), d["enc"], atol=2e-4)
    np.testing.assert_allclose(pred.numpy(), d["pred"], atol=1e-6)
    np.testing.assert_allclose(logits.numpy(), d["logits"], atol=1e-5)
    from oracle import rnnt
    costs, _ = rnnt.rnnt_loss_from_logits(logits.numpy(), d["y"], d["tlens"], d["ulens"])
    np.testing.assert_allclose(costs, d["costs"], rtol=1e-5)
    tl = om.frame_lens_after_encoder(torch.from_numpy(d["lens"]))
    assert tl.tolist() == d["tlens"].tolist()
