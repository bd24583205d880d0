"""Pins the oracle's transformer prediction net (oracle/model.py:conv_transformer_lm_forward) and the matching branch of the
beam-search oracle (oracle/decode.py) to the reference: tests/golden/model_xf.npz / decode_xf.npz were produced by executing
the reference's own trainer/model/rnnt_conv_transformer_lm.py and decoder/transducer_decoder.py (make_golden.py:golden_model_xf,
golden_decode_xf).  The drop-in module creates its parameters in the reference's order, so a seeded construction reproduces
the reference's initial weights; the fingerprints check exactly that."""
import os
import types

import numpy as np
import torch


def xf_args(V):
    return types.SimpleNamespace(rnn_size=1024, local_rank=0, decoder_type="transformer", brnn=True, encoder_type="transformer",
                                 embd_dim=100, padding_idx=V, dropout=0.2, dec_layers=2, enc_layers=9)


def build_xf(V=40, seed=778):
    from pika_b200.model.transducer import Net
    torch.manual_seed(seed)
    return Net(xf_args(V), 240, V)


def xf_inputs(seed, B, Tp, H=1024):
    return np.random.default_rng(seed).standard_normal((B, Tp, H)).astype(np.float32)


def test_drop_in_xf_init_matches_reference_weights(golden_dir):
    d = np.load(os.path.join(golden_dir, "model_xf.npz"))
    sd = build_xf().state_dict()
    keys = [k[2:] for k in d.files if k.startswith("w_")]
    assert len(keys) > 40 and any(k.startswith("decoder.conv.1") for k in keys)
    for k in keys:
        v = sd[k]
        fp = np.array([v.double().sum().item(), v.double().abs().sum().item(), float(v.flatten()[0]), float(v.flatten()[-1])])
        np.testing.assert_allclose(fp, d["w_" + k], rtol=1e-12, atol=0, err_msg=k)
    assert sd["decoder.mask"].dtype == torch.uint8 and tuple(sd["decoder.mask"].shape) == (1, 5000, 5000)   # state_dict key of the reference


def test_oracle_xf_prednet_and_joint_match_reference(golden_dir):
    from oracle import model as om
    d = np.load(os.path.join(golden_dir, "model_xf.npz"))
    V, B, Tp, U = [int(v) for v in d["dims"]]
    sd = {k: v.detach() for k, v in build_xf().state_dict().items()}
    with torch.no_grad():
        pred = om.prednet_forward(sd, torch.from_numpy(d["y"]))
        logits = om.joint_forward(sd, torch.from_numpy(xf_inputs(int(d["seed"]), B, Tp)), pred, softmax=False)
    np.testing.assert_allclose(pred.numpy(), d["pred"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(logits.numpy()[:, ::3], d["logits"], rtol=0, atol=2e-5)


def test_oracle_decode_xf_matches_reference(golden_dir):
    from fixture_utils import decode_fixture_reinit_xf
    from oracle import decode as od
    torch.set_num_threads(8)
    d = np.load(os.path.join(golden_dir, "decode_xf.npz"))
    V, B, Tp = [int(v) for v in d["dims"]]
    m = build_xf(V)
    decode_fixture_reinit_xf(m)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    enc = torch.from_numpy(xf_inputs(int(d["seed"]), B, Tp))
    tl = [int(t) for t in d["tlens"]]
    for name, beam, nbest in [("b4n2", 4, 2), ("b8n4", 8, 4)]:
        ret = od.decode_batch(sd, enc, tl, beam, n_best=nbest, max_len=[t + 30 for t in tl])
        for b in range(B):
            for n in range(nbest):
                assert ret["predictions"][b][n] == d["%s_pred_%d_%d" % (name, b, n)].tolist(), (name, b, n)
                assert abs(ret["scores"][b][n] - float(d["%s_score_%d_%d" % (name, b, n)])) < 1e-4 * abs(ret["scores"][b][n]) + 1e-4
