"""Transformer prediction net (decoder_type='transformer') on the GPU against fixtures produced by executing the reference's own
modules (tests/golden/model_xf.npz, decode_xf.npz; make_golden.py:golden_model_xf / golden_decode_xf):
trainer/model/rnnt_conv_transformer_lm.py:59-80 (training forward / backward through the joint and the loss) and
decoder/transducer_decoder.py:117-120,151-171 (beam search that re-runs the hypothesis history).

fp32-class mode: 1e-3 norm-relative on the prediction-net output, loss and every gradient; bf16: stated per assert."""
import os
import types

import numpy as np
import pytest
import torch

from test_oracle_xf_prednet import build_xf, xf_inputs

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).float().cpu(); b = torch.as_tensor(b).float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("precision,tol_act,tol_grad", [("fp32", 1e-3, 1e-3), ("bf16", 3e-2, 0.12)])   # measured: fp32 1.5e-5 / 3e-5, bf16 8e-3 / 8.5e-2
def test_xf_prednet_train_matches_reference(golden_dir, precision, tol_act, tol_grad):
    from fixture_utils import grad_fingerprint
    from pika_b200 import engine
    from test_model_gpu import _record
    d = np.load(os.path.join(golden_dir, "model_xf.npz"))
    V, B, Tp, U = [int(v) for v in d["dims"]]
    engine.set_precision(precision)
    engine.set_dropout_enabled(False)
    try:
        m = build_xf(V).cuda()
        m.train()
        y = torch.from_numpy(d["y"]).cuda()
        pred = engine.prednet_forward_act(m, y)
        assert tuple(pred.shape) == d["pred"].shape
        e_pred = rel(pred, d["pred"])
        assert e_pred < tol_act
        # module-level forward of the prediction net (drop-in for Net.forward of rnnt_conv_transformer_lm.py)
        sos = torch.zeros(B, 1, dtype=torch.long, device="cuda")
        assert rel(m.decoder(torch.cat((sos, y), 1)), d["pred"]) < tol_act
        enc = torch.from_numpy(xf_inputs(int(d["seed"]), B, Tp)).cuda().requires_grad_(precision == "fp32")
        costs = engine.JointLossFn.apply(engine._to_act(enc), pred, m, y.int().contiguous(), torch.from_numpy(d["tlens"]).cuda(),
                                         torch.from_numpy(d["ulens"]).cuda(), True)
        np.testing.assert_allclose(costs.detach().cpu().numpy(), d["costs"], rtol=1e-3)
        costs.sum().backward()
        worst = {}
        for k, p in m.named_parameters():
            if k.startswith("encoder."):
                continue
            ref = d["gs_" + k]
            assert p.grad is not None, k
            got = grad_fingerprint(p.grad.cpu(), 512)
            rn = np.linalg.norm(ref[3:])
            if ref[2] < 1e-7 or rn < 1e-3 * ref[2]:
                continue
            worst[k] = float(np.linalg.norm(got[3:] - ref[3:]) / rn)
            assert abs(got[2] - ref[2]) < tol_grad * ref[2], (k, got[2], ref[2])                # gradient norm
        top = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
        _record("model_xf_%s" % precision, dict(pred=e_pred, loss=float(np.abs(costs.detach().cpu().numpy() / d["costs"] - 1).max()),
                                                grad_worst=top[0][1], grad_top=[(k, round(v, 5)) for k, v in top], n=len(worst)))
        assert len(worst) > 40 and top[0][1] < tol_grad, top
        # the padding row of the shared embedding table receives no gradient (nn.Embedding(padding_idx), trainer/model/transducer.py:52-53)
        assert float(m.embed.weight.grad[V].abs().max()) == 0.0
        if precision == "fp32":
            ref = d["denc"]
            got = grad_fingerprint(enc.grad.cpu(), 512)
            assert np.linalg.norm(got[3:] - ref[3:]) / np.linalg.norm(ref[3:]) < tol_grad
    finally:
        engine.set_precision("bf16")
        engine.set_dropout_enabled(True)


def test_xf_prednet_dropout_path_runs():
    """training mode with dropout on (attention / residual / FFN dropout of the transformer layers): finite loss and gradients"""
    from pika_b200 import engine
    V, B, Tp, U = 40, 2, 12, 7
    m = build_xf(V).cuda()
    m.train()
    g = torch.Generator().manual_seed(3)
    y = torch.randint(1, V, (B, U), generator=g).cuda()
    enc = torch.from_numpy(xf_inputs(11, B, Tp)).cuda()
    pred = engine.prednet_forward_act(m, y)
    costs = engine.JointLossFn.apply(engine._to_act(enc), pred, m, y.int().contiguous(), torch.full((B,), Tp, dtype=torch.int32, device="cuda"),
                                     torch.full((B,), U, dtype=torch.int32, device="cuda"), True)
    costs.sum().backward()
    assert bool(torch.isfinite(costs).all())
    for k, p in m.named_parameters():
        if k.startswith("decoder."):
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k


@pytest.mark.parametrize("name,beam,nbest", [("b4n2", 4, 2), ("b8n4", 8, 4)])
def test_xf_decode_matches_reference(golden_dir, name, beam, nbest):
    """token ids bit-exact against the reference's decode_batch for every hypothesis whose reference score is separated from its
    n-best neighbours by more than 5e-3 (same rule as the beam-16 test); every score within 1e-3.  fp32-class mode."""
    from fixture_utils import decode_fixture_reinit_xf
    from pika_b200 import engine
    from pika_b200.decoder.beam_transducer import GlobalScorer
    from pika_b200.decoder.transducer_decoder import TransducerDecoder
    d = np.load(os.path.join(golden_dir, "decode_xf.npz"))
    V, B, Tp = [int(v) for v in d["dims"]]
    engine.set_precision("fp32")
    try:
        m = build_xf(V)
        decode_fixture_reinit_xf(m)
        m = m.cuda().eval()
        dargs = types.SimpleNamespace(las_rescorer=None, las_rescorer_bw=None, bilas_rescorer=None, nonblk_reward=0.0)
        dec = TransducerDecoder(m, B, beam, n_best=nbest, blk=0, global_scorer=GlobalScorer(), sm_scale=1.0, cuda=True, beam_prune=True,
                                args=dargs)
        enc = torch.from_numpy(xf_inputs(int(d["seed"]), B, Tp)).cuda()
        tl = torch.from_numpy(d["tlens"])
        ret, _ = dec.decode_batch(None, tl, max_len=[int(t) + 30 for t in tl], enc_out=enc)
    finally:
        engine.set_precision("bf16")
    exact = 0
    for b in range(B):
        ref_scores = [float(d["%s_score_%d_%d" % (name, b, n)]) for n in range(nbest)]
        for n in range(nbest):
            sc = float(ret["scores"][b][n])
            assert abs(sc - ref_scores[n]) < 1e-3 * abs(sc) + 1e-3, (b, n, sc, ref_scores[n])
            gap = min([abs(ref_scores[n] - ref_scores[j]) for j in (n - 1, n + 1) if 0 <= j < nbest])
            hyp = [int(t.item()) for t in ret["predictions"][b][n]]
            ref = d["%s_pred_%d_%d" % (name, b, n)].tolist()
            if gap > 5e-3 or n == 0:
                assert hyp == ref, (b, n, gap, hyp[:30], ref[:30])
                exact += 1
    assert exact >= B * nbest - 3
