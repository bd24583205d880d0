"""End-to-end GPU parity of the drop-in transducer against the fixtures produced by executing the
reference's own modules (tests/golden/model_small.npz, encoder_eval_T200.npz).

fp32-class mode: 1e-3 norm-relative on encoder activations / joint logits / loss (north star).
bf16 production mode: bf16 rounding accumulates through 12 GEMM layers, so activations are held to
6e-2 norm-relative and the loss to 1e-2 relative (documented in DESIGN.md)."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _record(name, obj):
    """measured parity figures land in gpurun_out/parity_measured.jsonl so that the stated tolerances can be checked against them"""
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_measured.jsonl"), "a") as f:
        f.write(json.dumps({"name": name, **obj}) + "\n")


def rel(a, b):
    a = torch.as_tensor(a).float().cpu(); b = torch.as_tensor(b).float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def build(V=40):
    from pika_b200.model.transducer import Net
    torch.manual_seed(777)
    args = types.SimpleNamespace(rnn_size=1024, local_rank=0, decoder_type="rnn", brnn=True, encoder_type="transformer",
                                 embd_dim=100, padding_idx=V, dropout=0.2, dec_layers=2, enc_layers=9)
    return Net(args, 240, V).cuda()


@pytest.mark.parametrize("precision,tol_act,tol_loss", [("fp32", 1e-3, 1e-3), ("bf16", 6e-2, 1e-3)])   # bf16 measured: activations 3.9e-2, loss 9e-5
def test_train_forward_backward_matches_reference(golden_dir, precision, tol_act, tol_loss):
    from pika_b200 import engine
    d = np.load(os.path.join(golden_dir, "model_small.npz"))
    engine.set_precision(precision)
    engine.set_dropout_enabled(False)
    try:
        m = build()
        m.train()
        x = torch.from_numpy(d["x"]).cuda()
        y = torch.from_numpy(d["y"]).long().cuda()
        enc = engine.encoder_forward(m.encoder, x)
        assert rel(enc, d["enc"]) < tol_act
        meas = {"enc": rel(enc, d["enc"])}
        for k in ("encoder.bn_in.running_mean", "encoder.hidden_bn.8.running_var", "encoder.bn_final.running_mean"):
            v = dict(m.named_buffers())[k]
            assert abs(v.double().sum().item() - d["bn_" + k][0]) < max(1e-3 * abs(d["bn_" + k][0]), 3e-2 if precision == "bf16" else 2e-3)
        pred = engine.prednet_forward_act(m, y)
        assert rel(pred, d["pred"]) < tol_act
        meas["pred"] = rel(pred, d["pred"])
        m2 = build(); m2.train()
        logits = m2.forward(x, y, None, softmax=False)
        assert logits.dtype == torch.float32 and tuple(logits.shape) == d["logits"].shape
        assert rel(logits, d["logits"]) < tol_act
        meas["logits"] = rel(logits, d["logits"])
        lp = m2.forward(x, y, None, softmax=True)
        assert rel(lp, torch.log_softmax(torch.from_numpy(d["logits"]), -1)) < tol_act
        # fused training path: loss + every parameter gradient
        m3 = build(); m3.train()
        costs = engine.transducer_loss(m3, x, y, torch.from_numpy(d["tlens"]).cuda(), torch.from_numpy(d["ulens"]).cuda())
        np.testing.assert_allclose(costs.detach().cpu().numpy(), d["costs"], rtol=tol_loss)
        meas["loss"] = float(np.abs(costs.detach().cpu().numpy() / d["costs"] - 1).max())
        _record("model_small_%s" % precision, meas)
        costs.sum().backward()
        gtol = 5e-3 if precision == "fp32" else 0.15
        bad = []
        for k, p in m3.named_parameters():
            ref = d["g_" + k]
            assert p.grad is not None, k
            nrm = p.grad.double().norm().item()
            # biases feeding a BatchNorm have an analytically zero gradient: compare on an absolute floor
            if abs(nrm - ref[0]) > gtol * max(ref[0], 1e-6) + (1e-5 if precision == "fp32" else 1e-3):
                bad.append((k, nrm, ref[0]))
        assert not bad, bad
        # full gradient DIRECTION, not just the norm: <= 512 strided samples of every parameter gradient (make_golden.py:golden_model)
        from fixture_utils import grad_fingerprint
        worst = {}
        for k, p in m3.named_parameters():
            ref = d["gs_" + k]
            got = grad_fingerprint(p.grad.cpu(), 512)
            rn = np.linalg.norm(ref[3:])
            if ref[2] < 1e-6:                          # analytically zero gradient (a bias in front of a BatchNorm)
                assert np.abs(got[3:]).max() < (1e-5 if precision == "fp32" else 2e-3), k
                continue
            if rn < 1e-3 * ref[2]:                     # the strided sample happens to hold none of the gradient's mass
                continue
            err = np.linalg.norm(got[3:] - ref[3:]) / rn
            cos = float(np.dot(got[3:], ref[3:]) / (np.linalg.norm(got[3:]) * rn + 1e-30))
            worst[k] = (err, cos)
        w_err = max(v[0] for v in worst.values()); w_cos = min(v[1] for v in worst.values())
        errs = sorted(v[0] for v in worst.values())
        top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:6]
        _record("model_small_grad_samples_%s" % precision, dict(worst_rel_err=w_err, worst_cos=w_cos, median_rel_err=errs[len(errs) // 2],
                                                                 p90_rel_err=errs[int(0.9 * len(errs))], n=len(errs),
                                                                 worst=[(k, round(v[0], 4), round(v[1], 5)) for k, v in top]))
        _record("model_small_grad_profile_%s" % precision, dict(per_param={k: round(v[0], 5) for k, v in worst.items()}))
        # measured (gpurun_out/parity_measured.jsonl): fp32-class mode median 0.8 %, worst 1.5 %, cosine >= 0.99989 -- an error ORTHOGONAL to
        # the gradient, of the same relative size on every encoder parameter; bf16 median 24 %, worst 34 %, cosine >= 0.94, again uniform.
        # It scales with the unit round-off of the arithmetic (1e-5 -> 4e-3), i.e. it is the conditioning of this randomly initialised
        # 12-layer batch-statistics / softmax stack (forward activations differ by 8e-5 and 4e-2), not a modelling difference: a wrong
        # mask, sign or transposition would show up at O(1) in the fp32-class mode, which is what the tight bound below guards.
        assert w_err < (3e-2 if precision == "fp32" else 0.5), top
        assert w_cos > (0.9995 if precision == "fp32" else 0.9), top
        assert errs[len(errs) // 2] < (1.5e-2 if precision == "fp32" else 0.35), errs[len(errs) // 2]
        # unfused compatibility path (model.forward + RNNTLoss.apply) gives the same loss
        from pika_b200.warp_rnnt import RNNTLoss
        m4 = build(); m4.train()
        lp4 = m4.forward(x, y, None, True)
        loss4 = RNNTLoss(blank=0, reduction="sum").apply(lp4, y.int(), torch.from_numpy(d["tlens"]).cuda(),
                                                          torch.from_numpy(d["ulens"]).cuda())
        np.testing.assert_allclose(loss4.detach().cpu().numpy(), d["costs"], rtol=tol_loss)
        loss4.sum().backward()
        g3 = m3.fc2.weight.grad; g4 = m4.fc2.weight.grad
        assert rel(g4, g3) < (2e-3 if precision == "fp32" else 5e-2)
    finally:
        engine.set_precision("bf16")
        engine.set_dropout_enabled(True)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
def test_encoder_eval_config1(golden_dir, precision, tol):
    """BASELINE config 1: encoder forward, 1 utterance, T=200, eval mode."""
    from pika_b200 import engine
    d = np.load(os.path.join(golden_dir, "encoder_eval_T200.npz"))
    engine.set_precision(precision)
    try:
        m = build(); m.eval()
        with torch.no_grad():
            enc = m.encoder(torch.from_numpy(d["x"]).cuda())
        assert tuple(enc.shape) == d["enc"].shape
        assert rel(enc, d["enc"]) < tol
    finally:
        engine.set_precision("bf16")


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_full_shape_T1000_U150_V6000_matches_reference(golden_dir, precision):
    """BASELINE shape (T=1000 frames -> T'=240, U=150, V=6000), B=2 with ragged lengths, against the reference's own modules
    run on the CPU (tests/golden/model_full_shape.npz, make_golden.py:golden_model_full).  Exercises what the small fixtures
    cannot: the 24-tile row-LSE epilogue of the fc2 GEMM inside the model, the 151-wide lattice, 6000-column gradients."""
    from pika_b200 import engine
    from fixture_utils import grad_fingerprint
    from make_inputs import full_shape_inputs
    d = np.load(os.path.join(golden_dir, "model_full_shape.npz"))
    V = int(d["V"])
    x_np, y_np, lens_np, ulens_np = full_shape_inputs(seed=int(d["seed"]), V=V)
    engine.set_precision(precision)
    engine.set_dropout_enabled(False)
    try:
        m = build(V)
        m.train()
        x, y = torch.from_numpy(x_np).cuda(), torch.from_numpy(y_np).cuda()
        tl, ul = torch.from_numpy(d["tlens"]).cuda(), torch.from_numpy(ulens_np).cuda()
        enc = engine.encoder_forward_act(m.encoder, x)
        pred = engine.prednet_forward_act(m, y)
        logits = engine.JointFn.apply(enc, pred, m)[..., :V]
        meas = dict(enc=rel(enc[:, ::5, ::7], d["enc"]), pred=rel(pred[:, ::3, ::7], d["pred"]),
                    logits=rel(logits[:, ::9, ::7, ::53], d["logits"]),
                    lse=rel(torch.logsumexp(logits[:, ::9, ::7].float(), -1), d["lse"]))
        del logits, enc, pred
        m3 = build(V); m3.train()
        costs = engine.transducer_loss(m3, x, y, tl, ul)
        got = costs.detach().cpu().numpy()
        meas["loss"] = float(np.abs(got / d["costs"] - 1).max())
        costs.sum().backward()
        worst = {}
        gmax = max(float(d["gs_" + k][2]) for k, _ in m3.named_parameters())
        for k, p in m3.named_parameters():
            ref = d["gs_" + k]
            g = grad_fingerprint(p.grad.cpu(), 256)
            rn = np.linalg.norm(ref[3:])
            # a bias in front of a BatchNorm has an analytically zero gradient (the reference holds rounding noise there): skipped,
            # like parameters whose strided sample misses the gradient's mass
            if ref[2] < 1e-5 * gmax or rn < 1e-3 * ref[2]:
                assert float(np.abs(g[3:]).max()) < 1e-3 * gmax, k
                continue
            worst[k] = float(np.linalg.norm(g[3:] - ref[3:]) / rn)
        errs = sorted(worst.values())
        meas["grad_worst"] = max(worst.values())
        meas["grad_worst_key"] = max(worst, key=worst.get)
        meas["grad_median"] = errs[len(errs) // 2]
        meas["grad_top"] = [(k, round(v, 4)) for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:6]]
        meas["grad_low"] = [(k, round(v, 5)) for k, v in sorted(worst.items(), key=lambda kv: kv[1])[:8]]
        meas["grad_fc2_bias"] = worst.get("fc2.bias")
        # the fc2 bias gradient IS the column sum of dlogits over all 72 480 rows: a 6000-wide signature of the fused loss gradient
        cs = m3.fc2.bias.grad.double().cpu().numpy()
        meas["dlogits_colsum"] = float(np.linalg.norm(cs - d["dlogits_colsum"]) / np.linalg.norm(d["dlogits_colsum"]))
        _record("model_full_shape_%s" % precision, meas)
        act_tol, loss_tol, grad_tol = (1e-3, 1e-3, 5e-2) if precision == "fp32" else (6e-2, 1e-3, 0.5)     # bf16 measured: 3.6e-2, 2e-6, 0.33
        assert meas["enc"] < act_tol and meas["pred"] < act_tol and meas["logits"] < act_tol, meas
        assert meas["lse"] < (1e-4 if precision == "fp32" else 2e-3), meas
        assert meas["loss"] < loss_tol, meas
        assert meas["dlogits_colsum"] < (2e-3 if precision == "fp32" else 3e-2), meas
        assert meas["grad_worst"] < grad_tol, sorted(worst.items(), key=lambda kv: -kv[1])[:5]
    finally:
        engine.set_precision("bf16")
        engine.set_dropout_enabled(True)
