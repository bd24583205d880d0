"""Pins oracle/mbr.py (the CPU restatement every MBR parity test compares the CUDA path with) against the REFERENCE's own
MBR loop body, which tests/golden/make_golden.py executed on CPU from the reference source text
(trainer/train_transducer_mbr_bmuf_otfaug.py, '#nbest genereation' .. 'out.backward(mbr_grad)'): same weights, batch and
reference-decoded N-best list -> same MBR loss, RNN-T loss and every parameter gradient (fingerprints: sum, |sum|, norm and
384 strided samples per tensor)."""
import os
import types

import numpy as np
import torch


def test_mbr_oracle_matches_reference_loop_body(golden_dir):
    from fixture_utils import decode_fixture_reinit, grad_fingerprint
    from oracle import mbr as ombr
    from pika_b200.model.transducer import Net
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    d = np.load(os.path.join(golden_dir, "mbr_small.npz"))
    dd = np.load(os.path.join(golden_dir, "decode_small.npz"))
    V = 40
    torch.manual_seed(777)
    margs = types.SimpleNamespace(rnn_size=1024, local_rank=0, decoder_type="rnn", brnn=True, encoder_type="transformer",
                                  embd_dim=100, padding_idx=V, dropout=0.2, dec_layers=2, enc_layers=9)
    m = Net(margs, 240, V)                                  # parameter container: reference init order and RNG stream
    decode_fixture_reinit(m)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    names = [k for k, _ in m.named_parameters()]
    for k in names:
        sd[k].requires_grad_(True)
    x = torch.from_numpy(dd["x"])
    tl = dd["tlens"].astype(np.int32)
    target, ul = torch.from_numpy(d["target"]), d["ulens"].astype(np.int32)
    hyps = [[[int(t) for t in h if t != -2] for h in row] for row in d["hyps"]]
    scores = [[float(s) for s in row] for row in d["scores"]]
    rnnt_scale, sm_scale = float(d["rnnt_scale"]), float(d["sm_scale"])
    mbr_loss, costs = ombr.mbr_loss_and_grads(sd, x, target, tl, ul, hyps, scores, 0, V, rnnt_scale, sm_scale)
    assert abs(mbr_loss - float(d["mbr_loss"])) < 1e-5 * max(1.0, abs(float(d["mbr_loss"])))
    assert abs(rnnt_scale * float(np.sum(costs)) - float(d["rnnt_loss"])) < 1e-4 * abs(float(d["rnnt_loss"]))
    checked = 0
    for k in names:
        ref = d["g:" + k]
        got = grad_fingerprint(sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k]))
        if ref[2] < 1e-4:                                   # analytically zero gradients (bias feeding a BatchNorm, key bias under the
            assert got[2] < 1e-3, k                         # softmax's shift invariance): rounding noise on both sides
            continue
        assert abs(got[2] - ref[2]) < 5e-4 * ref[2], (k, got[2], ref[2])
        # two fp32 CPU evaluations of a 40-GEMM-deep, deliberately high-gain network (Conv2d vs three matmul taps, different
        # reduction orders): measured <= 2.5e-3 sample-relative in the deepest TDNN layers, <= 1e-4 near the output
        srel = np.linalg.norm(got[3:] - ref[3:]) / np.linalg.norm(ref[3:])
        assert srel < 6e-3, (k, srel)
        checked += 1
    assert checked > 90


# ------------------------------------------------------------------------------------------------ host side of the MBR batch (CPU)
def test_edit_distance_batch_matches_scalar_and_oracle():
    from oracle import mbr as ombr
    from pika_b200.trainer.mbr import edit_distance, edit_distance_batch
    rng = np.random.default_rng(1)
    refs, hyps = [], []
    for _ in range(200):
        la, lb = int(rng.integers(0, 14)), int(rng.integers(0, 14))
        refs.append(rng.integers(1, 5, la).tolist())
        hyps.append(rng.integers(1, 5, lb).tolist())
    refs += [[], [1, 2, 3], []]
    hyps += [[], [], [4, 4]]
    got = edit_distance_batch(refs, hyps)
    for r, h, g in zip(refs, hyps, got):
        assert g == edit_distance(r, h) == ombr.levenshtein(r, h), (r, h, g)
    assert edit_distance_batch([], []).shape == (0,)
    assert edit_distance_batch([[], []], [[], []]).tolist() == [0, 0]


def test_nbest_risk_and_alignment_nodes_match_the_reference_loops():
    """pika_b200.trainer.mbr.nbest_risk / alignment_nodes (vectorised) against the per-token Python loops they replace
    (trainer/train_transducer_mbr_bmuf_otfaug.py:171-232), for hypotheses given as int lists, arrays and lists of 0-d tensors"""
    import torch
    from oracle import mbr as ombr
    from pika_b200.trainer.mbr import alignment_nodes, nbest_risk
    rng = np.random.default_rng(2)
    bsz, beam, blk, Tp, V = 3, 4, 0, 9, 12
    hyps = [[rng.integers(0, 4, int(rng.integers(Tp, Tp + 8))).tolist() for _ in range(beam)] for _ in range(bsz)]
    hyps[1][2] = [0] * Tp                                                       # an all-blank alignment
    scores = [[float(v) for v in rng.standard_normal(beam)] for _ in range(bsz)]
    ulens = np.array([5, 2, 4])
    targets = rng.integers(1, V, (bsz, 5))
    forms = [hyps, [[np.asarray(h) for h in row] for row in hyps], [[[torch.tensor(t) for t in h] for h in row] for row in hyps]]
    for hh in forms:
        nonblk, prob, dist, seq_grad, loss = nbest_risk(hh, scores, targets, ulens, blk)
        sc = torch.tensor(scores, dtype=torch.float32)
        p_ref = torch.softmax(sc, dim=1).numpy()
        np.testing.assert_allclose(prob, p_ref, rtol=1e-6)
        for i in range(bsz):
            for j in range(beam):
                nb = [t for t in hyps[i][j] if t != blk]
                assert nonblk[i][j] == nb
                assert dist[i, j] == ombr.levenshtein(targets[i][:ulens[i]].tolist(), nb)
        avg = (p_ref * dist).sum(1, keepdims=True)
        np.testing.assert_allclose(seq_grad, p_ref * (dist - avg), rtol=1e-5, atol=1e-7)
        assert abs(loss - float(avg.sum())) < 1e-5
        U = max(max(len(h) for row in nonblk for h in row), 5)
        ex, py, tk, cf = alignment_nodes(hh, seq_grad, Tp, U, blk)
        ex_r, py_r, tk_r, cf_r = [], [], [], []
        for i in range(bsz):
            for j in range(beam):
                t_i = u_i = 0
                sg = float(seq_grad[i, j])
                for t in hyps[i][j]:
                    ex_r.append(i * Tp + min(t_i, Tp - 1)); py_r.append((i * beam + j) * (U + 1) + min(u_i, U)); tk_r.append(t)
                    cf_r.append(sg / float(Tp) if t == blk else sg)
                    if t == blk:
                        t_i += 1
                    else:
                        u_i += 1
        assert ex.dtype == np.int32 and py.dtype == np.int32 and tk.dtype == np.int32 and cf.dtype == np.float32
        assert ex.tolist() == ex_r and py.tolist() == py_r and tk.tolist() == tk_r
        np.testing.assert_array_equal(cf, np.asarray(cf_r, np.float32))
