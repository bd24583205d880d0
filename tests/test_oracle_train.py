"""Pins oracle/train.py against what the reference's training loop literally calls
(trainer/train_transducer_bmuf_otfaug.py:46-55, 105-123): ``torch.nn.utils.clip_grad_norm_(..., norm_type=inf)``,
``torch.optim.SGD(momentum, nesterov=True)`` re-created after every BMUF sync, and the exponential learning-rate formula;
the BMUF update formula against the reference BmufTrainer's committed 2-rank trajectory (tests/golden/bmuf_2rank.npz)."""
import math
import os

import numpy as np
import torch
from math import inf

from oracle import train as ot


def test_clip_and_nesterov_sgd_match_torch_over_optimizer_recreation():
    torch.manual_seed(3)
    params = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11))]
    p_np = [p.detach().numpy().astype(np.float64).copy() for p in params]
    bufs = [None, None]
    lr, mom, clip = 3e-2, 0.9, 0.5
    opt = torch.optim.SGD(params, lr, momentum=mom, nesterov=True)
    for step in range(7):
        if step == 4:                                   # the reference builds a fresh optimiser after each sync (:121-123)
            lr = 1e-2
            opt = torch.optim.SGD(params, lr, momentum=mom, nesterov=True)
            bufs = [None, None]
        g = torch.Generator().manual_seed(50 + step)
        grads = [torch.randn(p.shape, generator=g) * (3.0 if step % 2 else 0.1) for p in params]   # clipped and unclipped steps
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(params, clip, norm_type=inf)
        opt.step()
        t_o, coef = ot.clip_coef_inf([gr.numpy() for gr in grads], clip)
        assert abs(t_o - float(total)) < 1e-6
        for i in range(2):
            gq = grads[i].numpy().astype(np.float64) * coef
            p_np[i], bufs[i] = ot.sgd_nesterov_step(p_np[i], gq, bufs[i], lr, mom, first=bufs[i] is None)
            np.testing.assert_allclose(p_np[i], params[i].detach().numpy(), rtol=0, atol=2e-6)


def test_lr_schedule_is_the_reference_expression():
    for n, tot in [(0, 100), (37, 100), (100, 100), (5, 7)]:
        want = 4e-4 * math.exp(n * math.log(2e-5 / 4e-4) / tot)        # trainer/train_transducer_bmuf_otfaug.py:115-118
        assert ot.lr_schedule(4e-4, 2e-5, n, tot) == want
    assert abs(ot.lr_schedule(4e-4, 2e-5, 100, 100) - 2e-5) < 1e-12


def test_bmuf_formula_reproduces_reference_trainer_trajectory(golden_dir):
    gold = np.load(os.path.join(golden_dir, "bmuf_2rank.npz"))["params"].astype(np.float64)
    glob, dprev = gold[0].copy(), np.zeros_like(gold[0])
    # rank r's model before sync `it`: the broadcast parameters + its own seeded perturbation (make_golden.py:_bmuf_ref_worker)
    for it in range(3):
        locals_ = []
        for rank in range(2):
            g = torch.Generator().manual_seed(7 * it + rank)
            noise = (0.01 * torch.randn(gold.shape[1], generator=g)).numpy()
            locals_.append((glob.astype(np.float32) + noise).astype(np.float64))
        glob, dprev = ot.bmuf_update(glob, locals_, dprev, 0.9, 1.0)
        np.testing.assert_allclose(glob, gold[it + 1], rtol=0, atol=3e-7)
