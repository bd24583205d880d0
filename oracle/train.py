"""Oracle: gradient clipping, Nesterov SGD and the BMUF block update, numpy/torch-CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Paths relative to /root/reference.
"""
import math
import numpy as np


def lr_schedule(initial_lr, final_lr, num_batches_processed, total_num_batches):
    """trainer/train_transducer_bmuf_otfaug.py:46-51, 115-120: exponential decay."""
    return initial_lr * math.exp(num_batches_processed * math.log(final_lr / initial_lr) / total_num_batches)


def clip_coef_inf(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type=inf)
    (trainer/train_transducer_bmuf_otfaug.py:106-109): total = max |g|; coef = max_norm/(total+1e-6),
    applied only when < 1."""
    total = max(float(np.abs(g).max()) for g in grads)
    coef = max_norm / (total + 1e-6)
    return total, min(coef, 1.0)


def sgd_nesterov_step(p, g, buf, lr, momentum, first):
    """torch.optim.SGD(momentum, nesterov=True), dampening 0, weight_decay 0
    (trainer/train_transducer_bmuf_otfaug.py:53-55, 110).  ``first`` = momentum buffer not yet
    created (it is re-created after every BMUF sync, :121-123): buf := g.
    Returns (p_new, buf_new)."""
    buf = g.copy() if first else momentum * buf + g
    d = g + momentum * buf
    return p - lr * d, buf


def bmuf_update(param_global, local_params_per_rank, delta_prev, block_momentum, block_lr):
    """BmufTrainer.update_and_sync (trainer/bmuf.py:76-100):
        delta = sum_r (param_global - local_r) / N
        delta_prev = bm * delta_prev + block_lr * (1 - bm) * delta
        param_global -= (1 + bm) * delta_prev
    Returns (param_global_new, delta_prev_new)."""
    n = len(local_params_per_rank)
    delta = sum((param_global - l) for l in local_params_per_rank) / float(n)
    delta_prev = block_momentum * delta_prev + block_lr * (1 - block_momentum) * delta
    return param_global - (1 + block_momentum) * delta_prev, delta_prev
