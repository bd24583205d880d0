"""Flat fp32 parameter / gradient storage and the fused clip + Nesterov-SGD step.

All parameters of the model become views into ONE contiguous fp32 buffer (and their ``.grad`` views
into a second one), so that gradient clipping (inf-norm), the optimiser and the BMUF exchange are
single HBM streams over 91 M floats instead of ~200 small kernels plus
``parameters_to_vector`` / ``_copy_vec_to_param`` copies (trainer/bmuf.py:14-35,62-63,83-84,98).
"""
import math

import torch

from .. import engine
from .. import kernels as K


class FlatParams:
    def __init__(self, model):
        params = [p for p in model.parameters()]
        assert all(p.dtype == torch.float32 for p in params), "model parameters must be fp32"
        # 16-byte aligned slots so that every parameter can be a TMA source / destination
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        dev = params[0].device
        self.data = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, offs):
                self.data[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.data[o:o + p.numel()].view(p.shape)
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        self.params, self.offsets, self.numel = params, offs, n
        self.num_param = sum(p.numel() for p in params)
        engine.invalidate_weights()

    def zero_grad(self):
        self.grad.zero_()


def lr_at(initial_lr, final_lr, num_batches_processed, total_num_batches):
    """exponential decay (trainer/train_transducer_bmuf_otfaug.py:46-51,115-120)"""
    return initial_lr * math.exp(num_batches_processed * math.log(final_lr / initial_lr) / total_num_batches)


class SgdNesterovClip:
    """clip_grad_norm_(params, max_norm, norm_type=inf) + optim.SGD(lr, momentum, nesterov=True).step()
    (trainer/train_transducer_bmuf_otfaug.py:53-55,105-110) as two kernels over the flat buffers.
    ``reset()`` drops the momentum buffer, as re-creating the optimiser after every BMUF sync does in the
    reference (:121-123)."""

    def __init__(self, flat, lr, momentum=0.9, max_norm=-1.0):
        self.flat, self.lr, self.momentum, self.max_norm = flat, lr, momentum, max_norm
        self.buf = torch.zeros_like(flat.data)
        self.first = True
        self.absmax = torch.zeros(1, dtype=torch.float32, device=flat.data.device)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=flat.data.device)

    def reset(self, lr=None):
        self.first = True
        if lr is not None:
            self.lr = lr

    def step(self):
        f = self.flat
        if self.max_norm > 0:
            self.nan_flag.zero_()
            K.absmax(f.grad, self.absmax, self.nan_flag)
        K.sgd_nesterov_clip(f.data, f.grad, self.buf, self.lr, self.momentum, self.max_norm,
                            self.absmax if self.max_norm > 0 else None, self.first,
                            nan_flag=self.nan_flag if self.max_norm > 0 else None)
        self.first = False
        engine.invalidate_weights()
