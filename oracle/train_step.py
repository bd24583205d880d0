"""Oracle / CPU baseline: one full RNN-T training batch on the host CPU, restating
trainer/train_transducer_bmuf_otfaug.py:71-123 (front end of loader/otf_utt_loader.py:218-250 included).

TEST INFRASTRUCTURE / REPORTED BASELINE ONLY (see oracle/__init__.py).  This is the "port" kind of
``cpu_baseline`` in bench.py: the reference's own Python cannot travel to the GPU box
(/root/reference does not exist there) and depends on PyKaldi / warp_rnnt, so the timed CPU arm is
this restatement -- torch-CPU fp32 primitives for the dense layers (all host threads), the C lattice
DP (oracle/rnnt_c.c) for the loss, numpy for the front end.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

from . import frontend as fe
from . import model as om
from . import train as ot

_ROOT = os.path.dirname(os.path.abspath(__file__))


def _c_lib():
    so = os.path.join(_ROOT, "_build", "liboracle_rnnt.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", _ROOT])
    return ctypes.CDLL(so)


def features(pcm_list, rates, target_dbs, offset=None, scale=None, cmn=True, specaug=None, lctx=1, rctx=1):
    """loader + trainer front end for a list of int16 utterances -> (data [B,T,240] f32, lens)."""
    feats = [fe.kaldi_fbank(fe.augment(p, r, g).astype(np.float32)) for p, r, g in zip(pcm_list, rates, target_dbs)]
    data, _, lens, _ = fe.assemble_batch(feats, [[1]] * len(feats), lctx, rctx, tu_limit=1 << 30)
    if offset is not None:
        data = fe.apply_cmvn(data, offset, scale, cmn)
    elif cmn:
        data = data - data.mean(axis=1, dtype=np.float32, keepdims=True)
    if specaug is not None:
        data = fe.spec_augment(data, *specaug)
    return data, lens


def train_step(sd, data, labels, frame_lens, label_lens, lr, momentum, max_norm, bufs=None):
    """sd: dict name -> torch fp32 tensor (parameters have requires_grad=True).  In place SGD update.
    Returns (costs [B] numpy, bufs)."""
    x = torch.from_numpy(np.ascontiguousarray(data))
    y = torch.from_numpy(np.ascontiguousarray(labels)).long()
    params = [v for v in sd.values() if v.requires_grad]
    for p in params:
        p.grad = None
    enc = om.encoder_forward(sd, x, train=True)
    pred = om.prednet_forward(sd, y)
    logits = om.joint_forward(sd, enc, pred, softmax=False)           # [B,T',U1,V]
    B, T, U1, V = logits.shape
    z = np.ascontiguousarray(logits.detach().numpy())
    costs = np.zeros(B)
    dz = np.empty_like(z)
    lab = np.ascontiguousarray(labels, np.int32)
    fl = np.ascontiguousarray(frame_lens, np.int32)
    ll = np.ascontiguousarray(label_lens, np.int32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = _c_lib().oracle_rnnt_loss(P(z), P(lab), P(fl), P(ll), B, T, U1, V, V, lab.shape[1], P(costs), P(dz))
    assert rc == 0
    logits.backward(torch.from_numpy(dz))
    grads = [p.grad.numpy() for p in params]
    _, coef = ot.clip_coef_inf(grads, max_norm) if max_norm > 0 else (0.0, 1.0)
    first = bufs is None
    if first:
        bufs = [None] * len(params)
    with torch.no_grad():
        for i, p in enumerate(params):
            g = p.grad * coef
            bufs[i] = g.clone() if first else bufs[i].mul_(momentum).add_(g)
            p.add_(g + momentum * bufs[i], alpha=-lr)
    return costs, bufs
