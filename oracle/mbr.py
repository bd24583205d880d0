"""Oracle: MBR training batch restated with torch-CPU autograd, following
trainer/train_transducer_mbr_bmuf_otfaug.py:140-235 line by line (given an N-best list).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: PINNED -- the MBR trainer is a script that cannot be
imported, so tests/golden/make_golden.py (``golden_mbr``) reads the loop-body source text from the reference, and executes
it on CPU with the reference Net and TransducerDecoder (repairs: ``.cuda()`` -> CPU, ``editdistance.eval`` = Levenshtein,
warp_rnnt -> torchaudio rnnt_loss); tests/test_oracle_mbr.py checks this restatement against that run: MBR loss equal,
RNN-T loss to 3e-6, every parameter gradient to <= 2.5e-3 (norms <= 2e-4) on the high-gain decode fixture.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import model as om
from . import rnnt as orn


def levenshtein(a, b):
    d = np.zeros((len(a) + 1, len(b) + 1), dtype=np.int64)
    d[:, 0] = np.arange(len(a) + 1)
    d[0, :] = np.arange(len(b) + 1)
    for i in range(1, len(a) + 1):
        for j in range(1, len(b) + 1):
            d[i, j] = min(d[i - 1, j] + 1, d[i, j - 1] + 1, d[i - 1, j - 1] + (a[i - 1] != b[j - 1]))
    return int(d[len(a), len(b)])


def mbr_loss_and_grads(sd, x, target, tlens, ulens, hyps, scores, blk, padding_idx, rnnt_scale, sm_scale):
    """sd: name -> tensor (parameters require grad).  hyps[i][j]: list of ints (alignment incl. blanks);
    scores[i][j]: float.  Returns (mbr_loss, rnnt_costs); gradients are left in the parameters' .grad."""
    bsz, beam = len(hyps), len(hyps[0])
    bb = bsz * beam
    enc = om.encoder_forward(sd, x, train=True)                                   # :130-138
    T = enc.shape[1]
    # RNN-T branch (:140-161)
    pred = om.prednet_forward(sd, target, blank=blk)
    logits = om.joint_forward(sd, enc, pred, softmax=False)
    costs, dz = orn.rnnt_loss_from_logits(logits.detach().numpy(), target.numpy(), tlens, ulens)
    logits.backward(torch.from_numpy(dz * rnnt_scale).float(), retain_graph=True)
    # MBR branch (:164-235)
    xx = enc.unsqueeze(1).expand(-1, beam, -1, -1).contiguous().view(bb, T, -1)
    prob = F.softmax(torch.tensor(scores, dtype=torch.float32).view(bsz, beam), dim=1)
    nonblk = [[[t for t in h if t != blk] for h in row] for row in hyps]
    max_nb = max(len(h) for row in nonblk for h in row)
    U = max_nb + 1
    dist = torch.zeros(bsz, beam)
    for i in range(bsz):
        ref = target[i][:int(ulens[i])].tolist()
        for j in range(beam):
            dist[i, j] = levenshtein(ref, nonblk[i][j])
    avg = (prob * dist).sum(dim=1)
    mbr_loss = avg.sum()
    seq_grad = prob * (dist - avg.unsqueeze(1))
    y = torch.full((bb, max_nb), padding_idx, dtype=torch.long)
    for i in range(bsz):
        for j in range(beam):
            if nonblk[i][j]:
                y[i * beam + j, :len(nonblk[i][j])] = torch.tensor(nonblk[i][j])
    yp = om.prednet_forward(sd, y, blank=blk)                                     # [bb, U, H]
    assert yp.shape[1] == U
    V = sd["fc2.weight"].shape[0]
    mbr_grad = torch.zeros(bb, T + U, V)
    b_idx, x_idx, y_idx = [], [], []
    for i in range(bsz):
        for j in range(beam):
            h = hyps[i][j]
            t_idx, u_idx = [0], [0]
            for t in range(1, len(h)):
                t_idx.append(t_idx[t - 1] + int(h[t - 1] == blk))
                u_idx.append(u_idx[t - 1] + int(h[t - 1] != blk))
            t_idx.extend((T + U - len(t_idx)) * [0])
            u_idx.extend((T + U - len(u_idx)) * [0])
            x_idx.extend(t_idx); y_idx.extend(u_idx)
            b_idx.extend([i * beam + j] * (T + U))
            mbr_grad[i * beam + j, torch.arange(len(h)), torch.tensor(h, dtype=torch.long)] = seq_grad[i, j]
    joint = torch.cat((xx[b_idx, x_idx, :], yp[b_idx, y_idx, :]), dim=-1).view(bb, T + U, -1)
    out = F.linear(torch.tanh(F.linear(joint, sd["fc1.weight"], sd["fc1.bias"])) *
                   torch.sigmoid(F.linear(joint, sd["fc_gate.weight"], sd["fc_gate.bias"])), sd["fc2.weight"], sd["fc2.bias"])
    out = F.log_softmax(sm_scale * out, dim=-1)
    mbr_grad[:, :, blk] = mbr_grad[:, :, blk] / float(T)                           # :234
    out.backward(mbr_grad)
    return float(mbr_loss), costs
