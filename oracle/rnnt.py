"""Oracle: RNN-T loss (alpha/beta lattice + gradient), numpy, log-space.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call site: trainer/train_transducer_bmuf_otfaug.py:58,97-99
    transducer_loss = RNNTLoss(blank=0, reduction='sum').apply
    loss = transducer_loss(outputs, target_batch.int(), len_batch, ali_lens); loss = loss.sum()
The arithmetic lives in the un-vendored ``warp_rnnt`` (README.md:34-36, no version pinned);
this restates its published recurrences (Graves 2012, eq. 16-20) with blank = 0:

    alpha[0,0] = 0
    alpha[t,u] = LSE(alpha[t-1,u] + lp[t-1,u,blank], alpha[t,u-1] + lp[t,u-1,y_u])
    beta[T-1,U] = lp[T-1,U,blank]
    beta[t,u]  = LSE(beta[t+1,u] + lp[t,u,blank], beta[t,u+1] + lp[t,u,y_{u+1}])
    cost = -beta[0,0]
    d cost / d lp[t,u,blank]   = -exp(alpha[t,u] + beta[t+1,u] + lp[t,u,blank] - ll)   (beta[T,U] := 0)
    d cost / d lp[t,u,y_{u+1}] = -exp(alpha[t,u] + beta[t,u+1] + lp[t,u,y_{u+1}] - ll)
    0 elsewhere and outside (T_n, U_n).
"""
import itertools
import numpy as np

NEG_INF = -np.inf


def _lse(a, b):
    m = np.maximum(a, b)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = m + np.log(np.exp(a - m) + np.exp(b - m))
    return np.where(np.isneginf(m), NEG_INF, r)


def log_softmax(x, axis=-1):
    """trainer/model/transducer.py:110-111 (F.log_softmax over V)."""
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=axis, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=axis, keepdims=True))


def rnnt_alpha_beta(lp_blank, lp_label, T, U):
    """lp_blank [T,U+1], lp_label [T,U] (label u+1 emitted at node (t,u)) -> alpha, beta [T,U+1]."""
    alpha = np.full((T, U + 1), NEG_INF)
    beta = np.full((T, U + 1), NEG_INF)
    alpha[0, 0] = 0.0
    for t in range(T):
        for u in range(U + 1):
            if t == 0 and u == 0:
                continue
            a = alpha[t - 1, u] + lp_blank[t - 1, u] if t > 0 else NEG_INF
            b = alpha[t, u - 1] + lp_label[t, u - 1] if u > 0 else NEG_INF
            alpha[t, u] = _lse(a, b)
    beta[T - 1, U] = lp_blank[T - 1, U]
    for t in range(T - 1, -1, -1):
        for u in range(U, -1, -1):
            if t == T - 1 and u == U:
                continue
            a = beta[t + 1, u] + lp_blank[t, u] if t < T - 1 else NEG_INF
            b = beta[t, u + 1] + lp_label[t, u] if u < U else NEG_INF
            beta[t, u] = _lse(a, b)
    return alpha, beta


def rnnt_loss(log_probs, labels, frame_lens, label_lens, blank=0):
    """log_probs [B,T,U+1,V] (already log-softmaxed), labels [B,U] int, lens [B] int.

    Returns (costs [B] f64, grads [B,T,U+1,V] f64 = d costs.sum() / d log_probs).
    """
    lp = np.asarray(log_probs, dtype=np.float64)
    B, Tm, U1, V = lp.shape
    costs = np.zeros(B)
    grads = np.zeros_like(lp)
    for n in range(B):
        T = int(frame_lens[n])
        U = int(label_lens[n])
        y = np.asarray(labels[n][:U], dtype=np.int64)
        lpb = lp[n, :T, :U + 1, blank]
        lpl = lp[n, :T, np.arange(U), y].T if U > 0 else np.zeros((T, 0))   # [T,U]
        alpha, beta = rnnt_alpha_beta(lpb, lpl, T, U)
        ll = beta[0, 0]
        costs[n] = -ll
        # blank grads: beta[t+1,u], terminal beta[T,U] := 0
        beta_next_t = np.full((T, U + 1), NEG_INF)
        beta_next_t[:T - 1] = beta[1:]
        beta_next_t[T - 1, U] = 0.0
        with np.errstate(invalid="ignore"):
            gb = -np.exp(alpha + beta_next_t + lpb - ll)
        gb[np.isnan(gb)] = 0.0
        grads[n, :T, :U + 1, blank] = gb
        if U > 0:
            with np.errstate(invalid="ignore"):
                gl = -np.exp(alpha[:, :U] + beta[:, 1:] + lpl - ll)
            gl[np.isnan(gl)] = 0.0
            for u in range(U):
                grads[n, :T, u, y[u]] += gl[:, u]
    return costs, grads


def rnnt_loss_from_logits(logits, labels, frame_lens, label_lens, blank=0):
    """Fused form used by the CUDA path: logits -> (costs, d costs.sum()/d logits).

    d/d logits of cost(log_softmax(logits)) = g - softmax * sum_v g   (chain rule through
    trainer/model/transducer.py:110-111).
    """
    z = np.asarray(logits, dtype=np.float64)
    lp = log_softmax(z)
    costs, g = rnnt_loss(lp, labels, frame_lens, label_lens, blank)
    dz = g - np.exp(lp) * g.sum(axis=-1, keepdims=True)
    return costs, dz


def rnnt_brute_force(log_probs, labels, T, U, blank=0):
    """-log sum over all monotone paths, by enumeration (tiny T,U only)."""
    lp = np.asarray(log_probs, dtype=np.float64)
    total = NEG_INF
    # a path = T blanks and U labels; last symbol must be the blank at (T-1,U)
    for pos in itertools.combinations(range(T + U - 1), U):
        pos = set(pos)
        t = u = 0
        s = 0.0
        for i in range(T + U - 1):
            if i in pos:
                s += lp[t, u, labels[u]]
                u += 1
            else:
                s += lp[t, u, blank]
                t += 1
        assert t == T - 1 and u == U
        s += lp[T - 1, U, blank]
        total = np.logaddexp(total, s)
    return -total
