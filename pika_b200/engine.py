"""Execution engine: the transducer's forward/backward expressed as autograd Functions whose
bodies are calls into the hand-written sm_100a kernels (pika_b200/csrc) through the C ABI.

torch supplies device memory, streams and the autograd tape; no torch operator computes anything
on this path (boundary dtype casts and view/reshape bookkeeping aside).

Precision modes (``set_precision``):
  * "bf16" (production): activations and GEMM operands bf16, fp32 accumulation in TMEM, fp32
    statistics / master weights / gradients.
  * "fp32" (parity): activations fp32; every GEMM runs as three bf16 tensor-core products on hi/lo
    split operands (A_hi B_hi + A_hi B_lo + A_lo B_hi), i.e. fp32-class accuracy from the same kernel.

Reference semantics reproduced (paths relative to the reference root): trainer/model/transducer.py:74-112,
trainer/model/rnnt_tdnn_transformer.py:73-89, trainer/model/modules/{transformer.py:85-100,
multi_headed_attn.py:110-241, position_ffn.py:27-39}.
"""
import math
import os
import weakref

import torch

from . import kernels as K

_PRECISION = "bf16"
_WEIGHT_EPOCH = 0           # bumped whenever parameters are modified through raw pointers
_SEED = [0x5EED]
_DROPOUT_ENABLED = True


def set_precision(p):
    global _PRECISION
    assert p in ("bf16", "fp32")
    _PRECISION = p


def get_precision():
    return _PRECISION


def act_dtype():
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


def invalidate_weights():
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def set_seed(seed):
    _SEED[0] = seed & 0x7FFFFFFF


def set_dropout_enabled(flag):
    """Parity runs disable dropout while BatchNorm stays in train mode."""
    global _DROPOUT_ENABLED
    _DROPOUT_ENABLED = bool(flag)


def _next_seed():
    _SEED[0] = (_SEED[0] * 1103515245 + 12345) & 0x7FFFFFFF
    return _SEED[0]


def _drop(p, training):
    return float(p) if (training and _DROPOUT_ENABLED and p > 0) else 0.0


# ------------------------------------------------------------------------------------------------
# operand staging
def stage_act(x):
    """Activation (act dtype, any shape, contiguous) -> GEMM operand parts [hi] or [hi, lo] (bf16)."""
    if x.dtype == torch.bfloat16:
        return [x]
    x2 = x.reshape(-1, x.shape[-1])
    hi = torch.empty(x2.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    K.cast_split(x2, hi, lo)
    return [hi.view(x.shape), lo.view(x.shape)]


_wcache = {}


def _prune_wcache():
    if len(_wcache) > 512:
        for k in [k for k, v in _wcache.items() if any(r() is None for r in v[2])]:
            del _wcache[k]


def _same_objects(refs, params):
    """id() values are recycled: a cache entry only counts if its weak references still point at these very tensors
    (a freed parameter whose id, version and storage address all reappear would otherwise alias a stale staging)."""
    return len(refs) == len(params) and all(r() is p for r, p in zip(refs, params))


def stage_weight(params, cols_pad=None, scale=None):
    """Parameter(s) -> staged bf16 operand parts of the row-concatenated matrix [sum N_i, K(_pad)].
    Cached until the parameters change (torch version counter or ``invalidate_weights``)."""
    params = params if isinstance(params, (list, tuple)) else [params]
    key = (tuple(id(p) for p in params), _PRECISION, cols_pad)
    stamp = (tuple(p._version for p in params), _WEIGHT_EPOCH, tuple(p.data_ptr() for p in params))
    hit = _wcache.get(key)
    if hit is not None and hit[0] == stamp and _same_objects(hit[2], params):
        return hit[1]
    mats = [p.detach().reshape(p.shape[0], -1) for p in params]
    Kdim = mats[0].shape[1]
    kp = cols_pad or Kdim
    n = sum(m.shape[0] for m in mats)
    hi = torch.empty(n, kp, dtype=torch.bfloat16, device=mats[0].device)
    lo = torch.empty_like(hi) if _PRECISION == "fp32" else None
    r = 0
    for i, m in enumerate(mats):
        K.cast_split(m, hi[r:r + m.shape[0]], None if lo is None else lo[r:r + m.shape[0]], cols_pad=kp,
                     scale=1.0 if scale is None else scale[i])
        r += m.shape[0]
    parts = [hi] if lo is None else [hi, lo]
    _prune_wcache()
    _wcache[key] = (stamp, parts, tuple(weakref.ref(p) for p in params))
    return parts


def stage_conv1d_weight(weight, ld):
    """nn.Conv1d weight [N, C, Kw] -> staged bf16 parts of the tap-major matrix [N, Kw * ld] (tap k at columns [k*ld, k*ld + C),
    zero padding up to the activation pitch ``ld``): the B operand of the causal convolution written as one GEMM over
    overlapping (Toeplitz) activation rows.  Cached like ``stage_weight``."""
    key = (id(weight), _PRECISION, "conv1d", ld)
    stamp = (weight._version, _WEIGHT_EPOCH, weight.data_ptr())
    hit = _wcache.get(key)
    if hit is not None and hit[0] == stamp and _same_objects(hit[2], [weight]):
        return hit[1]
    N, C, Kw = weight.shape
    mat = torch.zeros(N, Kw, ld, dtype=torch.float32, device=weight.device)
    mat[:, :, :C] = weight.detach().permute(0, 2, 1)
    mat = mat.view(N, Kw * ld)
    hi = torch.empty(N, Kw * ld, dtype=torch.bfloat16, device=weight.device)
    lo = torch.empty_like(hi) if _PRECISION == "fp32" else None
    K.cast_split(mat, hi, lo)
    parts = [hi] if lo is None else [hi, lo]
    _prune_wcache()
    _wcache[key] = (stamp, parts, (weakref.ref(weight),))
    return parts


# opt-in: same-box A/B of the full step showed no gain (88.2 vs 87.9 ms) although isolated dgrads are 5-18 % faster with a
# K-major B on random data; the extra transposes cancel it (profiles/r01_notes.md)
_DGRAD_KMAJOR = os.environ.get("PK_DGRAD_KMAJOR", "0") != "0"


def transposed_parts(parts):
    """Transposed bf16 copies [K, N] of staged weight parts [N, K] so a dgrad reads its B operand K-major
    (measured 5-18 % faster than MN-major B, profiles/r01_notes.md).  The copy hangs off the staged tensor, so it
    lives exactly as long as that staging does."""
    out = []
    for p in parts:
        t = getattr(p, "_pk_transposed", None)
        if t is None:
            t = torch.empty(p.shape[1], p.shape[0], dtype=torch.bfloat16, device=p.device)
            K.transpose_bf16(p, t)
            p._pk_transposed = t
        out.append(t)
    return out


def dgrad_b(parts, rows=None, cols=None):
    """B operand of dx = dy @ W for staged W parts [N, K] (optionally the sub-block rows x cols):
    returns (b_parts, b_mn)."""
    if _DGRAD_KMAJOR and all(p.shape[0] % 8 == 0 for p in parts):
        t = transposed_parts(parts)                      # [K, N]
        if cols is not None:
            t = [p[cols[0]:cols[1]] for p in t]
        if rows is not None:
            t = [p[:, rows[0]:rows[1]] for p in t]
        return t, False
    sub = parts
    if rows is not None:
        sub = [p[rows[0]:rows[1]] for p in sub]
    if cols is not None:
        sub = [p[:, cols[0]:cols[1]] for p in sub]
    return sub, True


def _cat_bias(params):
    """Concatenated f32 bias for a row-concatenated weight (cached like the weights)."""
    if len(params) == 1:
        return params[0].detach()
    key = (tuple(id(p) for p in params), "bias")
    stamp = (tuple(p._version for p in params), _WEIGHT_EPOCH, tuple(p.data_ptr() for p in params))
    hit = _wcache.get(key)
    if hit is not None and hit[0] == stamp and _same_objects(hit[2], params):
        return hit[1]
    out = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    r = 0
    for p in params:
        out[r:r + p.numel()].copy_(p.detach())
        r += p.numel()
    _wcache[key] = (stamp, out, tuple(weakref.ref(p) for p in params))
    return out


def gemm_parts(a_taps, b_taps, c, **kw):
    """Accumulate sum over taps of A_tap @ B_tap^T where each tap is a parts list ([hi] | [hi, lo])."""
    A, Bm, aro, bro = [], [], [], []
    a_off = kw.pop("a_row_off", None)
    b_off = kw.pop("b_row_off", None)
    for t, (ap, bp) in enumerate(zip(a_taps, b_taps)):
        combos = [(0, 0)]
        if len(ap) > 1 and len(bp) > 1:
            combos += [(0, 1), (1, 0)]
        elif len(ap) > 1:
            combos += [(1, 0)]
        elif len(bp) > 1:
            combos += [(0, 1)]
        for (i, j) in combos:
            A.append(ap[i])
            Bm.append(bp[j])
            aro.append(a_off[t] if a_off else 0)
            bro.append(b_off[t] if b_off else 0)
    return K.gemm(A, Bm, c, a_row_off=aro, b_row_off=bro, **kw)


def grad_of(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


def _new(shape, dtype=None, like=None, zero=False):
    dev = like.device if like is not None else "cuda"
    f = torch.zeros if zero else torch.empty
    return f(shape, dtype=dtype or act_dtype(), device=dev)


# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = dropout(act(x W^T + b)) + residual; W = row-concatenation of ``weights``."""

    @staticmethod
    def forward(ctx, x, residual, act, drop_p, seed, nw, premasked, mask_input_scale, *params):
        # premasked: the incoming gradient is already d(pre-activation) -- the consumer's backward applied this layer's
        #   ReLU(+dropout) mask (BatchNorm backward with relu_mask, or the next Linear's dgrad epilogue), so no mask pass runs here.
        # mask_input_scale (> 0): this layer's INPUT x is relu(+dropout) output of the previous Linear; its dgrad epilogue
        #   multiplies dx by (x != 0) * scale, i.e. hands the previous layer d(pre-activation) directly (AUX_MASK_NZ).
        weights = list(params[:nw])
        biases = list(params[nw:]) if len(params) > nw else None
        M, Kd = x.shape
        w_parts = stage_weight(weights, cols_pad=x.shape[1] if x.shape[1] != weights[0].shape[1] else None)
        N = w_parts[0].shape[0]
        y = _new((M, N), like=x)
        bias = _cat_bias(biases) if biases is not None else None
        a_parts = stage_act(x)
        gemm_parts([a_parts], [w_parts], y, bias=bias, act=K.ACT_RELU if act else K.ACT_NONE, drop_p=drop_p,
                   drop_seed=seed, aux=residual, aux_mode=K.AUX_ADD if residual is not None else K.AUX_NONE)
        ctx.weights, ctx.biases, ctx.act, ctx.drop_p, ctx.seed = weights, biases, act, drop_p, seed
        ctx.premasked, ctx.mask_input_scale = premasked, mask_input_scale
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, y if act else None)
        ctx.a_parts, ctx.w_parts = a_parts, w_parts
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dy = dy.contiguous()
        M, N = dy.shape
        if ctx.premasked:
            dpre = dy
        elif ctx.act:
            dpre = torch.empty_like(dy)
            K.mask_nz(dy, y, dpre, 1.0 / (1.0 - ctx.drop_p) if ctx.drop_p > 0 else 1.0)
        elif ctx.drop_p > 0:
            dpre = torch.empty_like(dy)
            K.dropout(dy, dpre, ctx.drop_p, ctx.seed)
        else:
            dpre = dy
        d_parts = stage_act(dpre)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            wb, wmn = dgrad_b(ctx.w_parts)
            if ctx.mask_input_scale > 0:
                gemm_parts([d_parts], [wb], dx, b_mn=wmn, aux=x, aux_mode=K.AUX_MASK_NZ, aux_scale=ctx.mask_input_scale)
            else:
                gemm_parts([d_parts], [wb], dx, b_mn=wmn)
            if dx.shape[1] != x.shape[1]:
                dx = dx[:, :x.shape[1]]
        # dW_i = dpre[:, rows_i]^T x ; db_i = colsum(dpre)[rows_i]
        r = 0
        dbias = None
        if ctx.biases is not None:
            dbias = torch.empty(N, dtype=torch.float32, device=dy.device)
            K.colsum(dpre, dbias)
        for i, w in enumerate(ctx.weights):
            n_i = w.shape[0]
            g = grad_of(w).view(n_i, -1)
            gemm_parts([[p[:, r:r + n_i] for p in d_parts]], [[p[:, :g.shape[1]] for p in ctx.a_parts]], g, a_mn=True, b_mn=True)
            if ctx.biases is not None:
                grad_of(ctx.biases[i]).copy_(dbias[r:r + n_i])
            r += n_i
        n_par = len(ctx.weights) + (len(ctx.biases) if ctx.biases is not None else 0)
        return (dx, (dy if ctx.has_res else None), None, None, None, None, None, None) + (None,) * n_par


def linear(x, weights, biases=None, act=False, drop_p=0.0, residual=None, premasked=False, mask_input_scale=0.0):
    weights = list(weights) if isinstance(weights, (list, tuple)) else [weights]
    if biases is not None and not isinstance(biases, (list, tuple)):
        biases = [biases]
    params = weights + (list(biases) if biases is not None else [])
    if x.dtype != torch.bfloat16 or x.shape[1] % 8 != 0:
        mask_input_scale = 0.0                 # the epilogue mask reads x as a 16-byte-aligned aux operand of the activation dtype
    return LinearFn.apply(x, residual, act, drop_p, _next_seed() if drop_p > 0 else 0, len(weights), premasked, mask_input_scale, *params)


class TdnnFn(torch.autograd.Function):
    """relu(Conv2d(1, N, (3, C), dilation=(d,1), stride=(s,1))) on [B,T,C] as three accumulated GEMM
    taps over strided views (trainer/model/rnnt_tdnn_transformer.py:44-59, 81-82)."""

    @staticmethod
    def forward(ctx, x, weight, bias, dil, stride, premasked=False):
        B, T, C = x.shape
        ctx.premasked = premasked
        N = weight.shape[0]
        t_out = (T - 2 * dil - 1) // stride + 1
        w_parts = stage_weight(weight)                 # [N, 3*C] row-major == weight[:, 0, k, :] at cols k*C
        a_parts = stage_act(x)
        span = (t_out - 1) * stride + 1
        a_taps = [[p[:, k * dil: k * dil + span: stride, :] for p in a_parts] for k in range(3)]
        b_taps = [[p[:, k * C:(k + 1) * C] for p in w_parts] for k in range(3)]
        y = _new((B, t_out, N), like=x)
        gemm_parts(a_taps, b_taps, y, a_sel=(K.SEL_ZB0, K.SEL_ZERO), b_sel=(K.SEL_ZERO, K.SEL_ZERO), bias=bias.detach(),
                   act=K.ACT_RELU)
        ctx.save_for_backward(x, y)
        ctx.a_parts, ctx.w_parts, ctx.weight, ctx.bias = a_parts, w_parts, weight, bias
        ctx.dil, ctx.stride, ctx.t_out = dil, stride, t_out
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        B, T, C = x.shape
        N = ctx.weight.shape[0]
        dil, stride, t_out = ctx.dil, ctx.stride, ctx.t_out
        if ctx.premasked:
            dpre = dy.contiguous()                  # BatchNormFn.backward already masked by (y > 0)
        else:
            dpre = torch.empty_like(y)
            K.mask_nz(dy.contiguous(), y, dpre, 1.0)
        d_parts = stage_act(dpre)
        b_taps, wmn = [], True
        for k in range(3):
            bt, wmn = dgrad_b(ctx.w_parts, cols=(k * C, (k + 1) * C))
            b_taps.append(bt)
        dx = None
        if ctx.needs_input_grad[0]:
            if stride == 1:
                dx = torch.empty_like(x)
                gemm_parts([d_parts] * 3, b_taps, dx, b_mn=wmn, a_sel=(K.SEL_ZB0, K.SEL_ZERO),
                           b_sel=(K.SEL_ZERO, K.SEL_ZERO), a_row_off=[0, -dil, -2 * dil])
            else:
                # rows tau = k*dil + t*stride of the three taps are disjoint when the residues differ
                assert len({(k * dil) % stride for k in range(3)}) == 3, "overlapping strided taps not supported"
                dx = torch.zeros_like(x)
                span = (t_out - 1) * stride + 1
                for k in range(3):
                    gemm_parts([d_parts], [b_taps[k]], dx[:, k * dil: k * dil + span: stride, :], b_mn=wmn,
                               a_sel=(K.SEL_ZB0, K.SEL_ZERO), b_sel=(K.SEL_ZERO, K.SEL_ZERO))
        gw = grad_of(ctx.weight).view(N, 3 * C)
        span = (t_out - 1) * stride + 1
        for k in range(3):
            xt = [p[:, k * dil: k * dil + span: stride, :] for p in ctx.a_parts]
            gemm_parts([d_parts], [xt], gw[:, k * C:(k + 1) * C], a_mn=True, b_mn=True, a_sel=(K.SEL_KZ, K.SEL_ZERO),
                       b_sel=(K.SEL_KZ, K.SEL_ZERO), kz_count=B)
        K.colsum(dpre.view(B * t_out, N), grad_of(ctx.bias))
        return dx, None, None, None, None, None


class CausalConvFn(torch.autograd.Function):
    """relu(Conv1d(C, N, Kw, padding=Kw-1)(x)[..., :-(Kw-1)]) on [B,T,ld] (ld >= C, pad columns zero) -- the causal convolution of
    the transformer prediction net (trainer/model/rnnt_conv_transformer_lm.py:36-46, 74-75).
    Forward: ONE GEMM.  Row (b, t) of the im2col matrix is the window x[b, t-Kw+1 .. t, :], which is contiguous in a left-padded
    [B, T+Kw-1, ld] buffer, so the A operand is an overlapping strided view of that buffer (row pitch ld, K = Kw*ld) and the B
    operand the tap-major staged weight.  dgrad: Kw accumulated taps with row offsets (rows past T read as zero); wgrad: the same
    overlapping view read MN-major, one batched-reduction GEMM into the tap-major gradient, permuted into the [N, C, Kw] layout."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        B, T, ld = x.shape
        N, C, Kw = weight.shape
        assert ld >= C and ld % 8 == 0
        xp = _new((B, T + Kw - 1, ld), like=x, zero=True)
        xp[:, Kw - 1:, :] = x
        a_parts = [p.as_strided((B, T, Kw * ld), ((T + Kw - 1) * ld, ld, 1)) for p in stage_act(xp)]
        w_parts = stage_conv1d_weight(weight, ld)
        y = _new((B, T, N), like=x)
        gemm_parts([a_parts], [w_parts], y, a_sel=(K.SEL_ZB0, K.SEL_ZERO), b_sel=(K.SEL_ZERO, K.SEL_ZERO), bias=bias.detach(),
                   act=K.ACT_RELU)
        ctx.save_for_backward(y)
        ctx.a_parts, ctx.w_parts, ctx.weight, ctx.bias, ctx.shape = a_parts, w_parts, weight, bias, (B, T, ld, N, C, Kw)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        B, T, ld, N, C, Kw = ctx.shape
        dpre = torch.empty_like(y)
        K.mask_nz(dy.contiguous(), y, dpre, 1.0)
        d_parts = stage_act(dpre)
        dx = None
        if ctx.needs_input_grad[0]:
            # dx[t] = sum_k dpre[t + Kw-1-k] W_k; at most 9 (A,B) pairs per launch: 3 taps in the split-bf16 mode, all of them in bf16
            per = 3 if len(d_parts) > 1 else Kw
            dx = _new((B, T, ld), dtype=torch.float32 if len(d_parts) > 1 else None, like=y)
            for k0 in range(0, Kw, per):
                ks = list(range(k0, min(Kw, k0 + per)))
                b_taps, wmn = [], True
                for k in ks:
                    bt, wmn = dgrad_b(ctx.w_parts, cols=(k * ld, (k + 1) * ld))
                    b_taps.append(bt)
                gemm_parts([d_parts] * len(ks), b_taps, dx, b_mn=wmn, a_sel=(K.SEL_ZB0, K.SEL_ZERO), b_sel=(K.SEL_ZERO, K.SEL_ZERO),
                           a_row_off=[Kw - 1 - k for k in ks], accumulate=k0 > 0)
        gw = torch.empty(N, Kw * ld, dtype=torch.float32, device=y.device)
        gemm_parts([d_parts], [ctx.a_parts], gw, a_mn=True, b_mn=True, a_sel=(K.SEL_KZ, K.SEL_ZERO), b_sel=(K.SEL_KZ, K.SEL_ZERO),
                   kz_count=B)
        grad_of(ctx.weight).copy_(gw.view(N, Kw, ld)[:, :, :C].permute(0, 2, 1))
        K.colsum(dpre.view(B * T, N), grad_of(ctx.bias))
        return dx, None, None


class BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bn, train, _w=None, _b=None, relu_input=False):
        # relu_input: x is a ReLU output whose producer was told ``premasked``; backward returns d(pre-ReLU) = dx * (x > 0)
        ctx.relu_input = relu_input
        rows, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        ws = torch.empty(2 * C, dtype=torch.float32, device=x.device)
        K.bn_fwd(x, y, bn.weight.detach(), bn.bias.detach(), bn.eps, train, bn.momentum if bn.momentum is not None else 0.1,
                 bn.running_mean, bn.running_var, mean, rstd, ws)
        if train and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.save_for_backward(x, mean, rstd)
        ctx.bn, ctx.train = bn, train
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        dx = torch.empty_like(x)
        K.bn_bwd(dy.contiguous(), x, dx, ctx.bn.weight.detach(), mean, rstd, ctx.train, ctx.relu_input, grad_of(ctx.bn.weight),
                 grad_of(ctx.bn.bias))
        return dx, None, None, None, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln, _w=None, _b=None):
        rows, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        K.layernorm_fwd(x, y, ln.weight.detach(), ln.bias.detach(), ln.eps, mean, rstd)
        ctx.save_for_backward(x, mean, rstd)
        ctx.ln = ln
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        dx = torch.empty_like(x)
        K.layernorm_bwd(dy.contiguous(), x, dx, ctx.ln.weight.detach(), mean, rstd, grad_of(ctx.ln.weight), grad_of(ctx.ln.bias))
        return dx, None, None, None


_FUSED_ATTN = os.environ.get("PK_FUSED_ATTN", "1") != "0"
_FOLD_MASKS = os.environ.get("PK_FOLD_MASKS", "1") != "0"


class AttentionFn(torch.autograd.Function):
    """Multi-head self-attention on a fused QKV tensor [B,T,3D] (q | k | v blocks):
    softmax((Q/sqrt(d)) K^T [masked]) -> dropout -> V   (trainer/model/modules/multi_headed_attn.py:199-223).
    Unmasked (the encoder): the fused tcgen05 kernels.  ``causal`` / ``key_pad`` (uint8 [B,T], 1 = padding key; the transformer
    prediction net, trainer/model/rnnt_conv_transformer_lm.py:66-70): batched GEMMs + the masked softmax kernel (short label
    sequences; the mask only enters the forward softmax -- a dropped key has probability 0, so its dS is 0 as well)."""

    @staticmethod
    def forward(ctx, qkv, heads, drop_p, seed, causal=False, key_pad=None):
        B, T, D3 = qkv.shape
        D = D3 // 3
        dh = D // heads
        masked = causal or key_pad is not None
        ctx.fused = _FUSED_ATTN and qkv.dtype == torch.bfloat16 and dh == 64 and not masked
        if ctx.fused:
            # scores / probabilities never leave the SM (pika_b200/csrc/attention_tc.cu)
            qkv = qkv.contiguous()
            out = _new((B, T, D), like=qkv)
            lse = torch.zeros(B * heads * K.attention_lse_stride(T), dtype=torch.float32, device=qkv.device)   # pad entries finite
            alpha = 1.0 / math.sqrt(dh)
            K.attention_fwd(qkv, out, lse, heads, alpha, drop_p, seed)
            ctx.save_for_backward(qkv, out, lse)
            ctx.meta = (B, T, D, heads, dh, 0, drop_p, seed, alpha)
            return out
        Tp = (T + 7) // 8 * 8
        parts = stage_act(qkv)

        def head_view(p, which):
            return p[:, :, which * D:(which + 1) * D].view(B, T, heads, dh).permute(0, 2, 1, 3)

        q, k, v = ([head_view(p, i) for p in parts] for i in range(3))
        S = torch.empty(B, heads, T, Tp, dtype=torch.float32, device=qkv.device)
        alpha = 1.0 / math.sqrt(dh)
        gemm_parts([q], [k], S[:, :, :, :T], alpha=alpha)
        P = _new((B, heads, T, Tp), like=qkv)
        Pd = _new((B, heads, T, Tp), like=qkv) if drop_p > 0 else P
        if masked:
            K.softmax_masked_fwd(S, P, Pd, T, T, heads, causal, key_pad, drop_p, seed)
        else:
            K.softmax_fwd(S, P, Pd, T, drop_p, seed)
        del S
        out = _new((B, T, D), like=qkv)
        pd_parts = stage_act(Pd)
        gemm_parts([[p[:, :, :, :T] for p in pd_parts]], [v], out.view(B, T, heads, dh).permute(0, 2, 1, 3), b_mn=True)
        ctx.save_for_backward(qkv, P)
        ctx.pd_parts, ctx.parts = pd_parts, parts
        ctx.meta = (B, T, D, heads, dh, Tp, drop_p, seed, alpha)
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.fused:
            qkv, out, lse = ctx.saved_tensors
            B, T, D, heads, dh, _, drop_p, seed, alpha = ctx.meta
            dqkv = torch.empty_like(qkv)
            K.attention_bwd(qkv, out, dout.contiguous(), lse, dqkv, heads, alpha, drop_p, seed)
            return dqkv, None, None, None, None, None
        qkv, P = ctx.saved_tensors
        B, T, D, heads, dh, Tp, drop_p, seed, alpha = ctx.meta
        dout = dout.contiguous()

        def head_view(p, which, d=D):
            return p[:, :, which * d:(which + 1) * d].view(B, T, heads, dh).permute(0, 2, 1, 3)

        q, k, v = ([head_view(p, i) for p in ctx.parts] for i in range(3))
        do_parts = stage_act(dout)
        do = [p.view(B, T, heads, dh).permute(0, 2, 1, 3) for p in do_parts]
        dqkv = torch.empty_like(qkv)
        # dV = Pd^T dO
        gemm_parts([[p[:, :, :, :T] for p in ctx.pd_parts]], [do], head_view(dqkv, 2), a_mn=True, b_mn=True)
        # dPd = dO V^T
        dPd = torch.empty(B, heads, T, Tp, dtype=torch.float32, device=qkv.device)
        gemm_parts([do], [v], dPd[:, :, :, :T])
        dS = torch.empty_like(P)
        K.softmax_bwd(dPd, P, dS, T, drop_p, seed)
        del dPd
        ds_parts = [p[:, :, :, :T] for p in stage_act(dS)]
        # dQ = alpha * dS K ; dK = alpha * dS^T Q
        gemm_parts([ds_parts], [k], head_view(dqkv, 0), b_mn=True, alpha=alpha)
        gemm_parts([ds_parts], [q], head_view(dqkv, 1), a_mn=True, b_mn=True, alpha=alpha)
        return dqkv, None, None, None, None, None


class EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, emb, ld, _w=None):
        n = idx.numel()
        out = _new((n, ld), like=emb.weight)
        K.embedding_fwd(idx.reshape(-1).contiguous(), emb.weight.detach(), out)
        ctx.idx, ctx.emb = idx.reshape(-1).contiguous(), emb
        return out

    @staticmethod
    def backward(ctx, dout):
        K.embedding_bwd(ctx.idx, dout.contiguous(), grad_of(ctx.emb.weight), ctx.emb.padding_idx)
        return None, None, None, None


class LstmLayerFn(torch.autograd.Function):
    """One nn.LSTM layer (batch_first, zero initial state) over x [B,U,E]: the input projection is one
    GEMM over all steps; each step is a recurrent GEMM (h_{t-1} W_hh^T) + a fused cell kernel."""

    @staticmethod
    def forward(ctx, x, lstm, layer, w_ih, w_hh, b_ih, b_hh):
        B, U, E = x.shape
        H = w_hh.shape[1]
        wih_parts = stage_weight(w_ih, cols_pad=E if E != w_ih.shape[1] else None)
        whh_parts = stage_weight(w_hh)
        bsum = torch.empty(4 * H, dtype=torch.float32, device=x.device)
        K.add(b_ih.detach(), b_hh.detach(), bsum)
        x_parts = stage_act(x)
        gx = torch.empty(B, U, 4 * H, dtype=torch.float32, device=x.device)
        gemm_parts([[p.view(B * U, E) for p in x_parts]], [wih_parts], gx.view(B * U, 4 * H), bias=bsum)
        out = _new((B, U, H), like=x)
        gates = torch.empty(U, B, 4 * H, dtype=torch.float32, device=x.device)
        cs = torch.empty(U, B, H, dtype=torch.float32, device=x.device)
        ctx.persistent = (x.dtype == torch.bfloat16 and H % 64 == 0 and H // 8 <= 148)      # any batch: 32 sequences per cooperative launch
        if ctx.persistent:
            # whole recurrence in one cooperative launch (pika_b200/csrc/lstm_seq.cu)
            K.lstm_seq_fwd(gx, whh_parts[0], out, gates, cs)
        gh = torch.empty(B, 4 * H, dtype=torch.float32, device=x.device)
        for t in range(0 if not ctx.persistent else U, U):
            if t > 0:
                gemm_parts([stage_act_view(out[:, t - 1, :])], [whh_parts], gh, block_n=64)
            K.lstm_cell_fwd(gx[:, t, :], gh if t > 0 else None, cs[t - 1] if t > 0 else None, cs[t], out[:, t, :], gates[t], B, H)
        ctx.save_for_backward(x, out, gates, cs)
        ctx.x_parts, ctx.wih_parts, ctx.whh_parts = x_parts, wih_parts, whh_parts
        ctx.params = (w_ih, w_hh, b_ih, b_hh)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, out, gates, cs = ctx.saved_tensors
        w_ih, w_hh, b_ih, b_hh = ctx.params
        B, U, E = x.shape
        H = w_hh.shape[1]
        dout = dout.contiguous()
        dG = _new((U, B, 4 * H), like=out)                 # time-major: dG[t] is a contiguous [B,4H] matrix
        dh_rec = torch.empty(B, H, dtype=torch.float32, device=x.device)
        dc = [torch.empty(B, H, dtype=torch.float32, device=x.device) for _ in range(2)]
        if ctx.persistent:
            K.lstm_seq_bwd(dout, gates, cs, ctx.whh_parts[0], dG)
        for t in range(U - 1 if not ctx.persistent else -1, -1, -1):
            last = (t == U - 1)
            K.lstm_cell_bwd(dout[:, t, :], None if last else dh_rec, None if last else dc[(t + 1) & 1], gates[t], cs[t],
                            cs[t - 1] if t > 0 else None, dG[t], dc[t & 1], B, H)
            if t > 0:
                gemm_parts([stage_act_view(dG[t])], [ctx.whh_parts], dh_rec, b_mn=True, block_n=64)
        dg_parts = stage_act(dG)                            # [U,B,4H]
        out_parts = stage_act(out)                          # [B,U,H]
        sel = dict(a_sel=(K.SEL_KZ, K.SEL_ZERO), b_sel=(K.SEL_KZ, K.SEL_ZERO))
        # dW_hh = sum_{t>=1} dG[t]^T h[t-1]   (reduction over batch rows, batched over t)
        if U > 1:
            gemm_parts([[p[1:] for p in dg_parts]], [[p.permute(1, 0, 2)[:-1] for p in out_parts]], grad_of(w_hh),
                       a_mn=True, b_mn=True, kz_count=U - 1, **sel)
        else:
            grad_of(w_hh).zero_()
        gemm_parts([dg_parts], [[p.permute(1, 0, 2) for p in ctx.x_parts]], grad_of(w_ih), a_mn=True, b_mn=True, kz_count=U, **sel)
        K.colsum(dG.view(U * B, 4 * H), grad_of(b_ih))
        grad_of(b_hh).copy_(b_ih.grad)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            wb, wmn = dgrad_b(ctx.wih_parts)
            gemm_parts([dg_parts], [wb], dx.permute(1, 0, 2), b_mn=wmn, a_sel=(K.SEL_ZB0, K.SEL_ZERO),
                       b_sel=(K.SEL_ZERO, K.SEL_ZERO))
        return dx, None, None, None, None, None, None


def stage_act_view(v):
    """stage a (possibly strided-row) 2-D activation view [B, H]."""
    if v.dtype == torch.bfloat16:
        return [v]
    hi = torch.empty(v.shape, dtype=torch.bfloat16, device=v.device)
    lo = torch.empty_like(hi)
    K.cast_split(v, hi, lo)
    return [hi, lo]


class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        y = torch.empty_like(x)
        K.dropout(x.contiguous(), y, p, seed)
        ctx.p, ctx.seed = p, seed
        return y

    @staticmethod
    def backward(ctx, dy):
        dx = torch.empty_like(dy)
        K.dropout(dy.contiguous(), dx, ctx.p, ctx.seed)
        return dx, None, None


def _ldv(V):
    return (V + 7) // 8 * 8


# on by default: same-box A/B 84.5 vs 86.2 ms/step (the fc2 GEMM gets ~19 % slower, the 13.9 GB first pass of the loss goes away);
# PK_FUSED_LSE=0 restores the stand-alone first pass (profiles/r01_notes.md)
_FUSED_LSE = os.environ.get("PK_FUSED_LSE", "1") != "0"


# measurement hook (bench.py): when set to a dict, single launches of the step are bracketed by CUDA events on the launching stream,
# e.g. EVENT_TAPS["fc2_fwd"] = [(start, end), ...] -- the duration of that kernel INSIDE a real step, not in a loop of its own
EVENT_TAPS = None


class _Tap:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if EVENT_TAPS is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *exc):
        if EVENT_TAPS is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            EVENT_TAPS.setdefault(self.name, []).append((self.s, e))


def _joint_forward(enc, pred, model, want_lse=False):
    """factored gated joint -> (logits [B,T,U1,ldv] act dtype, saved state).  ``want_lse``: the fc2 GEMM also
    reduces every logits row to per-tile (max, sum-exp) pairs (state["row_lse"]) for the fused loss."""
    B, T, H = enc.shape
    U1 = pred.shape[1]
    V = model.fc2.weight.shape[0]
    ldv = _ldv(V)
    fc1, fcg, fc2 = model.fc1, model.fc_gate, model.fc2
    wx = stage_weight([fc1.weight, fcg.weight])                       # [2H, 2H]; x half = cols [0,H), y half = [H,2H)
    enc_parts = stage_act(enc.reshape(B * T, H))
    pred_parts = stage_act(pred.reshape(B * U1, H))
    ex = _new((B * T, 2 * H), like=enc)
    py = _new((B * U1, 2 * H), like=enc)
    gemm_parts([enc_parts], [[p[:, :H] for p in wx]], ex, bias=_cat_bias([fc1.bias, fcg.bias]))
    gemm_parts([pred_parts], [[p[:, H:] for p in wx]], py)
    R = B * T * U1
    h = _new((R, H), like=enc)
    K.joint_gate_fwd(ex, py, h, B, T, U1, H)
    w2 = stage_weight(fc2.weight)
    logits = _new((B, T, U1, ldv), like=enc, zero=(ldv != V))
    h_parts = stage_act(h)
    row_lse = None
    if want_lse and _FUSED_LSE and logits.dtype == torch.bfloat16 and V % 8 == 0:
        row_lse = torch.empty(K.row_lse_parts(R, V, 256), R, 2, dtype=torch.float32, device=enc.device)
    with _Tap("fc2_fwd"):
        gemm_parts([h_parts], [w2], logits.view(R, ldv)[:, :V], bias=fc2.bias.detach(), row_lse=row_lse,
                   **({"block_n": 256} if row_lse is not None else {}))
    state = dict(row_lse=row_lse, ex=ex, py=py, h_parts=h_parts, enc_parts=enc_parts, pred_parts=pred_parts, wx=wx, w2=w2, dims=(B, T, U1, H, V, ldv))
    return logits, state


def _joint_backward(dlogits, st, model, need_enc=True, need_pred=True, db2=None):
    """dlogits [B,T,U1,ldv] (padding columns zero) -> (d_enc, d_pred); parameter grads written in place."""
    B, T, U1, H, V, ldv = st["dims"]
    R = B * T * U1
    fc1, fcg, fc2 = model.fc1, model.fc_gate, model.fc2
    dl_parts = [p.view(R, ldv) for p in stage_act(dlogits)]
    dl_v = [p[:, :V] for p in dl_parts]
    dh = _new((R, H), like=dlogits)
    w2b, w2mn = dgrad_b(st["w2"])
    gemm_parts([dl_v], [w2b], dh, b_mn=w2mn)
    gemm_parts([dl_v], [st["h_parts"]], grad_of(fc2.weight), a_mn=True, b_mn=True)
    if db2 is None:
        db2 = torch.empty(ldv, dtype=torch.float32, device=dlogits.device)
        K.colsum(dlogits.view(R, ldv), db2)
    grad_of(fc2.bias).copy_(db2[:V])
    dex = _new((B * T, 2 * H), like=dlogits)
    dpy = _new((B * U1, 2 * H), like=dlogits)
    K.joint_gate_bwd(st["ex"], st["py"], dh, dex, dpy, B, T, U1, H)
    del dh
    dex_parts, dpy_parts = stage_act(dex), stage_act(dpy)
    g1, gg = grad_of(fc1.weight), grad_of(fcg.weight)
    for (dparts, xparts, lo) in ((dex_parts, st["enc_parts"], 0), (dpy_parts, st["pred_parts"], H)):
        gemm_parts([[p[:, :H] for p in dparts]], [xparts], g1[:, lo:lo + H], a_mn=True, b_mn=True)
        gemm_parts([[p[:, H:] for p in dparts]], [xparts], gg[:, lo:lo + H], a_mn=True, b_mn=True)
    dbx = torch.empty(2 * H, dtype=torch.float32, device=dlogits.device)
    K.colsum(dex, dbx)
    grad_of(fc1.bias).copy_(dbx[:H])
    grad_of(fcg.bias).copy_(dbx[H:])
    d_enc = d_pred = None
    if need_enc:
        d_enc = _new((B * T, H), like=dlogits)
        wb, wmn = dgrad_b(st["wx"], cols=(0, H))
        gemm_parts([dex_parts], [wb], d_enc, b_mn=wmn)
        d_enc = d_enc.view(B, T, H)
    if need_pred:
        d_pred = _new((B * U1, H), like=dlogits)
        wb, wmn = dgrad_b(st["wx"], cols=(H, 2 * H))
        gemm_parts([dpy_parts], [wb], d_pred, b_mn=wmn)
        d_pred = d_pred.view(B, U1, H)
    return d_enc, d_pred


class JointFn(torch.autograd.Function):
    """enc [B,T,H], pred [B,U1,H] -> logits [B,T,U1,ldv] (trainer/model/transducer.py:96-108)."""

    @staticmethod
    def forward(ctx, enc, pred, model):
        logits, st = _joint_forward(enc, pred, model)
        ctx.st, ctx.model = st, model
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        d_enc, d_pred = _joint_backward(dlogits.contiguous(), ctx.st, ctx.model, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return d_enc, d_pred, None


_UNIT_LOSS_GRAD = False


def assume_unit_loss_grad(flag):
    """The training step calls ``costs.sum().backward()`` (trainer/train_transducer_bmuf_otfaug.py:97-103), i.e. the upstream
    gradient of every cost is exactly 1.  TrainStep declares that here so JointLossFn.backward does no re-scaling work;
    any other caller gets the general (scaled) backward."""
    global _UNIT_LOSS_GRAD
    _UNIT_LOSS_GRAD = bool(flag)


class JointLossFn(torch.autograd.Function):
    """Fused joint + log-softmax + RNN-T loss: the logits tensor is produced, consumed by the loss
    kernels and overwritten IN PLACE by d(loss)/d(logits); the joint backward then runs immediately, so
    only one [B,T,U1,V] tensor ever exists.  Returns costs [B].

    Gradient contract (one place): every Function of this engine OVERWRITES the ``.grad`` of the parameters it owns
    (EmbeddingFn zero-fills, then scatters), so one forward + one backward per ``zero_grad`` is the supported pattern, as in
    the reference loop (:74, :103).  The joint's parameter gradients are produced here in forward for an upstream gradient
    of 1 per utterance; ``backward`` re-scales them (and d_enc / d_pred) when the upstream gradient is a different UNIFORM
    scalar, and scales d_enc / d_pred per utterance otherwise -- per-utterance weights on the joint's OWN parameters are not
    representable after the fact and raise.  With grad mode off (``need_grad`` False: the eval branch of run_one_epoch) only the
    costs are computed and no gradient buffer is touched."""

    @staticmethod
    def forward(ctx, enc, pred, model, labels, frame_lens, label_lens, need_grad=True):
        logits, st = _joint_forward(enc, pred, model, want_lse=True)
        V = st["dims"][4]
        ctx.need_grad = need_grad
        if not need_grad:
            costs, _ = K.rnnt_loss_fwd_bwd(logits, labels, frame_lens, label_lens, V=V, want_grad=False, row_lse=st.pop("row_lse"))
            return costs
        db2 = torch.empty(logits.shape[-1], dtype=torch.float32, device=logits.device)
        with _Tap("rnnt_loss"):
            costs, _ = K.rnnt_loss_fwd_bwd(logits, labels, frame_lens, label_lens, V=V, dlogits=logits, colsum=db2,
                                           row_lse=st.pop("row_lse"))
        d_enc, d_pred = _joint_backward(logits, st, model, db2=db2)
        del logits, st
        ctx.save_for_backward(d_enc, d_pred)
        ctx.model = model
        return costs

    @staticmethod
    def backward(ctx, dcosts):
        if not ctx.need_grad:
            raise RuntimeError("JointLossFn: forward ran with grad mode off; there is nothing to back-propagate")
        d_enc, d_pred = ctx.saved_tensors
        if not _UNIT_LOSS_GRAD:
            w = dcosts.detach().float()
            if not bool((w == w[0]).all()):
                raise RuntimeError("JointLossFn: per-utterance loss weights are not supported by the fused joint backward "
                                   "(its parameter gradients were formed for a uniform upstream gradient)")
            if float(w[0]) != 1.0:
                m = ctx.model
                for p in (m.fc1.weight, m.fc1.bias, m.fc_gate.weight, m.fc_gate.bias, m.fc2.weight, m.fc2.bias):
                    p.grad.mul_(w[0])
                d_enc = d_enc * w[0].to(d_enc.dtype)
                d_pred = d_pred * w[0].to(d_pred.dtype)
        return d_enc, d_pred, None, None, None, None, None


class RNNTLossFn(torch.autograd.Function):
    """warp_rnnt.RNNTLoss.apply-compatible: log_probs [B,T,U1,V] -> costs [B] (reference call site
    trainer/train_transducer_bmuf_otfaug.py:58,97-98)."""

    @staticmethod
    def forward(ctx, log_probs, labels, frame_lens, label_lens):
        lp = log_probs.contiguous()
        B, T, U1, V = lp.shape
        ldv = V if (V % (8 if lp.dtype == torch.bfloat16 else 4) == 0) else None
        if ldv is None:
            ldv = _ldv(V)
            buf = torch.zeros(B, T, U1, ldv, dtype=lp.dtype, device=lp.device)
            buf[..., :V].copy_(lp)
            lp = buf
        costs, grads = K.rnnt_loss_fwd_bwd(lp, labels.int().contiguous(), frame_lens.int().contiguous(),
                                           label_lens.int().contiguous(), V=V)
        ctx.save_for_backward(grads[..., :V])
        return costs

    @staticmethod
    def backward(ctx, dcosts):
        (g,) = ctx.saved_tensors
        return g * dcosts.view(-1, 1, 1, 1).to(g.dtype), None, None, None


# ------------------------------------------------------------------------------------------------
# model assembly
def _to_act(x):
    x = x.contiguous()
    if x.dtype == act_dtype():
        return x
    if act_dtype() == torch.bfloat16:
        x2 = x.reshape(-1, x.shape[-1]).float()
        out = torch.empty(x2.shape, dtype=torch.bfloat16, device=x.device)
        K.cast_split(x2, out)
        return out.view(x.shape)
    return x.float()


def transformer_layer(layer, x2, B, T, training, causal=False, key_pad=None):
    """x2 [B*T, D] -> [B*T, D]   (pre-LN attention block + position-wise FFN); ``causal`` / ``key_pad``: see AttentionFn."""
    att, ff = layer.self_attn, layer.feed_forward
    p = _drop(layer.dropout_p, training)
    ln = LayerNormFn.apply(x2, layer.layer_norm, layer.layer_norm.weight, layer.layer_norm.bias)
    qkv = linear(ln, [att.linear_query.weight, att.linear_keys.weight, att.linear_values.weight],
                 [att.linear_query.bias, att.linear_keys.bias, att.linear_values.bias])
    ctxv = AttentionFn.apply(qkv.view(B, T, -1), att.head_count, p, _next_seed() if p > 0 else 0, causal, key_pad)
    h1 = linear(ctxv.view(B * T, -1), att.final_linear.weight, att.final_linear.bias, drop_p=p, residual=x2)
    ln2 = LayerNormFn.apply(h1, ff.layer_norm, ff.layer_norm.weight, ff.layer_norm.bias)
    fold = _FOLD_MASKS and ln2.dtype == torch.bfloat16
    inter = linear(ln2, ff.w_1.weight, ff.w_1.bias, act=True, drop_p=p, premasked=fold)
    # w_2's dgrad epilogue applies w_1's ReLU(+dropout) mask: inter != 0 <=> active and kept
    return linear(inter, ff.w_2.weight, ff.w_2.bias, drop_p=p, residual=h1,
                  mask_input_scale=(1.0 / (1.0 - p) if p > 0 else 1.0) if fold else 0.0)


def encoder_forward_act(enc, x):
    """x [B,T,D] -> [B,T',H] in the activation dtype."""
    training = enc.training
    B, T, D = x.shape
    C = enc.tdnn_nhid
    if T < 43:
        raise ValueError("encoder input has %d frames; the TDNN stack needs at least 43 (receptive field 21+1+21)" % T)
    fold = _FOLD_MASKS            # ReLU masks ride in the BatchNorm backward (pk_bn_bwd relu_mask) instead of separate passes
    h = linear(_to_act(x).view(B * T, D), enc.fc_in.weight, enc.fc_in.bias, act=True, premasked=fold)
    h = BatchNormFn.apply(h, enc.bn_in, training, enc.bn_in.weight, enc.bn_in.bias, fold)
    for l, (conv, bn) in enumerate(zip(enc.hidden_conv, enc.hidden_bn)):
        dil, stride = enc.TDNN_DIL_STRIDE[l]
        h3 = TdnnFn.apply(h.view(B, T, C), conv.weight, conv.bias, dil, stride, fold)
        T = h3.shape[1]
        h = BatchNormFn.apply(h3.view(B * T, C), bn, training, bn.weight, bn.bias, fold)
        if (l + 1) % 3 == 0:
            h = transformer_layer(enc.transformer[l // 3], h, B, T, training)
    h = BatchNormFn.apply(h, enc.bn_final, training, enc.bn_final.weight, enc.bn_final.bias)
    h = linear(h, enc.fc_out.weight, enc.fc_out.bias)
    return h.view(B, T, -1)


def encoder_forward(enc, x):
    return encoder_forward_act(enc, x).float()


def prednet_forward_act(model, y):
    """y [B,U] int64 -> [B,U+1,H]: SOS(blank=0) prepend, embedding, LSTM stack with inter-layer dropout
    (trainer/model/transducer.py:90-95)."""
    B, U = y.shape
    sos = torch.zeros(B, 1, dtype=torch.long, device=y.device)
    yy = torch.cat((sos, y.long()), dim=1).contiguous()
    if getattr(model, "decoder_type", "rnn") != "rnn":
        return conv_transformer_lm_forward_act(model.decoder, yy)                  # trainer/model/transducer.py:96-97
    E = model.embed.weight.shape[1]
    ld = (E + 7) // 8 * 8
    h = EmbeddingFn.apply(yy, model.embed, ld, model.embed.weight).view(B, U + 1, ld)
    lstm = model.decoder
    p = _drop(lstm.dropout, lstm.training)
    for l in range(lstm.num_layers):
        h = LstmLayerFn.apply(h, lstm, l, getattr(lstm, "weight_ih_l%d" % l), getattr(lstm, "weight_hh_l%d" % l),
                              getattr(lstm, "bias_ih_l%d" % l), getattr(lstm, "bias_hh_l%d" % l))
        if p > 0 and l < lstm.num_layers - 1:
            h = DropoutFn.apply(h, p, _next_seed())
    return h


def conv_transformer_lm_forward_act(dec, src):
    """Transformer prediction net (trainer/model/rnnt_conv_transformer_lm.py:59-80): src [B,L] int64 (SOS already prepended) ->
    [B,L,output_dim].  Embedding -> num_layers x (causal Conv1d(k=5) + ReLU -> pre-LN transformer layer under the causal +
    padding-key mask) -> LayerNorm -> linear_out."""
    if getattr(dec, "max_relative_positions", 0) > 0:
        raise NotImplementedError("pika_b200: relative position embeddings (max_relative_positions > 0) are not on the hot path")
    B, L = src.shape
    emb = dec.embeddings
    ld = (emb.weight.shape[1] + 7) // 8 * 8
    src = src.contiguous()
    h = EmbeddingFn.apply(src, emb, ld, emb.weight).view(B, L, ld)
    key_pad = src.eq(emb.padding_idx).to(torch.uint8).contiguous() if emb.padding_idx is not None else None     # :66-68
    training = dec.training
    for conv, layer in zip(dec.conv, dec.transformer):
        assert conv.kernel_size[0] - 1 == conv.padding[0] and conv.dilation[0] == 1 and conv.stride[0] == 1
        h = CausalConvFn.apply(h, conv.weight, conv.bias)                          # :74-75
        h = transformer_layer(layer, h.view(B * L, -1), B, L, training, causal=True, key_pad=key_pad).view(B, L, -1)
    hn = LayerNormFn.apply(h.reshape(B * L, -1), dec.layer_norm, dec.layer_norm.weight, dec.layer_norm.bias)
    return linear(hn, dec.linear_out.weight, dec.linear_out.bias).view(B, L, -1)


def transducer_forward(model, x, y, softmax=True):
    """Net.forward drop-in: -> [B,T',U+1,V] fp32 log-probs (softmax=True) or logits."""
    enc = encoder_forward_act(model.encoder, x)
    pred = prednet_forward_act(model, y)
    logits = JointFn.apply(enc, pred, model)
    V = model.fc2.weight.shape[0]
    if not softmax:
        return logits[..., :V].float()
    return LogSoftmaxFn.apply(logits, V)


class LogSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, V):
        out = torch.empty(logits.shape[:-1] + (V,), dtype=torch.float32, device=logits.device)
        K.log_softmax(logits, out, V)
        ctx.save_for_backward(out)
        ctx.ldv, ctx.dt = logits.shape[-1], logits.dtype
        return out

    @staticmethod
    def backward(ctx, dlp):
        # d logits = dlp - softmax * sum(dlp); boundary op of the compatibility API only (the training
        # path uses JointLossFn, which never materialises log-probs)
        (lp,) = ctx.saved_tensors
        g = dlp - lp.exp() * dlp.sum(-1, keepdim=True)
        out = torch.zeros(lp.shape[:-1] + (ctx.ldv,), dtype=ctx.dt, device=lp.device)
        out[..., :lp.shape[-1]] = g.to(ctx.dt)
        return out, None


def transducer_loss(model, x, y, frame_lens, label_lens):
    """Fused training path: costs [B] with gradients wired to every parameter."""
    enc = encoder_forward_act(model.encoder, x)
    pred = prednet_forward_act(model, y)
    return JointLossFn.apply(enc, pred, model, y.int().contiguous(), frame_lens.int().contiguous(), label_lens.int().contiguous(),
                             torch.is_grad_enabled())
