"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers + sizes across the boundary.

torch is used here only for device memory, streams and shape bookkeeping.
"""
import ctypes

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, AUX_ADD, AUX_MASK_NZ, AUX_NONE, PK_BF16, PK_F32, SEL_KZ, SEL_ZB0, SEL_ZB1,
                   SEL_ZERO, GemmDesc, View4, check, lib)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return PK_BF16
    if t.dtype == torch.float32:
        return PK_F32
    raise TypeError("unsupported dtype %s" % t.dtype)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _fill_view(v, t):
    """torch view with <= 4 dims, last dim contiguous -> View4 (dim[0] = contiguous extent)."""
    assert t.is_cuda and t.dim() >= 1 and t.dim() <= 4, "views must be CUDA tensors with 1..4 dims"
    assert t.stride(-1) == 1 or t.shape[-1] == 1, "innermost dimension must be contiguous"
    shape = list(t.shape)[::-1]
    strides = list(t.stride())[::-1]
    v.ptr = t.data_ptr()
    for i in range(4):
        v.dim[i] = shape[i] if i < len(shape) else 1
    for i in range(3):
        v.stride[i] = strides[i + 1] if i + 1 < len(strides) else 0
    return v


def gemm(a, b, c, a_mn=False, b_mn=False, a_sel=(SEL_ZB0, SEL_ZB1), b_sel=(SEL_ZB0, SEL_ZB1), kz_count=1,
         a_row_off=None, b_row_off=None, alpha=1.0, bias=None, act=ACT_NONE, drop_p=0.0, drop_seed=0,
         aux=None, aux_mode=AUX_NONE, aux_scale=1.0, accumulate=False, block_n=0):
    """C = epilogue(alpha * sum_p A_p @ B_p^T) on the tcgen05 tensor cores (include/pika_b200.h).

    a, b: a bf16 view or a list of views (pairs).  Views are torch tensors of <= 4 dims laid out
    (z3, z2, rows, contiguous):  K-major A = (.., M, K), MN-major A = (.., K, M); same for B with N.
    c: (zb1, zb0, M, N) view, bf16 or f32.  aux: tensor broadcast-compatible with c's logical shape,
    given as a view with the same number of dims as c.
    """
    a = a if isinstance(a, (list, tuple)) else [a]
    b = b if isinstance(b, (list, tuple)) else [b]
    assert len(a) == len(b) and 1 <= len(a) <= _lib.MAX_PAIRS
    d = GemmDesc()
    d.n_pairs = len(a)
    for i, (ai, bi) in enumerate(zip(a, b)):
        assert ai.dtype == torch.bfloat16 and bi.dtype == torch.bfloat16, "GEMM operands must be bf16"
        _fill_view(d.a[i], ai)
        _fill_view(d.b[i], bi)
        d.a_row_off[i] = a_row_off[i] if a_row_off else 0
        d.b_row_off[i] = b_row_off[i] if b_row_off else 0
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    # selectors only matter for dims that exist (extent > 1); default maps z2<-zb0, z3<-zb1
    d.a_sel2, d.a_sel3 = a_sel
    d.b_sel2, d.b_sel3 = b_sel
    d.kz_count = kz_count
    _fill_view(d.c, c)
    d.c_dtype = _dt(c)
    d.c_accumulate = int(accumulate)
    d.alpha = alpha
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        d.bias = bias.data_ptr()
    d.act = act
    d.drop_p = drop_p
    d.drop_seed = drop_seed & 0xFFFFFFFF
    if aux is not None:
        assert aux.dim() == c.dim() and aux.stride(-1) == 1
        d.aux_mode = aux_mode
        d.aux = aux.data_ptr()
        d.aux_dtype = _dt(aux)
        st = list(aux.stride())[::-1]          # (n, m, zb0, zb1)
        for i in range(3):
            d.aux_stride[i] = st[i + 1] if i + 1 < len(st) else 0
        d.aux_scale = aux_scale
    d.block_n = block_n
    check(lib.pk_gemm_bf16(ctypes.byref(d), _stream()), "pk_gemm_bf16")
    return c


def rnnt_loss_fwd_bwd(logits, labels, frame_lens, label_lens, V=None, grad_scale=None, dlogits=None, want_grad=True):
    """logits [B,T,U1,ldv] (bf16|f32) -> (costs [B] f32, dlogits).  dlogits may alias logits."""
    B, T, U1, ldv = logits.shape
    V = ldv if V is None else V
    assert logits.is_contiguous() and labels.dtype == torch.int32 and labels.dim() == 2
    assert frame_lens.dtype == torch.int32 and label_lens.dtype == torch.int32
    ws_bytes = int(lib.pk_rnnt_loss_workspace_bytes(B, T, U1))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
    costs = torch.empty(B, dtype=torch.float32, device=logits.device)
    if want_grad and dlogits is None:
        dlogits = torch.empty_like(logits)
    check(lib.pk_rnnt_loss_fwd_bwd(_ptr(logits), _dt(logits), _ptr(labels), _ptr(frame_lens), _ptr(label_lens),
                                   B, T, U1, V, ldv, max(labels.stride(0), 1), _ptr(grad_scale), _ptr(costs),
                                   _ptr(dlogits if want_grad else None), _ptr(ws), ws_bytes, _stream()),
          "pk_rnnt_loss_fwd_bwd")
    return costs, dlogits
