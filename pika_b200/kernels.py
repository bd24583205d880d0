"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers + sizes across the boundary.

torch is used here only for device memory, streams and shape bookkeeping.
"""
import ctypes

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, AUX_ADD, AUX_MASK_NZ, AUX_NONE, PK_BF16, PK_F32, SEL_KZ, SEL_ZB0, SEL_ZB1,  # noqa: F401  (re-exported: engine uses K.ACT_RELU ...)
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
         aux=None, aux_mode=AUX_NONE, aux_scale=1.0, accumulate=False, block_n=0, k_splits=0, two_sm=0, row_lse=None):
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
    d.k_splits = k_splits
    d.two_sm = two_sm
    if row_lse is not None:
        assert row_lse.dtype == torch.float32 and row_lse.is_contiguous() and c.dim() == 2
        assert tuple(row_lse.shape) == (row_lse_parts(c.shape[-2], c.shape[-1], block_n, two_sm), c.shape[-2], 2)
        d.row_lse = row_lse.data_ptr()
    check(lib.pk_gemm_bf16(ctypes.byref(d), _stream()), "pk_gemm_bf16")
    return c


def row_lse_parts(M, N, block_n=0, two_sm=0):
    """number of per-row (max, sum-exp) partials the GEMM writes into ``row_lse`` for an [M, N] output: one per 256-wide
    N tile, two when the CTA-pair kernel (two epilogue groups per tile) runs it"""
    return int(lib.pk_gemm_row_lse_parts(M, N, block_n, two_sm))


def rnnt_loss_fwd_bwd(logits, labels, frame_lens, label_lens, V=None, grad_scale=None, dlogits=None, want_grad=True, colsum=None,
                      row_lse=None):
    """logits [B,T,U1,ldv] (bf16|f32) -> (costs [B] f32, dlogits).  dlogits may alias logits.
    row_lse [n_parts, B*T*U1, 2]: per-row log-sum-exp partials written by the producing GEMM (skips the first pass)."""
    B, T, U1, ldv = logits.shape
    V = ldv if V is None else V
    assert logits.is_contiguous() and labels.dtype == torch.int32 and labels.dim() == 2
    assert frame_lens.dtype == torch.int32 and label_lens.dtype == torch.int32
    ws_bytes = int(lib.pk_rnnt_loss_workspace_bytes(B, T, U1))
    if colsum is not None:
        ws_bytes += int(lib.pk_rnnt_loss_colsum_workspace_bytes(B, T, U1, ldv))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
    costs = torch.empty(B, dtype=torch.float32, device=logits.device)
    if want_grad and dlogits is None:
        dlogits = torch.empty_like(logits)
    if row_lse is not None:
        assert row_lse.dtype == torch.float32 and row_lse.is_contiguous() and tuple(row_lse.shape[1:]) == (B * T * U1, 2)
        check(lib.pk_rnnt_loss_fwd_bwd_lse(_ptr(logits), _dt(logits), _ptr(labels), _ptr(frame_lens), _ptr(label_lens),
                                           B, T, U1, V, ldv, max(labels.stride(0), 1), _ptr(grad_scale), _ptr(costs),
                                           _ptr(dlogits if want_grad else None), _ptr(colsum), _ptr(ws), ws_bytes,
                                           _ptr(row_lse), int(row_lse.shape[0]), _stream()), "pk_rnnt_loss_fwd_bwd_lse")
        return costs, dlogits
    check(lib.pk_rnnt_loss_fwd_bwd(_ptr(logits), _dt(logits), _ptr(labels), _ptr(frame_lens), _ptr(label_lens),
                                   B, T, U1, V, ldv, max(labels.stride(0), 1), _ptr(grad_scale), _ptr(costs),
                                   _ptr(dlogits if want_grad else None), _ptr(colsum), _ptr(ws), ws_bytes, _stream()),
          "pk_rnnt_loss_fwd_bwd")
    return costs, dlogits


# ------------------------------------------------------------------------------------------------
# thin wrappers for the memory-bound kernels (argument marshalling only)
_P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
_L = ctypes.c_longlong
_I = ctypes.c_int
_F = ctypes.c_float
_U = ctypes.c_uint32


def cast_split(src, hi, lo=None, cols_pad=None, scale=1.0):
    """src [rows, cols] (f32|bf16, row stride free) -> hi (and lo) bf16 [rows, cols_pad]."""
    rows, cols = src.shape
    cols_pad = cols_pad or cols
    assert hi.shape[-1] == cols_pad and hi.stride(-1) == 1 and src.stride(-1) == 1
    check(lib.pk_cast_split(_P(src), _I(_dt(src)), _L(src.stride(0)), _P(hi), _P(lo), _L(hi.stride(0)), _L(rows),
                            _I(cols), _I(cols_pad), _F(scale), _stream()), "pk_cast_split")


def transpose_bf16(src, dst):
    """dst [cols, rows] = src [rows, cols]^T (bf16, row strides free)"""
    rows, cols = src.shape
    assert src.dtype == torch.bfloat16 and dst.dtype == torch.bfloat16 and dst.shape == (cols, rows)
    assert src.stride(1) == 1 and dst.stride(1) == 1
    check(lib.pk_transpose_bf16(_P(src), _L(src.stride(0)), _P(dst), _L(dst.stride(0)), _I(rows), _I(cols), _stream()),
          "pk_transpose_bf16")


def attention_lse_stride(T):
    """row pitch of the per-(batch, head) lse / D vectors (64-element aligned so that tiles of them are bulk-copy sources)"""
    return int(lib.pk_attention_lse_stride(T))


def attention_fwd(qkv, out, lse, heads, alpha, drop_p=0.0, seed=0):
    """qkv [B,T,3D] bf16 (q | k | v column blocks) -> out [B,T,D], lse [B*heads*T] f32 (fused attention, head dim 64)"""
    B, T, D3 = qkv.shape
    D = D3 // 3
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and out.is_contiguous() and out.shape == (B, T, D)
    assert lse.dtype == torch.float32 and lse.numel() == B * heads * attention_lse_stride(T)
    base, es = qkv.data_ptr(), 2
    vp = ctypes.c_void_p
    check(lib.pk_attention_fwd(vp(base), vp(base + D * es), vp(base + 2 * D * es), _L(D3), _P(out), _L(D), _P(lse), _I(B), _I(T),
                               _I(heads), _I(D // heads), _F(alpha), _F(drop_p), _U(seed & 0xFFFFFFFF), _stream()), "pk_attention_fwd")


def attention_bwd(qkv, out, dout, lse, dqkv, heads, alpha, drop_p=0.0, seed=0):
    B, T, D3 = qkv.shape
    D = D3 // 3
    assert dout.is_contiguous() and dqkv.is_contiguous() and dqkv.shape == qkv.shape and dout.dtype == torch.bfloat16
    ws = torch.zeros(B * heads * attention_lse_stride(T), dtype=torch.float32, device=qkv.device)     # pad entries must be finite (0)
    base, gb, es = qkv.data_ptr(), dqkv.data_ptr(), 2
    vp = ctypes.c_void_p
    check(lib.pk_attention_bwd(vp(base), vp(base + D * es), vp(base + 2 * D * es), _L(D3), _P(out), _L(D), _P(dout), _L(D), _P(lse), _P(ws),
                               vp(gb), vp(gb + D * es), vp(gb + 2 * D * es), _L(D3), _I(B), _I(T), _I(heads), _I(D // heads), _F(alpha),
                               _F(drop_p), _U(seed & 0xFFFFFFFF), _stream()), "pk_attention_bwd")


_col_ws = {}


def col_ws(C, device):
    """persistent scratch for the two-stage column reductions (per device, per C)"""
    key = (C, str(device))
    if key not in _col_ws:
        lib.pk_colstats_ws_floats.restype = ctypes.c_longlong
        _col_ws[key] = torch.empty(int(lib.pk_colstats_ws_floats(C)) + 2 * C, dtype=torch.float32, device=device)
    return _col_ws[key]


def bn_fwd(x, y, w, b, eps, train, momentum, run_mean, run_var, mean, rstd, ws=None):
    rows, C = x.shape
    ws = col_ws(C, x.device)
    check(lib.pk_bn_fwd(_P(x), _P(y), _I(_dt(x)), _L(rows), _I(C), _P(w), _P(b), _F(eps), _I(int(train)), _F(momentum),
                        _P(run_mean), _P(run_var), _P(mean), _P(rstd), _P(ws), _stream()), "pk_bn_fwd")


def bn_bwd(dy, x, dx, w, mean, rstd, train, relu_mask, dw, db):
    rows, C = x.shape
    check(lib.pk_bn_bwd(_P(dy), _P(x), _P(dx), _I(_dt(x)), _L(rows), _I(C), _P(w), _P(mean), _P(rstd), _I(int(train)),
                        _I(int(relu_mask)), _P(dw), _P(db), _P(col_ws(C, x.device)), _stream()), "pk_bn_bwd")


def colsum(x, out):
    rows, C = x.shape
    assert x.is_contiguous()
    check(lib.pk_colsum(_P(x), _I(_dt(x)), _L(rows), _I(C), _P(out), _P(col_ws(C, x.device)), _stream()), "pk_colsum")


def layernorm_fwd(x, y, w, b, eps, mean, rstd):
    rows, C = x.shape
    check(lib.pk_layernorm_fwd(_P(x), _P(y), _I(_dt(x)), _L(rows), _I(C), _P(w), _P(b), _F(eps), _P(mean), _P(rstd),
                               _stream()), "pk_layernorm_fwd")


def layernorm_bwd(dy, x, dx, w, mean, rstd, dw, db):
    rows, C = x.shape
    check(lib.pk_layernorm_bwd(_P(dy), _P(x), _P(dx), _I(_dt(x)), _L(rows), _I(C), _P(w), _P(mean), _P(rstd), _P(dw),
                               _P(db), _stream()), "pk_layernorm_bwd")


def softmax_fwd(S, P, Pd, n, drop_p, seed):
    rows = S.numel() // S.shape[-1]
    check(lib.pk_softmax_fwd(_P(S), _L(S.shape[-1]), _P(P), _P(Pd), _I(_dt(P)), _L(P.shape[-1]), _L(rows), _I(n),
                             _F(drop_p), _U(seed & 0xFFFFFFFF), _stream()), "pk_softmax_fwd")


def softmax_masked_fwd(S, P, Pd, n, q_len, heads, causal, key_pad, drop_p, seed):
    """rows of S = (sequence, head, query); key c of query i is dropped when c > i (causal) or key_pad[sequence, c] != 0"""
    rows = S.numel() // S.shape[-1]
    if key_pad is not None:
        assert key_pad.dtype == torch.uint8 and key_pad.is_contiguous() and key_pad.shape[-1] == n
    check(lib.pk_softmax_masked_fwd(_P(S), _L(S.shape[-1]), _P(P), _P(Pd), _I(_dt(P)), _L(P.shape[-1]), _L(rows), _I(n), _I(q_len),
                                    _I(heads), _I(int(bool(causal))), _P(key_pad), _F(drop_p), _U(seed & 0xFFFFFFFF), _stream()),
          "pk_softmax_masked_fwd")


def softmax_bwd(dPd, P, dS, n, drop_p, seed):
    rows = P.numel() // P.shape[-1]
    check(lib.pk_softmax_bwd(_P(dPd), _L(dPd.shape[-1]), _P(P), _L(P.shape[-1]), _P(dS), _I(_dt(P)), _L(rows), _I(n),
                             _F(drop_p), _U(seed & 0xFFFFFFFF), _stream()), "pk_softmax_bwd")


def dropout(x, y, p, seed):
    check(lib.pk_dropout(_P(x), _P(y), _I(_dt(x)), _L(x.numel()), _F(p), _U(seed & 0xFFFFFFFF), _stream()), "pk_dropout")


def mask_nz(dy, y, dx, scale):
    check(lib.pk_mask_nz(_P(dy), _P(y), _P(dx), _I(_dt(y)), _L(y.numel()), _F(scale), _stream()), "pk_mask_nz")


def add(a, b, o):
    check(lib.pk_add(_P(a), _P(b), _P(o), _I(_dt(a)), _L(a.numel()), _stream()), "pk_add")


def log_softmax(x, y, n, scale=1.0):
    rows = x.numel() // x.shape[-1]
    check(lib.pk_log_softmax(_P(x), _I(_dt(x)), _L(x.shape[-1]), _P(y), _L(rows), _I(n), _F(scale), _stream()),
          "pk_log_softmax")


def joint_gate_fwd(ex, py, h, B, T, U1, H):
    check(lib.pk_joint_gate_fwd(_P(ex), _P(py), _P(h), _I(_dt(ex)), _I(B), _I(T), _I(U1), _I(H), _I(h.stride(0)), _stream()),
          "pk_joint_gate_fwd")


def joint_gate_bwd(ex, py, dh, dex, dpy, B, T, U1, H):
    check(lib.pk_joint_gate_bwd(_P(ex), _P(py), _P(dh), _P(dex), _P(dpy), _I(_dt(ex)), _I(B), _I(T), _I(U1), _I(H),
                                _stream()), "pk_joint_gate_bwd")


def lstm_cell_fwd(gx, gh, c_prev, c_out, h_out, gates_save, B, H):
    check(lib.pk_lstm_cell_fwd(_P(gx), _L(gx.stride(0)), _P(gh), _L(gh.stride(0) if gh is not None else 0), _P(c_prev),
                               _P(c_out), _P(h_out), _I(_dt(h_out)), _L(h_out.stride(0)), _P(gates_save), _I(B), _I(H),
                               _stream()), "pk_lstm_cell_fwd")


def lstm_cell_bwd(dh_out, dh_rec, dc_next, gates, c, c_prev, dgates, dc_prev, B, H):
    check(lib.pk_lstm_cell_bwd(_P(dh_out), _L(dh_out.stride(0) if dh_out is not None else 0), _P(dh_rec), _P(dc_next),
                               _P(gates), _P(c), _P(c_prev), _P(dgates), _I(_dt(dgates)), _P(dc_prev), _I(B), _I(H),
                               _stream()), "pk_lstm_cell_bwd")


def embedding_fwd(idx, table, out):
    n, ld = out.shape
    check(lib.pk_embedding_fwd(_P(idx), _P(table), _I(table.shape[1]), _P(out), _I(_dt(out)), _I(ld), _L(n), _stream()),
          "pk_embedding_fwd")


def embedding_bwd(idx, dout, dtable, padding_idx):
    n, ld = dout.shape
    check(lib.pk_embedding_bwd(_P(idx), _P(dout), _I(_dt(dout)), _I(ld), _I(dtable.shape[1]), _P(dtable), _L(n),
                               _L(padding_idx if padding_idx is not None else -1), _stream()), "pk_embedding_bwd")


def absmax(x, out, nan_flag=None):
    check(lib.pk_absmax(_P(x), _L(x.numel()), _P(out), _P(nan_flag), _stream()), "pk_absmax")


def sgd_nesterov_clip(p, g, buf, lr, momentum, max_norm, absmax_t, first, nan_flag=None):
    check(lib.pk_sgd_nesterov_clip(_P(p), _P(g), _P(buf), _L(p.numel()), _F(lr), _F(momentum), _F(max_norm), _P(absmax_t),
                                   _P(nan_flag), _I(int(first)), _stream()), "pk_sgd_nesterov_clip")


def bmuf_delta(glob, local, delta):
    check(lib.pk_bmuf_delta(_P(glob), _P(local), _P(delta), _L(glob.numel()), _stream()), "pk_bmuf_delta")


def bmuf_update(glob, local, delta_prev, delta_sum, world, bm, blr):
    check(lib.pk_bmuf_update(_P(glob), _P(local), _P(delta_prev), _P(delta_sum), _L(glob.numel()), _I(world), _F(bm), _F(blr),
                             _stream()), "pk_bmuf_update")


_lstm_ws = {}


def _lstm_scratch(H, device):
    key = (H, str(device))
    if key not in _lstm_ws:
        lib.pk_lstm_seq_workspace_bytes.restype = ctypes.c_longlong
        _lstm_ws[key] = torch.zeros(int(lib.pk_lstm_seq_workspace_bytes(H)), dtype=torch.uint8, device=device)
    return _lstm_ws[key]


def lstm_seq_fwd(gx, w_hh, out, gates_save, cs):
    B, U, G4 = gx.shape
    H = G4 // 4
    assert gx.is_contiguous() and out.is_contiguous() and w_hh.dtype == torch.bfloat16 and w_hh.is_contiguous()
    check(lib.pk_lstm_seq_fwd(_P(gx), _P(w_hh), _P(out), _I(_dt(out)), _P(gates_save), _P(cs), _I(B), _I(U), _I(H),
                              _P(_lstm_scratch(H, gx.device)), _stream()), "pk_lstm_seq_fwd")


def lstm_seq_bwd(dout, gates_save, cs, w_hh, dG):
    B, U, H = dout.shape
    assert dout.is_contiguous() and dG.dtype == torch.bfloat16 and dG.is_contiguous()
    check(lib.pk_lstm_seq_bwd(_P(dout), _I(_dt(dout)), _P(gates_save), _P(cs), _P(w_hh), _P(dG), _I(B), _I(U), _I(H),
                              _P(_lstm_scratch(H, dout.device)), _stream()), "pk_lstm_seq_bwd")


def gather_rows(src, idx, dst):
    rows, C = dst.shape
    check(lib.pk_gather_rows(_P(src), _P(idx), _P(dst), _I(_dt(src)), _L(rows), _I(C), _stream()), "pk_gather_rows")


def scatter_add_rows(src, idx, dst):
    rows, C = src.shape
    assert dst.dtype == torch.float32
    check(lib.pk_scatter_add_rows(_P(src), _P(idx), _P(dst), _I(_dt(src)), _L(rows), _I(C), _stream()), "pk_scatter_add_rows")


def ce_grad(z, tok, coef, scale, dz, n):
    rows, ld = z.shape
    check(lib.pk_ce_grad(_P(z), _I(_dt(z)), _L(ld), _P(tok), _P(coef), _F(scale), _P(dz), _L(rows), _I(n), _stream()), "pk_ce_grad")
