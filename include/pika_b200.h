/* pika_b200 -- C ABI of the B200-native RNN-Transducer hot path.
 *
 * This is the drop-in boundary.  The reference (tencent-ailab/pika) has
 * no FFI of its own: its hot path is Python calling torch / warp_rnnt / PyKaldi.  Each entry
 * point below names the reference call it replaces (paths relative to the reference root).
 * A maintainer binds these with ctypes; pika_b200/_lib.py is that binding.
 *
 * Conventions: plain pointers and sizes only (no torch types); every function returns 0 on
 * success, <0 on error (pk_last_error() gives the message); device pointers unless noted;
 * functions never allocate device memory, never synchronise the stream, and are re-entrant per
 * stream.  `stream` is a cudaStream_t passed as void*.  dtype codes: 0 = float32, 1 = bfloat16.
 */
#ifndef PIKA_B200_H
#define PIKA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_F32 0
#define PK_BF16 1

const char* pk_last_error(void);
int pk_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
long long pk_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense contractions on the 5th-gen tensor cores (tcgen05 + TMEM, TMA-fed, bf16 x bf16 -> fp32).
 * Replaces every nn.Linear / nn.Conv2d(TDNN) / torch.matmul on the path:
 *   trainer/model/rnnt_tdnn_transformer.py:44-59,76-89 (fc_in, 9 TDNN convs, fc_out)
 *   trainer/model/modules/multi_headed_attn.py:180-182,207,223,233 (QKV, QK^T, PV, final_linear)
 *   trainer/model/modules/position_ffn.py:37-38 (w_1, w_2)
 *   trainer/model/transducer.py:56-61 (LSTM input/recurrent products), :108 (fc1, fc_gate, fc2)
 * and their autograd backward (dgrad / wgrad).
 *
 *   C[z][m][n] = epilogue( alpha * sum_p sum_kz sum_k A_p[z|kz][m + a_row_off_p][k] * B_p[z|kz][n + b_row_off_p][k] )
 *
 * Every operand is a 4-D strided view; dim[0] is the contiguous one.
 *   A, K-major : dim = (K, M, z2, z3)      A, MN-major: dim = (M, K, z2, z3)
 *   B, K-major : dim = (K, N, z2, z3)      B, MN-major: dim = (N, K, z2, z3)
 *   C          : dim = (N, M, zb0, zb1)
 * coordinates 2/3 of A and B are taken from (0 | zb0 | zb1 | kz) as chosen by *_sel.
 * Up to PK_GEMM_MAX_PAIRS (A_p, B_p) pairs accumulate into one tile: the 3 taps of a TDNN
 * layer (no im2col), and/or the 3 partial products of the split-bf16 "fp32-class" mode.
 */
#define PK_GEMM_MAX_PAIRS 9
#define PK_SEL_ZERO 0
#define PK_SEL_ZB0 1
#define PK_SEL_ZB1 2
#define PK_SEL_KZ 3
#define PK_ACT_NONE 0
#define PK_ACT_RELU 1
#define PK_AUX_NONE 0
#define PK_AUX_ADD 1      /* out += aux[m][n]                         (residual)            */
#define PK_AUX_MASK_NZ 2  /* out  = aux[m][n] != 0 ? out*aux_scale : 0 (ReLU/dropout backward) */

typedef struct {
    const void* ptr;    /* bf16 for A/B; bf16 or f32 for C */
    int64_t dim[4];     /* extents, dim[0] contiguous */
    int64_t stride[3];  /* strides of dims 1..3, in elements (multiples of 8 for bf16, 4 for f32) */
} pk_view4;

typedef struct {
    int n_pairs;
    pk_view4 a[PK_GEMM_MAX_PAIRS];
    pk_view4 b[PK_GEMM_MAX_PAIRS];
    int a_row_off[PK_GEMM_MAX_PAIRS]; /* added to the M (K-major) or K (MN-major) coordinate; may be negative */
    int b_row_off[PK_GEMM_MAX_PAIRS];
    int a_mn_major, b_mn_major;       /* 0 = K-major (reduction dim contiguous), 1 = MN-major */
    int a_sel2, a_sel3, b_sel2, b_sel3;
    int kz_count;                     /* extra (batched) reduction loop, >= 1 */
    pk_view4 c;
    int c_dtype;                      /* PK_F32 | PK_BF16 */
    int c_accumulate;                 /* 1: C += result (TMA reduce-add, f32 only) */
    /* epilogue, applied in this order */
    float alpha;
    const float* bias;                /* [N] or NULL */
    int act;
    float drop_p;                     /* 0 = off; keep-scale 1/(1-p) */
    uint32_t drop_seed;
    int aux_mode;
    const void* aux;                  /* same logical shape as C */
    int aux_dtype;
    int64_t aux_stride[3];            /* strides of (m, zb0, zb1) in elements */
    float aux_scale;
    int block_n;                      /* 0 = auto; else 64 | 128 | 256 */
} pk_gemm_desc;

int pk_gemm_bf16(const pk_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------
 * RNN-T loss + gradient, fused with the log-softmax over V.
 * Replaces F.log_softmax (trainer/model/transducer.py:110-111) + warp_rnnt RNNTLoss.apply
 * (trainer/train_transducer_bmuf_otfaug.py:58,97-99; trainer/train_transducer_mbr_bmuf_otfaug.py:157-160)
 * and autograd's backward through both.
 *   logits  [B, T, U1, ldv] (first V of each row valid), dtype f32 | bf16, blank = 0
 *   labels  [B, ld_labels] int32; frame_lens, label_lens [B] int32 (T_n <= T, U_n <= U1-1)
 *   grad_scale [B] or NULL (dLoss/dcost_n, 1 when NULL)
 *   costs   [B] f32 = -log P(y_n | x_n)
 *   dlogits same shape/dtype as logits, may alias it (in place); NULL = loss only.
 *           Entries of padded nodes and of the row padding [V, ldv) are written as 0.
 */
long long pk_rnnt_loss_workspace_bytes(int B, int T, int U1);
int pk_rnnt_loss_fwd_bwd(const void* logits, int dtype, const int* labels, const int* frame_lens,
                         const int* label_lens, int B, int T, int U1, int V, int ldv, int ld_labels,
                         const float* grad_scale, float* costs, void* dlogits, void* workspace,
                         long long workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
