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
    int k_splits;                     /* 0 = auto; 1 = off; >1 = split the reduction (plain f32 2-D C only) */
    int two_sm;                       /* 0 = auto (CTA-pair kernel for 256-wide tiles when M > 128); 1 = force; -1 = never */
    float* row_lse;                   /* optional out [pk_gemm_row_lse_parts()][M][2] f32: per row and column group (max*log2(e),
                                         sum_j 2^(c_ij*log2(e) - max)) over the ROUNDED bf16 outputs -- the first pass of the fused
                                         log-softmax + RNN-T loss (pk_rnnt_loss_fwd_bwd_lse) computed while the logits tile is still in TMEM.
                                         Needs a 2-D bf16 C with N % 8 == 0, K-major operands, block_n 256. */
} pk_gemm_desc;

int pk_gemm_bf16(const pk_gemm_desc* desc, void* stream);
/* number of partials per row that pk_gemm_bf16 writes into row_lse for an [M, N] output with these block_n / two_sm settings:
 * one per 256-wide N tile on the single-CTA kernel, two on the CTA-pair kernel (two epilogue groups per tile) */
int pk_gemm_row_lse_parts(long long M, long long N, int block_n, int two_sm);

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
 *   dlogits_colsum [ldv] f32 or NULL: sum of dlogits over all (b,t,u) rows = the joint fc2 bias gradient,
 *           produced by the gradient pass itself (each thread owns fixed columns), so dlogits is not re-read.
 *           Deterministic: per-CTA partial rows are added in a fixed order (no atomics); when requested the workspace
 *           must hold pk_rnnt_loss_workspace_bytes + pk_rnnt_loss_colsum_workspace_bytes bytes.
 */
long long pk_rnnt_loss_workspace_bytes(int B, int T, int U1);
long long pk_rnnt_loss_colsum_workspace_bytes(int B, int T, int U1, int ldv);
int pk_rnnt_loss_fwd_bwd(const void* logits, int dtype, const int* labels, const int* frame_lens,
                         const int* label_lens, int B, int T, int U1, int V, int ldv, int ld_labels,
                         const float* grad_scale, float* costs, void* dlogits, float* dlogits_colsum, void* workspace,
                         long long workspace_bytes, void* stream);
/* Same, with the first pass (row log-sum-exp) already done by the GEMM that produced the logits:
 * row_lse [n_parts][B*T*U1][2] as written through pk_gemm_desc.row_lse.  Saves one full read of the logits. */
int pk_rnnt_loss_fwd_bwd_lse(const void* logits, int dtype, const int* labels, const int* frame_lens,
                             const int* label_lens, int B, int T, int U1, int V, int ldv, int ld_labels,
                             const float* grad_scale, float* costs, void* dlogits, float* dlogits_colsum, void* workspace,
                             long long workspace_bytes, const float* row_lse, int n_parts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Memory-bound layers around the GEMMs (pika_b200/csrc/elementwise.cu).  `dtype` is the
 * activation type (PK_BF16 production, PK_F32 fp32-class parity mode); statistics are f32.
 */
/* bf16 (hi) and optional residual (lo) copies of an f32/bf16 matrix, zero-padded to cols_pad:
 * weight/activation staging for pk_gemm_bf16 (replaces the implicit casts of torch autocast-free fp32). */
int pk_cast_split(const void* src, int src_dtype, long long ld_src, void* hi, void* lo, long long ld_dst,
                  long long rows, int cols, int cols_pad, float scale, void* stream);
/* dst[c, r] = src[r, c] for a bf16 matrix [rows, cols]: K-major copies of staged weights for the dgrad GEMMs
 * (the reference relies on cuBLAS's transposed-operand modes, e.g. nn.Linear backward). */
int pk_transpose_bf16(const void* src, long long ld_src, void* dst, long long ld_dst, int rows, int cols, void* stream);
/* Fused unmasked multi-head self-attention, head dim 64, bf16 (pika_b200/csrc/attention_tc.cu):
 *   O = dropout(softmax(alpha * Q K^T)) V  per (batch, head)   -- MultiHeadedAttention.forward,
 *   trainer/model/modules/multi_headed_attn.py:199-223 (scale, softmax, dropout, context) and its autograd backward,
 * without materialising the [B, heads, T, T] score / probability tensors.
 *   q, k, v: element (b, t, h, d) at ptr[(b*T + t)*ld_qkv + h*64 + d]  (three column blocks of a fused [B,T,3D] projection)
 *   out / dout [B, T, heads*64] with row strides ld_out / ld_dout;  lse [B*heads*T] f32 saved for the backward
 *   dq, dk, dv: same addressing with row stride ld_dqkv;  dsum_ws: B*heads*T floats of scratch
 * Dropout masks are the same counter-based masks as pk_softmax_fwd/bwd for equal (drop_p, seed). */
int pk_attention_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out, float* lse,
                     int B, int T, int heads, int dh, float alpha, float drop_p, uint32_t seed, void* stream);
int pk_attention_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                     const void* dout, long long ld_dout, const float* lse, float* dsum_ws, void* dq, void* dk, void* dv,
                     long long ld_dqkv, int B, int T, int heads, int dh, float alpha, float drop_p, uint32_t seed, void* stream);
/* nn.BatchNorm1d over rows [rows, C] (trainer/model/rnnt_tdnn_transformer.py:41,58-59,69,76-82,85):
 * train: batch statistics incl. padded frames, running stats updated (momentum 0.1); eval: running stats.
 * stats_ws: pk_colstats_ws_floats(C) + 2*C floats scratch.  mean/rstd [C] are saved for the backward. */
long long pk_colstats_ws_floats(int C);
int pk_bn_fwd(const void* x, void* y, int dtype, long long rows, int C, const float* w, const float* b, float eps,
              int train, float momentum, float* run_mean, float* run_var, float* mean, float* rstd, float* stats_ws,
              void* stream);
/* BN backward; relu_mask=1 additionally multiplies by (x > 0): x is the BN input = ReLU output, so the
 * result is the gradient w.r.t. the pre-ReLU TDNN/Linear output.  dw, db [C] are overwritten. */
int pk_bn_bwd(const void* dy, const void* x, void* dx, int dtype, long long rows, int C, const float* w,
              const float* mean, const float* rstd, int train, int relu_mask, float* dw, float* db, float* ws /* pk_colstats_ws_floats(C) */,
              void* stream);
/* out[c] = sum_r x[r,c]  (bias gradients); ws: pk_colstats_ws_floats(C) floats */
int pk_colsum(const void* x, int dtype, long long rows, int C, float* out, float* ws, void* stream);
/* nn.LayerNorm(C, eps=1e-6) (trainer/model/modules/transformer.py:82, position_ffn.py:21) */
int pk_layernorm_fwd(const void* x, void* y, int dtype, long long rows, int C, const float* w, const float* b, float eps,
                     float* mean, float* rstd, void* stream);
int pk_layernorm_bwd(const void* dy, const void* x, void* dx, int dtype, long long rows, int C, const float* w,
                     const float* mean, const float* rstd, float* dw, float* db, void* stream);
/* attention softmax over keys + dropout on the probabilities
 * (trainer/model/modules/multi_headed_attn.py:220-221): S f32 [rows, ld_s] -> P, Pd=dropout(P) [rows, ld_p] */
int pk_softmax_fwd(const float* S, long long ld_s, void* P, void* Pd, int dtype, long long ld_p, long long rows, int n,
                   float drop_p, uint32_t seed, void* stream);
/* masked forward (transformer prediction net, trainer/model/rnnt_conv_transformer_lm.py:66-70 + modules/multi_headed_attn.py:214-216):
 * rows = sequences * heads * q_len; key c of query i = row % q_len is dropped when (causal && c > i) or key_pad[sequence][c] != 0
 * (key_pad: uint8 [sequences][n] or NULL).  The backward is pk_softmax_bwd (a dropped key has P = 0). */
int pk_softmax_masked_fwd(const float* S, long long ld_s, void* P, void* Pd, int dtype, long long ld_p, long long rows, int n,
                          int q_len, int heads, int causal, const uint8_t* key_pad, float drop_p, uint32_t seed, void* stream);
int pk_softmax_bwd(const float* dPd, long long ld_d, const void* P, long long ld_p, void* dS, int dtype, long long rows,
                   int n, float drop_p, uint32_t seed, void* stream);
/* nn.Dropout with the counter-based RNG shared with the GEMM epilogue; mask_nz: dx = dy * (y != 0) * scale */
int pk_dropout(const void* x, void* y, int dtype, long long n, float p, uint32_t seed, void* stream);
int pk_mask_nz(const void* dy, const void* y, void* dx, int dtype, long long n, float scale, void* stream);
int pk_add(const void* a, const void* b, void* o, int dtype, long long n, void* stream);
/* F.log_softmax(scale * x) rows -> f32 (trainer/model/transducer.py:110-111; decoder/transducer_decoder.py:177) */
int pk_log_softmax(const void* x, int dtype, long long ld, float* y, long long rows, int n, float scale, void* stream);
/* gated joint, factored: h[b,t,u,:] = tanh(e1[b,t]+p1[b,u]) * sigmoid(eg[b,t]+pg[b,u])
 * (trainer/model/transducer.py:102-108 without materialising the 2H-wide concat).
 * ex [B*T, 2H], py [B*U1, 2H] = the x / y halves of fc1 | fc_gate applied to encoder / prediction outputs. */
/* h has row pitch ld_h = H or H+8; with H+8 the pad columns are written as (1,0,..,0) so that the fc2 bias gradient
 * falls out of the fc2 wgrad GEMM as one extra column */
int pk_joint_gate_fwd(const void* ex, const void* py, void* h, int dtype, int B, int T, int U1, int H, int ld_h, void* stream);
int pk_joint_gate_bwd(const void* ex, const void* py, const void* dh, void* dex, void* dpy, int dtype, int B, int T, int U1,
                      int H, void* stream);
/* one LSTM time step, pointwise part (nn.LSTM, gate order i,f,g,o; trainer/model/transducer.py:56-61) */
int pk_lstm_cell_fwd(const float* gx, long long ld_gx, const float* gh, long long ld_gh, const float* c_prev, float* c_out,
                     void* h_out, int dtype, long long ld_h, float* gates_save, int B, int H, void* stream);
int pk_lstm_cell_bwd(const void* dh_out, long long ld_dho, const float* dh_rec, const float* dc_next, const float* gates,
                     const float* c, const float* c_prev, void* dgates, int dtype, float* dc_prev, int B, int H, void* stream);
/* nn.Embedding (trainer/model/transducer.py:52-53,94); rows padded to ld_out; padding_idx gets no gradient */
int pk_embedding_fwd(const long long* idx, const float* table, int E, void* out, int dtype, int ld_out, long long n, void* stream);
int pk_embedding_bwd(const long long* idx, const void* dout, int dtype, int ld, int E, float* dtable, long long n,
                     long long padding_idx, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser / BMUF on the flat fp32 parameter vector (pika_b200/csrc/optim.cu).
 *   pk_absmax + pk_sgd_nesterov_clip: clip_grad_norm_(params, max_norm, inf) + optim.SGD(nesterov).step()
 *                                     (trainer/train_transducer_bmuf_otfaug.py:53-55,105-110)
 *   pk_bmuf_delta / pk_bmuf_update  : BmufTrainer.update_and_sync (trainer/bmuf.py:76-100); the sum over
 *                                     ranks between the two is an NCCL all-reduce issued by the host side.
 */
int pk_absmax(const float* x, long long n, float* out, int* nan_flag, void* stream);
/* nan_flag (may be NULL): the flag pk_absmax raised; when set the clip coefficient is NaN, as torch's clip_grad_norm_(inf) gives */
int pk_sgd_nesterov_clip(float* p, const float* g, float* buf, long long n, float lr, float momentum, float max_norm,
                         const float* absmax, const int* nan_flag, int first, void* stream);
int pk_bmuf_delta(const float* glob, const float* local, float* delta, long long n, void* stream);
int pk_bmuf_update(float* glob, float* local, float* delta_prev, const float* delta_sum, long long n, int world,
                   float block_momentum, float block_lr, void* stream);

/* ------------------------------------------------------------------------------------------
 * On-the-fly front end (pika_b200/csrc/frontend.cu): int16 PCM -> speed perturbation + RMS gain +
 * int16 re-quantisation -> Kaldi fbank -> splice -> last-frame padding -> CMN/CMVN -> SpecAugment.
 * Replaces loader/audio.py (AudioSegment.change_speed/normalize/_convert_*), the PyKaldi
 * Fbank.compute_features call and splice() in loader/otf_utt_loader.py:28-46,195-201,218-234,262-270,
 * and trainer/train_transducer_bmuf_otfaug.py:86-93 + utils/spec_augment.py:10-20.
 *   pcm [B, ld_pcm] int16; n_samples, new_len (= int(n/rate)), n_frames (= 1+(new_len-400)/160) [B] int32
 *   rate, target_db [B] f32 (host-drawn, as the reference draws them in the loader thread)
 *   window [400], twiddle [256 x (re,im)], mel_w [n_mel,256], mel_lo/hi [n_mel]: host-built tables
 *   offset/scale [D] CMVN (NULL = off); cmn: subtract the per-utterance mean over the PADDED time axis
 *   (f0,fs,t0,ts): SpecAugment freq/time mask start and span (span 0 = off), shared by the batch
 *   out [B, t_max, D] f32|bf16; wave_i16_out [B, n_max] optional copy of the augmented samples
 *   err_flag: set to 1 if a gain above 300 dB was requested (the reference raises ValueError)
 *   dither, dither_seed: FbankOptions.dither (egs/fbank.conf: dither=1): dither * N(0,1) added to every sample of every extracted
 *   window, as Kaldi's Dither() does; the draws come from a counter-based generator keyed by (utterance, frame, sample, seed) --
 *   Kaldi's own RNG stream is not reproduced.  0 = off (bit-reproducible features, what the parity tests use).
 */
long long pk_frontend_workspace_bytes(int B, int n_max, int t_max, int n_mel, int D);
int pk_frontend_fwd(const short* pcm, long long ld_pcm, const int* n_samples, const float* rate, const int* new_len,
                    const float* target_db, const int* n_frames, int B, int n_max, int t_max, int n_mel, int lctx,
                    int rctx, const float* window, const float* twiddle, const float* mel_w, const int* mel_lo,
                    const int* mel_hi, float preemph, int cmn, const float* offset, const float* scale, int f0, int fs,
                    int t0, int ts, void* out, int out_dtype, short* wave_i16_out, void* workspace,
                    long long workspace_bytes, int* err_flag, float dither, unsigned int dither_seed, void* stream);
int pk_fbank(const float* wave, long long ld_wave, const int* n_frames, int B, int t_max, int n_mel, const float* window,
             const float* twiddle, const float* mel_w, const int* mel_lo, const int* mel_hi, float preemph, float* feats,
             float dither, unsigned int dither_seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched beam search (pika_b200/csrc/beam.cu): one launch per step for the whole batch.
 * Replaces decoder/beam_transducer.py:82-187 (BeamMergeTransducer.advance) and the gather / masked LSTM
 * update / state reordering of decoder/transducer_decoder.py:127-150,173-178,188-202.  Rows are
 * utterance-major (row = b*K + k); blank = blk, EOS = -1.
 * The loop is meant to be replayed from a CUDA graph, so nothing step-dependent is a kernel argument: step_ctx [2] int32 lives in
 * device memory -- step_ctx[0] = step index (the kernels address next_ys[step] / prev_ks[step] themselves), step_ctx[1] = 1 while
 * `not all(b.done() for b in beam)` (decoder/transducer_decoder.py:123) and the history buffers have room; pk_beam_step_end
 * advances / latches both at the end of a step and every kernel of a dead step is a no-op.  Initialise step_ctx = {0, 1}.
 */
int pk_beam_prepare(const int* next_ys, const int* step_ctx, int* t_idx, const void* enc, int dtype, int Tenc, int H, void* enc_hid,
                    const float* embed, int E, void* x_emb, int ld_x, int K, int blk, int rows, void* stream);
int pk_beam_lstm_cell(const float* gates, const int* next_ys, const int* step_ctx, int blk, void* h, int dtype, float* c, int rows, int H,
                      void* stream);
int pk_beam_step_end(int* step_ctx, const int* not_done, int max_steps, void* stream);
int pk_beam_gate(const float* a, void* h, int dtype, int rows, int H, void* stream);
/* one BeamMergeTransducer.advance for every utterance.  The word scores are given as the joint's logits [B*K rows, pitch ldv] f32 plus
 * the per-row log-sum-exp of sm_scale * logits (pk_row_lse): word_probs[r, v] = logits[r, v] * sm_scale - row_lse[r], the very expression
 * pk_log_softmax evaluates (decoder/transducer_decoder.py:177), formed on the fly so that the [B*K, V] log-prob tensor is never written.
 * t_idx [B*K], histories next_ys [S+1,B,K], prev_ks [S,B,K]; partial hypotheses hyp_tok [2,B,K,L] / hyp_len [2,B,K];
 * finished lists fin_* [B,cap]; *not_done_total is decremented when an utterance becomes done. */
int pk_row_lse(const void* x, int dtype, long long ld, float* lse, long long rows, int n, float scale, void* stream);
int pk_beam_advance(const float* logits, int ldv, const float* row_lse, float sm_scale, const int* t_idx, const int* num_frames, const int* max_len, float* scores,
                    int* next_ys, int* prev_ks, int* hyp_tok, int* hyp_len, float* fin_score, int* fin_step, int* fin_k,
                    int* fin_count, int* eos_top, int* done, int* not_done_total, int B, int K, int V, int L, int cap,
                    const int* step_ctx, int blk, int n_best, int beam_prune, void* stream);
int pk_beam_reorder(const int* prev_ks, const int* step_ctx, const void* h_in, const float* c_in, const int* t_in, void* h_out, float* c_out,
                    int* t_out, int dtype, int K, int H, int layers, int rows, void* stream);
/* On-the-fly FST shallow fusion inside the beam step: decoder/beam_transducer.py:135-159,167-176 with the arc search of
 * decoder/sorted_matcher.py:24-111 over a flattened arc table (arcs of a state sorted by input label, label = token + 1).
 * Per beam the active FST states are an insertion-ordered set of at most max_states (state, cost) pairs in double precision;
 * err_flag is raised when a set would overflow.  lm_scores [B,K] f32 and the sets ([2,B,K,max_states], [2,B,K]) are state the
 * caller zero-initialises before step 0. */
typedef struct {
    const int* arc_off;       /* [n_states + 1] */
    const int* arc_ilabel;    /* [n_arcs] */
    const double* arc_weight; /* [n_arcs] */
    const int* arc_next;      /* [n_arcs] */
    const double* finals;     /* [n_states], +inf = not final */
    int backoff_id, n_disambig;
    int disambig_ids[4];
} pk_lm_fst;
int pk_beam_advance_lm(const float* logits, int ldv, const float* row_lse, float sm_scale, const int* t_idx, const int* num_frames, const int* max_len, float* scores,
                       int* next_ys, int* prev_ks, int* hyp_tok, int* hyp_len, float* fin_score, int* fin_step, int* fin_k,
                       int* fin_count, int* eos_top, int* done, int* not_done_total, int B, int K, int V, int L, int cap,
                       const int* step_ctx, int blk, int n_best, int beam_prune, const pk_lm_fst* fst, double lm_scale, double nonblk_reward,
                       int* set_state, double* set_cost, int* set_n, float* lm_scores, int max_states, int* err_flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * Persistent LSTM layer (pika_b200/csrc/lstm_seq.cu): the whole recurrence of one nn.LSTM layer in one
 * cooperative launch (trainer/model/transducer.py:56-61,95).  B <= 32 sequences, zero initial state.
 *   fwd: gx f32 [B,U,4H] = x W_ih^T + b_ih + b_hh; w_hh bf16 [4H,H]; out [B,U,H] (f32|bf16);
 *        gates_save f32 [U,B,4H], cs f32 [U,B,H] are kept for the backward.
 *   bwd: dout [B,U,H] -> dG bf16 [U,B,4H] (gradient w.r.t. the pre-activation gates, time-major).
 *   ws : pk_lstm_seq_workspace_bytes(H) bytes of scratch (grid barrier + hidden-state exchange).
 */
long long pk_lstm_seq_workspace_bytes(int H);
int pk_lstm_seq_fwd(const float* gx, const void* w_hh_bf16, void* out, int out_dtype, float* gates_save, float* cs, int B,
                    int U, int H, void* ws, void* stream);
int pk_lstm_seq_bwd(const void* dout, int dtype, const float* gates_save, const float* cs, const void* w_hh_bf16, void* dG_bf16,
                    int B, int U, int H, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------
 * MBR training step helpers (trainer/train_transducer_mbr_bmuf_otfaug.py:197-235): the joint is evaluated only
 * on the (t,u) nodes of each N-best alignment, so rows are gathered, and the sparse mbr_grad through
 * log_softmax(sm_scale * out) is formed directly.
 */
int pk_gather_rows(const void* src, const int* idx, void* dst, int dtype, long long rows, int C, void* stream);
int pk_scatter_add_rows(const void* src, const int* idx, float* dst, int dtype, long long rows, int C, void* stream);
/* dz[r] = scale * coef[r] * (onehot(tok[r]) - softmax(scale * z[r]));  z, dz [rows, ld], first n columns valid */
int pk_ce_grad(const void* z, int dtype, long long ld, const int* tok, const float* coef, float scale, void* dz,
               long long rows, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif
