// RNN-T loss (alpha/beta lattice) + gradient w.r.t. the joint logits, fused with log-softmax.
//
// Replaces  F.log_softmax (trainer/model/transducer.py:110-111)  +  warp_rnnt.RNNTLoss.apply
// (trainer/train_transducer_bmuf_otfaug.py:58,97-99) and their autograd backward.
//
//   pass 1  rnnt_rowstats   one warp per joint node (b,t,u): streams the V logits once (16-byte
//                           coalesced loads, several in flight per lane), online log-sum-exp ->
//                           lse[b,t,u], lp_blank, lp_label written in a diagonal-major ("skewed")
//                           layout so that pass 2 reads each anti-diagonal contiguously.
//   pass 2  rnnt_lattice    one CTA per utterance: one thread group sweeps alpha, a second one beta
//                           (anti-diagonal wavefront, one lattice cell per thread, previous diagonal
//                           staged in shared memory, next diagonal's log-probs prefetched, fp64);
//                           then all threads emit the per-node gradient coefficients.  Never touches V.
//   pass 3  rnnt_grad       one warp per node: re-reads the logits row and writes
//                           dlogits = -softmax * (gb + gl) + [v==blank] gb + [v==label] gl
//                           (in place if dlogits aliases logits).
// Algorithmic HBM traffic: 3 * B*T*(U+1)*V * sizeof(elem)  (+ O(nodes) fp32).
#include <cstdlib>

#include "../../include/pika_b200.h"
#include "common.cuh"

namespace pk {
void count_launch();

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <typename T> struct Vec16;
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    PK_DEVICE static void unpack(const uint4& q, float (&f)[8]) {
        f[0] = bf16lo(q.x); f[1] = bf16hi(q.x); f[2] = bf16lo(q.y); f[3] = bf16hi(q.y);
        f[4] = bf16lo(q.z); f[5] = bf16hi(q.z); f[6] = bf16lo(q.w); f[7] = bf16hi(q.w);
    }
    PK_DEVICE static uint4 pack(const float (&f)[8]) {
        return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
    PK_DEVICE static float vmax(const uint4& q) {         // packed bf16x2 max: 3 + 1 instructions for 8 elements
        const __nv_bfloat162 a = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q.x), *reinterpret_cast<const __nv_bfloat162*>(&q.y));
        const __nv_bfloat162 b = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q.z), *reinterpret_cast<const __nv_bfloat162*>(&q.w));
        const __nv_bfloat162 c = __hmax2(a, b);
        return fmaxf(__low2float(c), __high2float(c));
    }
};
template <> struct Vec16<float> {
    static constexpr int N = 4;
    PK_DEVICE static void unpack(const uint4& q, float (&f)[4]) {
        f[0] = __uint_as_float(q.x); f[1] = __uint_as_float(q.y); f[2] = __uint_as_float(q.z); f[3] = __uint_as_float(q.w);
    }
    PK_DEVICE static uint4 pack(const float (&f)[4]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    PK_DEVICE static float vmax(const uint4& q) {
        return fmaxf(fmaxf(__uint_as_float(q.x), __uint_as_float(q.y)), fmaxf(__uint_as_float(q.z), __uint_as_float(q.w)));
    }
};

PK_DEVICE uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
PK_DEVICE uint4 ld_plain(const uint4* p) { return *p; }
PK_DEVICE void st_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct RnntDims {
    int B, T, U1, V, ldv;     // padded batch dims; ldv = row pitch of logits in elements
    int ld_labels;
    int ND;                   // T + U1 - 1 diagonals
};
PK_DEVICE size_t skew_index(const RnntDims& d, int b, int t, int u) { return ((size_t)b * d.ND + (t + u)) * d.U1 + u; }

constexpr int ROWSTATS_UNROLL = 4;

// ------------------------------------------------------------------------------------ pass 1
template <typename T>
__global__ void __launch_bounds__(256) rnnt_rowstats_kernel(const T* __restrict__ logits, const int* __restrict__ labels,
                                                            const int* __restrict__ frame_lens, const int* __restrict__ label_lens,
                                                            RnntDims d, float* __restrict__ lse_out, float* __restrict__ lpb_skew,
                                                            float* __restrict__ lpl_skew) {
    constexpr int VN = Vec16<T>::N;
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long rows = (long long)d.B * d.T * d.U1;
    const int nvec = (d.V + VN - 1) / VN;
    for (long long row = warp_global; row < rows; row += n_warps) {
        const int u = (int)(row % d.U1);
        const long long bt = row / d.U1;
        const int t = (int)(bt % d.T);
        const int b = (int)(bt / d.T);
        const int Tn = frame_lens[b], Un = label_lens[b];
        if (t >= Tn || u > Un) continue;       // padded node: never read (warp-uniform)
        const T* rp = logits + row * (long long)d.ldv;
        const uint4* vp = reinterpret_cast<const uint4*>(rp);
        float m = -INFINITY, s = 0.f;          // running max (in log2 units) and sum of 2^(x*log2e - m)
        const int nfull = d.V / VN;            // vectors without a tail; the (rare) partial vector is handled after the loop
        for (int i0 = lane; i0 < nfull; i0 += 32 * ROWSTATS_UNROLL) {
            uint4 q[ROWSTATS_UNROLL];
#pragma unroll
            for (int k = 0; k < ROWSTATS_UNROLL; ++k) {
                const int i = i0 + k * 32;
                if (i < nfull) q[k] = ld_stream(vp + i);
            }
#pragma unroll
            for (int k = 0; k < ROWSTATS_UNROLL; ++k) {
                const int i = i0 + k * 32;
                if (i < nfull) {
                    const float cm = Vec16<T>::vmax(q[k]) * kLog2e;
                    if (cm > m) { s *= exp2f(m - cm); m = cm; }   // rare after the first chunks
                    float f[VN];
                    Vec16<T>::unpack(q[k], f);
#pragma unroll
                    for (int e = 0; e < VN; ++e) s += ex2_approx(fmaf(f[e], kLog2e, -m));   // raw MUFU.EX2: no denormal fix-up sequence
                }
            }
        }
        if (nfull * VN < d.V && lane == 0) {   // partial last vector (V not a multiple of 16 bytes)
            for (int v = nfull * VN; v < d.V; ++v) {
                const float x = to_f32<T>(rp[v]) * kLog2e;
                if (x > m) { s *= exp2f(m - x); m = x; }
                s += exp2f(x - m);
            }
        }
        // combine the 32 lane-local (m, s) pairs
        const float mw = warp_max(m);
        s = (m == -INFINITY) ? 0.f : s * exp2f(m - mw);
        s = warp_sum(s);
        if (lane == 0) {
            const float lse = (mw + log2f(s)) * kLn2;
            const float zb = to_f32<T>(rp[0]);
            lse_out[row] = lse;
            const size_t sk = skew_index(d, b, t, u);
            lpb_skew[sk] = zb - lse;
            if (u < Un) {
                const int y = labels[(size_t)b * d.ld_labels + u];
                lpl_skew[sk] = to_f32<T>(rp[y]) - lse;
            }
        }
    }
}

// ------------------------------------------------------------------------------------ pass 1, fused variant
// The producing GEMM already reduced every logits row to per-N-tile (max, sum) pairs (pk_gemm_desc.row_lse);
// this kernel only merges the n_parts pairs of a row and gathers the blank / label logits: ~100 bytes per row
// instead of a full read of the row.
template <typename T>
__global__ void __launch_bounds__(256) rnnt_rowfinish_kernel(const T* __restrict__ logits, const int* __restrict__ labels,
                                                             const int* __restrict__ frame_lens, const int* __restrict__ label_lens,
                                                             RnntDims d, const float2* __restrict__ parts, int n_parts,
                                                             float* __restrict__ lse_out, float* __restrict__ lpb_skew,
                                                             float* __restrict__ lpl_skew) {
    const long long rows = (long long)d.B * d.T * d.U1;
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (long long)gridDim.x * blockDim.x) {
        const int u = (int)(row % d.U1);
        const long long bt = row / d.U1;
        const int t = (int)(bt % d.T);
        const int b = (int)(bt / d.T);
        const int Tn = frame_lens[b], Un = label_lens[b];
        if (t >= Tn || u > Un) continue;
        float m = -INFINITY;
        for (int i = 0; i < n_parts; ++i) m = fmaxf(m, parts[(size_t)i * rows + row].x);
        float sum = 0.f;
        for (int i = 0; i < n_parts; ++i) {
            const float2 ps = parts[(size_t)i * rows + row];
            sum += ps.y * exp2f(ps.x - m);
        }
        const float lse = (m + log2f(sum)) * kLn2;
        const T* rp = logits + row * (long long)d.ldv;
        lse_out[row] = lse;
        const size_t sk = skew_index(d, b, t, u);
        lpb_skew[sk] = to_f32<T>(rp[0]) - lse;
        if (u < Un) {
            const int y = labels[(size_t)b * d.ld_labels + u];
            lpl_skew[sk] = to_f32<T>(rp[y]) - lse;
        }
    }
}

// ------------------------------------------------------------------------------------ pass 2
// The lattice runs in double precision: alpha/beta reach magnitudes of (T+U)*log V ~ 3000 where an
// fp32 ulp (2.4e-4) would show up as a 1e-4-level relative error in the gradients.  The DP touches
// only O(T*U) values per utterance, so fp64 costs nothing measurable next to the V-axis passes.
typedef double lat_t;
#define LAT_NEG_INF (-(double)INFINITY)
PK_DEVICE lat_t lse2(lat_t a, lat_t b) {
    const lat_t mx = fmax(a, b), mn = fmin(a, b);
    if (mx == LAT_NEG_INF) return LAT_NEG_INF;
    return mx + log1p(exp(mn - mx));
}

// One CTA per utterance, 2*G threads: threads [0,G) sweep alpha over ascending anti-diagonals,
// threads [G,2G) sweep beta over descending ones; thread j owns label positions u = j, j+G, ...
// The previous diagonal lives in shared memory (ping-pong), each group synchronises with its own
// named barrier, and the log-probs of the NEXT diagonal are prefetched before the barrier so the
// L2 latency overlaps the dependent LSE chain.
constexpr int LAT_MAX_G = 512;
constexpr int LAT_MAX_CPT = 4;      // cells per thread => U1 <= 2048

__global__ void __launch_bounds__(2 * LAT_MAX_G) rnnt_lattice_kernel(
    const int* __restrict__ frame_lens, const int* __restrict__ label_lens, RnntDims d, int G, int cpt,
    const float* __restrict__ lpb_skew, const float* __restrict__ lpl_skew, lat_t* __restrict__ alpha_skew,
    lat_t* __restrict__ beta_skew, const float* __restrict__ grad_scale, float* __restrict__ costs,
    float* __restrict__ gb_out, float* __restrict__ gl_out) {
    extern __shared__ lat_t sm[];            // 2 groups x 2 diagonals x (U1 + 2)
    const int b = blockIdx.x;
    const int grp = threadIdx.x >= G ? 1 : 0;
    const int j = threadIdx.x - grp * G;
    const int T = frame_lens[b], U = label_lens[b];
    const int W = d.U1 + 2;
    lat_t* buf0 = sm + grp * 2 * W + 1;      // index -1 .. U1 valid
    lat_t* buf1 = buf0 + W;
    __shared__ lat_t s_ll;
    const size_t base = (size_t)b * d.ND * d.U1;
    const bool valid = (T > 0 && T <= d.T && U >= 0 && U < d.U1);
    if (valid) {
        for (int i = j - 1; i <= d.U1; i += G) { buf0[i] = LAT_NEG_INF; buf1[i] = LAT_NEG_INF; }
        named_bar_sync(1 + grp, G);
        const int last = T - 1 + U;           // last diagonal
        if (grp == 0) {
            // ---------------- alpha: diagonals ascending
            lat_t* prev = buf0;
            lat_t* cur = buf1;
            if (j == 0) { cur[0] = 0.0; alpha_skew[base] = 0.0; }
            float nb[LAT_MAX_CPT], nl[LAT_MAX_CPT];   // lpb(d-1,u), lpl(d-1,u-1) for the upcoming diagonal
#pragma unroll
            for (int c = 0; c < LAT_MAX_CPT; ++c) {
                const int u = j + c * G;
                nb[c] = (c < cpt && u <= U && last >= 1) ? lpb_skew[base + u] : 0.f;
                nl[c] = (c < cpt && u >= 1 && u <= U && last >= 1) ? lpl_skew[base + u - 1] : 0.f;
            }
            named_bar_sync(1, G);
            for (int dg = 1; dg <= last; ++dg) {
                lat_t* tsw = prev; prev = cur; cur = tsw;
                const int lo = max(0, dg - (T - 1)), hi = min(U, dg);
                float pb[LAT_MAX_CPT], pl[LAT_MAX_CPT];
#pragma unroll
                for (int c = 0; c < LAT_MAX_CPT; ++c) { pb[c] = nb[c]; pl[c] = nl[c]; }
                if (dg < last) {                       // prefetch diagonal dg (used at dg+1)
                    const size_t nbase = base + (size_t)dg * d.U1;
#pragma unroll
                    for (int c = 0; c < LAT_MAX_CPT; ++c) {
                        const int u = j + c * G;
                        if (c < cpt && u <= U) {
                            nb[c] = lpb_skew[nbase + u];
                            nl[c] = (u >= 1) ? lpl_skew[nbase + u - 1] : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < LAT_MAX_CPT; ++c) {
                    const int u = j + c * G;
                    if (c < cpt && u >= lo && u <= hi) {
                        lat_t a = LAT_NEG_INF, cc = LAT_NEG_INF;
                        if (u <= dg - 1) a = prev[u] + pb[c];              // from (t-1, u) via blank
                        if (u >= 1) cc = prev[u - 1] + pl[c];              // from (t, u-1) via label u
                        const lat_t v = lse2(a, cc);
                        cur[u] = v;
                        alpha_skew[base + (size_t)dg * d.U1 + u] = v;
                    }
                }
                named_bar_sync(1, G);
            }
        } else {
            // ---------------- beta: diagonals descending
            lat_t* prev = buf0;
            lat_t* cur = buf1;
            if (j == 0) {
                const lat_t v = lpb_skew[base + (size_t)last * d.U1 + U];
                cur[U] = v;
                beta_skew[base + (size_t)last * d.U1 + U] = v;
            }
            float nb[LAT_MAX_CPT], nl[LAT_MAX_CPT];   // lpb(d,u), lpl(d,u) of the upcoming diagonal
#pragma unroll
            for (int c = 0; c < LAT_MAX_CPT; ++c) {
                const int u = j + c * G;
                const bool ok = c < cpt && u <= U && last >= 1;
                nb[c] = ok ? lpb_skew[base + (size_t)(last - 1) * d.U1 + u] : 0.f;
                nl[c] = ok ? lpl_skew[base + (size_t)(last - 1) * d.U1 + u] : 0.f;
            }
            named_bar_sync(2, G);
            for (int dg = last - 1; dg >= 0; --dg) {
                lat_t* tsw = prev; prev = cur; cur = tsw;
                const int lo = max(0, dg - (T - 1)), hi = min(U, dg);
                float pb[LAT_MAX_CPT], pl[LAT_MAX_CPT];
#pragma unroll
                for (int c = 0; c < LAT_MAX_CPT; ++c) { pb[c] = nb[c]; pl[c] = nl[c]; }
                if (dg > 0) {
                    const size_t nbase = base + (size_t)(dg - 1) * d.U1;
#pragma unroll
                    for (int c = 0; c < LAT_MAX_CPT; ++c) {
                        const int u = j + c * G;
                        if (c < cpt && u <= U) { nb[c] = lpb_skew[nbase + u]; nl[c] = lpl_skew[nbase + u]; }
                    }
                }
#pragma unroll
                for (int c = 0; c < LAT_MAX_CPT; ++c) {
                    const int u = j + c * G;
                    if (c < cpt && u >= lo && u <= hi) {
                        const int t = dg - u;
                        lat_t a = LAT_NEG_INF, cc = LAT_NEG_INF;
                        if (t + 1 <= T - 1) a = prev[u] + pb[c];            // to (t+1, u) via blank
                        if (u + 1 <= U) cc = prev[u + 1] + pl[c];           // to (t, u+1) via label u+1
                        const lat_t v = lse2(a, cc);
                        cur[u] = v;
                        beta_skew[base + (size_t)dg * d.U1 + u] = v;
                    }
                }
                named_bar_sync(2, G);
            }
            if (j == 0) { s_ll = cur[0]; costs[b] = (float)(-cur[0]); }
        }
    } else if (threadIdx.x == 0) {
        costs[b] = 0.f;
    }
    __syncthreads();
    // ---------------- per-node gradient coefficients (natural [b,t,u] layout); zero for padded nodes
    const float gs = grad_scale ? grad_scale[b] : 1.f;
    const lat_t ll = valid ? s_ll : 0.0;
    const int nodes = d.T * d.U1;
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
        const int t = i / d.U1, u = i - t * d.U1;
        float gb = 0.f, gl = 0.f;
        if (valid && t < T && u <= U) {
            const size_t sk = skew_index(d, b, t, u);
            const lat_t a = alpha_skew[sk];
            lat_t bn;
            if (t < T - 1) bn = beta_skew[skew_index(d, b, t + 1, u)];
            else bn = (u == U) ? 0.0 : LAT_NEG_INF;
            gb = (float)(-exp(a + bn + (lat_t)lpb_skew[sk] - ll)) * gs;
            if (u < U) gl = (float)(-exp(a + beta_skew[skew_index(d, b, t, u + 1)] + (lat_t)lpl_skew[sk] - ll)) * gs;
            if (!(gb == gb)) gb = 0.f;
            if (!(gl == gl)) gl = 0.f;
        }
        gb_out[(size_t)b * nodes + i] = gb;
        gl_out[(size_t)b * nodes + i] = gl;
    }
}

// ------------------------------------------------------------------------------------ pass 3
// One CTA walks a contiguous block of joint nodes; thread i owns the 16-byte column groups i, i+256, ... of every
// row (32 columns per thread -> V <= 8192), so the column sums of dlogits (= the fc2 bias
// gradient) accumulate in registers for free.  GRAD_RU rows are in flight per thread.
constexpr int GRAD_THREADS = 256;
template <typename T> struct GradCfg { static constexpr int MAXG = 32 / Vec16<T>::N; };   // 16-byte groups per thread: V <= 8192 for bf16 and f32
template <typename T, int GRAD_RU, int MINB>
__global__ void __launch_bounds__(GRAD_THREADS, MINB) rnnt_grad_kernel(const T* logits, const int* __restrict__ labels,
                                                                    const int* __restrict__ label_lens, RnntDims d,
                                                                    const float* __restrict__ lse_in, const float* __restrict__ gb_in,
                                                                    const float* __restrict__ gl_in, T* dlogits, float* __restrict__ colsum,
                                                                    long long rows_per_cta) {
    constexpr int VN = Vec16<T>::N;
    constexpr int GRAD_MAXG = GradCfg<T>::MAXG;
    extern __shared__ float gsm[];                          // per-row scalars of this CTA's block: gb, gl, lse*log2e, label
    float* s_gb = gsm;
    float* s_gl = gsm + rows_per_cta;
    float* s_l2 = gsm + 2 * rows_per_cta;
    int* s_y = reinterpret_cast<int*>(gsm + 3 * rows_per_cta);
    const long long rows = (long long)d.B * d.T * d.U1;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = min(rows, r0 + rows_per_cta);
    const int nvec_ld = d.ldv / VN;
    const int tid = threadIdx.x;
    for (long long row = r0 + tid; row < r1; row += GRAD_THREADS) {
        const int i = (int)(row - r0);
        s_gb[i] = gb_in[row];
        s_gl[i] = gl_in[row];
        s_l2[i] = lse_in[row] * kLog2e;
        const int u = (int)(row % d.U1);
        const int b = (int)(row / ((long long)d.T * d.U1));
        s_y[i] = (u < label_lens[b]) ? labels[(size_t)b * d.ld_labels + u] : -1;
    }
    __syncthreads();
    float cs[GRAD_MAXG][VN];
#pragma unroll
    for (int gq = 0; gq < GRAD_MAXG; ++gq)
#pragma unroll
        for (int e = 0; e < VN; ++e) cs[gq][e] = 0.f;
    for (long long rb = r0; rb < r1; rb += GRAD_RU) {
        uint4 q[GRAD_RU][GRAD_MAXG];
#pragma unroll
        for (int k = 0; k < GRAD_RU; ++k) {
            const long long row = rb + k;
            if (row < r1 && (s_gb[row - r0] != 0.f || s_gl[row - r0] != 0.f)) {
                const uint4* vp = reinterpret_cast<const uint4*>(logits + row * (long long)d.ldv);
#pragma unroll
                for (int gq = 0; gq < GRAD_MAXG; ++gq) {
                    const int i = tid + gq * GRAD_THREADS;
                    if (i < nvec_ld) q[k][gq] = ld_plain(vp + i);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < GRAD_RU; ++k) {
            const long long row = rb + k;
            if (row >= r1) continue;
            const float gb = s_gb[row - r0], gl = s_gl[row - r0];
            uint4* op = reinterpret_cast<uint4*>(dlogits + row * (long long)d.ldv);
            if (gb == 0.f && gl == 0.f) {               // padded (or zero-probability) node: zeros, nothing read
#pragma unroll
                for (int gq = 0; gq < GRAD_MAXG; ++gq) {
                    const int i = tid + gq * GRAD_THREADS;
                    if (i < nvec_ld) st_stream(op + i, make_uint4(0, 0, 0, 0));
                }
                continue;
            }
            const int y = s_y[row - r0];
            const float l2 = s_l2[row - r0];
            const float gsum = -(gb + gl);
#pragma unroll
            for (int gq = 0; gq < GRAD_MAXG; ++gq) {
                const int i = tid + gq * GRAD_THREADS;
                if (i < nvec_ld) {
                    float f[VN];
                    Vec16<T>::unpack(q[k][gq], f);
#pragma unroll
                    for (int e = 0; e < VN; ++e) f[e] = ex2_approx(fmaf(f[e], kLog2e, -l2)) * gsum;
                    const int v0 = i * VN;
                    if (v0 == 0) f[0] += gb;                                   // blank column
                    if (y >= v0 && y < v0 + VN) {                              // label column (one thread per row)
#pragma unroll
                        for (int e = 0; e < VN; ++e) if (v0 + e == y) f[e] += gl;
                    }
                    if (v0 + VN > d.V) {                                       // row padding [V, ldv)
#pragma unroll
                        for (int e = 0; e < VN; ++e) if (v0 + e >= d.V) f[e] = 0.f;
                    }
                    const uint4 packed = Vec16<T>::pack(f);
                    st_stream(op + i, packed);
                    if (colsum) {                         // sum what was actually stored (bf16-rounded in production)
                        float w[VN];
                        Vec16<T>::unpack(packed, w);
#pragma unroll
                        for (int e = 0; e < VN; ++e) cs[gq][e] += w[e];
                    }
                }
            }
        }
    }
    if (colsum) {
        // one partial row per CTA (no atomics): colsum_partials_kernel adds them in a fixed order, so the fc2 bias
        // gradient is bit-reproducible from run to run
        float* part = colsum + (size_t)blockIdx.x * d.ldv;
#pragma unroll
        for (int gq = 0; gq < GRAD_MAXG; ++gq) {
            const int i = tid + gq * GRAD_THREADS;
            if (i < nvec_ld) {
#pragma unroll
                for (int e = 0; e < VN; ++e) part[i * VN + e] = cs[gq][e];
            }
        }
    }
}

// out[c] = sum over the n_part partial rows.  Block = 32 columns x 8 row groups; every group adds its rows in index order and the
// eight group sums are combined in a fixed order, so the result is bit-reproducible.
__global__ void __launch_bounds__(256) colsum_partials_kernel(const float* __restrict__ part, int n_part, int ld, float* __restrict__ out) {
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (c < ld) {
        const int per = (n_part + 7) / 8;
        const int i0 = ty * per, i1 = min(n_part, i0 + per);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int i = i0;
        for (; i + 4 <= i1; i += 4) {
            a0 += part[(size_t)(i + 0) * ld + c]; a1 += part[(size_t)(i + 1) * ld + c];
            a2 += part[(size_t)(i + 2) * ld + c]; a3 += part[(size_t)(i + 3) * ld + c];
        }
        for (; i < i1; ++i) a0 += part[(size_t)i * ld + c];
        a = (a0 + a1) + (a2 + a3);
    }
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < ld) {
        float t = sm[0][tx];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += sm[k][tx];
        out[c] = t;
    }
}

}  // namespace pk

extern "C" long long pk_rnnt_loss_workspace_bytes(int B, int T, int U1) {
    const long long nd = (long long)T + U1 - 1;
    const long long skew = (long long)B * nd * U1;
    const long long nodes = (long long)B * T * U1;
    return (2 * skew + 3 * nodes) * 4 + 2 * skew * 8 + 256;
}
// gradient-kernel grid: shared by the launch and by the column-sum workspace query
static void rnnt_grad_grid(long long rows, int* ru_out, int* ctas_per_sm_out, long long* rpc_out, int* ggrid_out, int* variant_out) {
    static int variant = -1;                         // tuning hook (PK_RNNT_GRAD_VARIANT=0..3), default chosen by measurement
    if (variant < 0) { const char* e = getenv("PK_RNNT_GRAD_VARIANT"); variant = e ? atoi(e) : 1; }
    const int ru = (variant == 0) ? 1 : (variant == 3 ? 4 : 2);
    const int ctas_per_sm = (variant == 2) ? 2 : 3;
    const int gcta = pk::num_sms() * ctas_per_sm * 4;
    long long rpc = (rows + gcta - 1) / gcta;
    rpc = (rpc + ru - 1) / ru * ru;
    if (rpc > 2048) rpc = 2048;                      // 16 bytes of shared memory per row
    *ru_out = ru; *ctas_per_sm_out = ctas_per_sm; *rpc_out = rpc; *ggrid_out = (int)((rows + rpc - 1) / rpc); *variant_out = variant;
}
extern "C" long long pk_rnnt_loss_colsum_workspace_bytes(int B, int T, int U1, int ldv) {
    int ru, cps, ggrid, variant; long long rpc;
    rnnt_grad_grid((long long)B * T * U1, &ru, &cps, &rpc, &ggrid, &variant);
    return (long long)ggrid * ldv * 4;
}

static int rnnt_loss_impl(const void* logits, int dtype, const int* labels, const int* frame_lens,
                          const int* label_lens, int B, int T, int U1, int V, int ldv, int ld_labels,
                          const float* grad_scale, float* costs, void* dlogits, float* dlogits_colsum, void* workspace,
                          long long workspace_bytes, const float* row_lse, int n_parts, void* stream_v) {
    using namespace pk;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    PK_CHECK_ARG(dtype == PK_F32 || dtype == PK_BF16, "bad dtype");
    PK_CHECK_ARG(B > 0 && T > 0 && U1 > 0 && V > 1, "bad dims");
    const int vn = dtype == PK_F32 ? 4 : 8;
    PK_CHECK_ARG(ldv >= V && ldv % vn == 0, "ldv must be >= V and a multiple of 16 bytes");
    PK_CHECK_ARG((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "logits not 16B aligned");
    PK_CHECK_ARG(dlogits == nullptr || (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0, "dlogits not 16B aligned");
    PK_CHECK_ARG(workspace_bytes >= pk_rnnt_loss_workspace_bytes(B, T, U1), "workspace too small");
    RnntDims d{B, T, U1, V, ldv, ld_labels, T + U1 - 1};
    const size_t skew = (size_t)B * d.ND * U1, nodes = (size_t)B * T * U1;
    double* wsd = reinterpret_cast<double*>(workspace);          // doubles first (8-byte alignment)
    double* alpha = wsd; double* beta = alpha + skew;
    float* lpb = reinterpret_cast<float*>(beta + skew); float* lpl = lpb + skew;
    float* lse = lpl + skew; float* gb = lse + nodes; float* gl = gb + nodes;

    const long long rows = (long long)nodes;
    const int warps_per_cta = 8;
    long long want = (rows + warps_per_cta - 1) / warps_per_cta;
    const long long cap = (long long)num_sms() * 8 * 4;      // persistent grid-stride: 8 CTAs/SM x 4 waves
    const int grid = (int)(want < cap ? want : cap);
    if (row_lse != nullptr) {
        PK_CHECK_ARG(n_parts >= 1, "n_parts must be >= 1");
        const int fgrid = (int)((rows + 255) / 256 < (long long)num_sms() * 8 ? (rows + 255) / 256 : (long long)num_sms() * 8);
        if (dtype == PK_BF16)
            rnnt_rowfinish_kernel<__nv_bfloat16><<<fgrid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), labels, frame_lens,
                                                                           label_lens, d, reinterpret_cast<const float2*>(row_lse), n_parts,
                                                                           lse, lpb, lpl);
        else
            rnnt_rowfinish_kernel<float><<<fgrid, 256, 0, stream>>>(reinterpret_cast<const float*>(logits), labels, frame_lens, label_lens, d,
                                                                   reinterpret_cast<const float2*>(row_lse), n_parts, lse, lpb, lpl);
    } else if (dtype == PK_BF16)
        rnnt_rowstats_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), labels,
                                                                     frame_lens, label_lens, d, lse, lpb, lpl);
    else
        rnnt_rowstats_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(logits), labels, frame_lens,
                                                              label_lens, d, lse, lpb, lpl);
    PK_CHECK_LAUNCH(); count_launch();
    const int lat_smem = 2 * 2 * (U1 + 2) * 8;
    int G = ((U1 + 31) / 32) * 32;
    if (G > LAT_MAX_G) G = LAT_MAX_G;
    const int cpt = (U1 + G - 1) / G;
    PK_CHECK_ARG(cpt <= LAT_MAX_CPT, "U too large for the lattice kernel (U+1 <= 2048)");
    rnnt_lattice_kernel<<<B, 2 * G, lat_smem, stream>>>(frame_lens, label_lens, d, G, cpt, lpb, lpl, alpha, beta, grad_scale,
                                                     costs, gb, gl);
    PK_CHECK_LAUNCH(); count_launch();
    if (dlogits != nullptr) {
        PK_CHECK_ARG(ldv <= 8192, "V too large for the gradient kernel (V <= 8192)");
        int ru, ctas_per_sm, ggrid, variant; long long rpc;
        rnnt_grad_grid(rows, &ru, &ctas_per_sm, &rpc, &ggrid, &variant);
        (void)ru; (void)ctas_per_sm;
        float* cs_part = nullptr;                        // [ggrid][ldv] per-CTA partial column sums, after the lattice workspace
        if (dlogits_colsum) {
            const long long base_bytes = pk_rnnt_loss_workspace_bytes(B, T, U1);
            PK_CHECK_ARG(workspace_bytes >= base_bytes + (long long)ggrid * ldv * 4,
                         "workspace too small for the column sums (add pk_rnnt_loss_colsum_workspace_bytes)");
            cs_part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + base_bytes);
        }
        const size_t gsmem = (size_t)rpc * 16;
#define PK_GRAD_LAUNCH(TT, RU, MB)                                                                                         \
        rnnt_grad_kernel<TT, RU, MB><<<ggrid, GRAD_THREADS, gsmem, stream>>>(reinterpret_cast<const TT*>(logits), labels, label_lens, d, lse, \
                                                                            gb, gl, reinterpret_cast<TT*>(dlogits), cs_part, rpc)
        if (dtype == PK_BF16) {
            if (variant == 0) PK_GRAD_LAUNCH(__nv_bfloat16, 1, 3);
            else if (variant == 2) PK_GRAD_LAUNCH(__nv_bfloat16, 2, 2);
            else if (variant == 3) PK_GRAD_LAUNCH(__nv_bfloat16, 4, 2);
            else PK_GRAD_LAUNCH(__nv_bfloat16, 2, 3);
        } else {
            PK_GRAD_LAUNCH(float, 2, 2);
        }
#undef PK_GRAD_LAUNCH
        PK_CHECK_LAUNCH(); count_launch();
        if (dlogits_colsum) {
            colsum_partials_kernel<<<(ldv + 31) / 32, 256, 0, stream>>>(cs_part, ggrid, ldv, dlogits_colsum);
            PK_CHECK_LAUNCH(); count_launch();
        }
    }
    return 0;
}

extern "C" int pk_rnnt_loss_fwd_bwd(const void* logits, int dtype, const int* labels, const int* frame_lens,
                                    const int* label_lens, int B, int T, int U1, int V, int ldv, int ld_labels,
                                    const float* grad_scale, float* costs, void* dlogits, float* dlogits_colsum, void* workspace,
                                    long long workspace_bytes, void* stream) {
    return rnnt_loss_impl(logits, dtype, labels, frame_lens, label_lens, B, T, U1, V, ldv, ld_labels, grad_scale, costs, dlogits,
                          dlogits_colsum, workspace, workspace_bytes, nullptr, 0, stream);
}

extern "C" int pk_rnnt_loss_fwd_bwd_lse(const void* logits, int dtype, const int* labels, const int* frame_lens,
                                        const int* label_lens, int B, int T, int U1, int V, int ldv, int ld_labels,
                                        const float* grad_scale, float* costs, void* dlogits, float* dlogits_colsum, void* workspace,
                                        long long workspace_bytes, const float* row_lse, int n_parts, void* stream) {
    PK_CHECK_ARG(row_lse != nullptr, "row_lse is null");
    return rnnt_loss_impl(logits, dtype, labels, frame_lens, label_lens, B, T, U1, V, ldv, ld_labels, grad_scale, costs, dlogits,
                          dlogits_colsum, workspace, workspace_bytes, row_lse, n_parts, stream);
}
