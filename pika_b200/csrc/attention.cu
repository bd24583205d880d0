// Fused multi-head self-attention for the Transformer encoder layers (unmasked, head dim 64, bf16):
//     O = dropout(softmax((Q / sqrt(d)) K^T)) V            (reference: trainer/model/modules/multi_headed_attn.py:199-223)
// without ever writing the [B, heads, T, T] score / probability tensors to HBM.  At the bench shape the
// materialised path moved ~4 GB per layer and direction through HBM (fp32 scores + bf16 probabilities);
// here each CTA keeps a 128-row slab of one head stationary in registers and streams 64-row tiles of the
// other operand through shared memory (cp.async double buffer), with the online-softmax recurrence on the
// accumulator fragments.
//
// Three instantiations of one skeleton (rows = the stationary index, tiles = the streamed index):
//   MODE 0  forward          rows = queries   S = Q K^T -> P -> O += P V ;  writes O and the row log-sum-exp
//   MODE 1  backward, dQ     rows = queries   recomputes P, dP = dO V^T, dS = P o (dP - D), dQ += dS K
//   MODE 2  backward, dK/dV  rows = keys      works on S^T: dV += P^T dO, dK += dS^T Q
// D_i = sum_d dO_id O_id comes from a small pre-pass.  Recomputing S in both backward kernels costs two extra
// T x T x 64 products per head but keeps every output owned by exactly one CTA (no atomics, deterministic).
//
// The products run on mma.sync.m16n8k16 (HMMA) rather than tcgen05: the tiles are 16 x 64 per warp with a
// softmax between the two products of every tile, which is the register-resident pattern; attention is ~2 % of
// the step's FLOPs, the point of the kernel is the HBM traffic it removes.
//
// Dropout uses the same counter-based mask as the stand-alone softmax kernels (index = (row * T + col) over the
// [B*heads*T, T] probability matrix), so both paths drop the same elements for a given seed.
#include "../../include/pika_b200.h"
#include "common.cuh"

namespace pk {
void count_launch();

constexpr int AT_ROWS = 128;      // stationary rows per CTA: 8 warps x 16
constexpr int AT_TILE = 64;       // streamed rows per iteration
constexpr int AT_THREADS = 256;
constexpr float AT_LOG2E = 1.4426950408889634f;
constexpr float AT_LN2 = 0.6931471805599453f;

struct AttnParams {
    const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;   // element (b, t, h, d) at base[(b*T + t)*ld + h*64 + d]
    long long ld_qkv;
    const __nv_bfloat16* o; long long ld_o;          // forward output (read by the backward pre-pass)
    const __nv_bfloat16* dout; long long ld_do;
    __nv_bfloat16* out;                              // MODE 0
    __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv; long long ld_dqkv;
    float* lse;                                      // [B*heads*T] natural-log row log-sum-exp of the scaled scores
    float* dsum;                                     // [B*heads*T] D_i
    int B, T, heads;
    int stat_ld;                                     // row pitch of lse / dsum per (batch, head): pk_attention_lse_stride(T)
    float alpha;
    uint32_t drop_thresh; float drop_scale; uint32_t seed;
};

PK_DEVICE void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
PK_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
PK_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

PK_DEVICE void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
PK_DEVICE void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
PK_DEVICE void hmma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A tile is rows x 128 B (64 bf16); the eight 16-byte chunks of a row are XOR-swizzled with the row index so
// that both ldmatrix flavours are bank-conflict free.
PK_DEVICE uint32_t tile_addr(uint32_t base, int row, int chunk) { return base + row * 128 + ((chunk ^ (row & 7)) << 4); }

// rows [r0, r0 + NROWS) of one head of a [B, T, ld] matrix -> swizzled tile; rows >= T are zero-filled
template <int NROWS>
PK_DEVICE void load_tile(uint32_t sbase, const __nv_bfloat16* g, long long ld, int r0, int T) {
    for (int c = threadIdx.x; c < NROWS * 8; c += AT_THREADS) {
        const int r = c >> 3, ch = c & 7;
        const int gr = r0 + r;
        const int grc = gr < T ? gr : T - 1;
        cp_async16(tile_addr(sbase, r, ch), g + (long long)grc * ld + ch * 8, gr < T ? 16 : 0);
    }
}

// A fragments (4 k-steps over d = 64) of this warp's 16 stationary rows
PK_DEVICE void load_a_frags(uint32_t sbase, int warp, int lane, uint32_t (&a)[4][4]) {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(tile_addr(sbase, row, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[nt] (16 x 8 each, nt = 0..7 over the 64 streamed rows) += A (16 x 64) * X^T, X tile stored [streamed row][d]
PK_DEVICE void mma_a_xt(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t xbase, int lane) {
    const int rsel = (lane & 7) + ((lane >> 4) & 1) * 8;
    const int csel = (lane >> 3) & 1;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(tile_addr(xbase, np * 16 + rsel, ks * 2 + csel), b0, b1, b2, b3);
            hmma_16816(acc[2 * np], a[ks], b0, b1);
            hmma_16816(acc[2 * np + 1], a[ks], b2, b3);
        }
    }
}

// acc[dt] (16 x 8 each, dt = 0..7 over d = 64) += P (16 x 64 streamed, as A fragments) * X, X tile stored [streamed row][d]
PK_DEVICE void mma_p_x(float (&acc)[8][4], const uint32_t (&p)[4][4], uint32_t xbase, int lane) {
    const int rsel = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int csel = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(tile_addr(xbase, ks * 16 + rsel, dp * 2 + csel), b0, b1, b2, b3);
            hmma_16816(acc[2 * dp], p[ks], b0, b1);
            hmma_16816(acc[2 * dp + 1], p[ks], b2, b3);
        }
    }
}

// accumulator-layout 16 x 64 tile -> A fragments of the next product (bf16)
PK_DEVICE void acc_to_a(const float (&x)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a[ks][0] = pack_bf16x2(x[2 * ks][0], x[2 * ks][1]);
        a[ks][1] = pack_bf16x2(x[2 * ks][2], x[2 * ks][3]);
        a[ks][2] = pack_bf16x2(x[2 * ks + 1][0], x[2 * ks + 1][1]);
        a[ks][3] = pack_bf16x2(x[2 * ks + 1][2], x[2 * ks + 1][3]);
    }
}

PK_DEVICE float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
PK_DEVICE float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// 16 x 64 accumulator tile -> rows (row0 + g, row0 + g + 8), columns h*64 + ..., scaled, bf16
PK_DEVICE void store_rows(__nv_bfloat16* base, long long ld, int row0, int T, int lane, const float (&acc)[8][4], float s_lo, float s_hi) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = row0 + g + half * 8;
        if (r >= T) continue;
        const float sc = half ? s_hi : s_lo;
        uint32_t* rp = reinterpret_cast<uint32_t*>(base + (long long)r * ld);
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) rp[dt * 4 + t] = pack_bf16x2(acc[dt][half * 2] * sc, acc[dt][half * 2 + 1] * sc);
    }
}

template <int MODE>
__global__ void __launch_bounds__(AT_THREADS, 1) attention_kernel(const AttnParams p) {
    // two stages x (X1 tile, X2 tile) of 8 KB each; the first 32 KB double as the staging area of the stationary slabs
    __shared__ __align__(1024) uint8_t smem[2 * 2 * AT_TILE * 128];
    __shared__ float s_lse[2][AT_TILE], s_dsum[2][AT_TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int T = p.T;
    const int bh = blockIdx.y, b = bh / p.heads, h = bh - b * p.heads;
    const int row_base = blockIdx.x * AT_ROWS;                 // first stationary row of this CTA
    const long long head_off = (long long)b * T * p.ld_qkv + h * 64;
    const __nv_bfloat16* gq = p.q + head_off;
    const __nv_bfloat16* gk = p.k + head_off;
    const __nv_bfloat16* gv = p.v + head_off;
    const __nv_bfloat16* gdo = (MODE == 0) ? nullptr : p.dout + (long long)b * T * p.ld_do + h * 64;
    const long long stat_row0 = (long long)bh * T;             // row offset into the dropout index space
    const long long stat_vec0 = (long long)bh * p.stat_ld;     // row offset into lse / dsum
    const uint32_t sbase = smem_u32(smem);

    // ---- stationary fragments: A1 (Q | Q | K), A2 (- | dO | V)
    uint32_t a1[4][4], a2[4][4];
    {
        load_tile<AT_ROWS>(sbase, MODE == 2 ? gk : gq, p.ld_qkv, row_base, T);
        if (MODE != 0) load_tile<AT_ROWS>(sbase + AT_ROWS * 128, MODE == 2 ? gv : gdo, MODE == 2 ? p.ld_qkv : p.ld_do, row_base, T);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        load_a_frags(sbase, warp, lane, a1);
        if (MODE != 0) load_a_frags(sbase + AT_ROWS * 128, warp, lane, a2);
        __syncthreads();
    }
    const int my_row_lo = row_base + warp * 16 + g, my_row_hi = my_row_lo + 8;   // the two stationary rows this thread's fragments cover
    const float c2 = p.alpha * AT_LOG2E;                       // scores -> log2 domain

    float acc1[8][4], acc2[8][4];                              // MODE 0: O, -   MODE 1: dQ, -   MODE 2: dK, dV
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc1[i][e] = 0.f; acc2[i][e] = 0.f; }
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;           // MODE 0 online softmax state
    float lse_lo = 0.f, lse_hi = 0.f, d_lo = 0.f, d_hi = 0.f;                   // MODE 1 row scalars (log2 domain lse)
    if (MODE == 1) {
        if (my_row_lo < T) { lse_lo = p.lse[stat_vec0 + my_row_lo] * AT_LOG2E; d_lo = p.dsum[stat_vec0 + my_row_lo]; }
        if (my_row_hi < T) { lse_hi = p.lse[stat_vec0 + my_row_hi] * AT_LOG2E; d_hi = p.dsum[stat_vec0 + my_row_hi]; }
    }

    const __nv_bfloat16* gx1 = (MODE == 2) ? gq : gk;          // streamed operand of the score product
    const __nv_bfloat16* gx2 = (MODE == 2) ? gdo : gv;
    const long long ld_x2 = (MODE == 2) ? p.ld_do : p.ld_qkv;
    const int n_tiles = (T + AT_TILE - 1) / AT_TILE;

    auto issue = [&](int j) {
        const uint32_t st = sbase + (j & 1) * (2 * AT_TILE * 128);
        load_tile<AT_TILE>(st, gx1, p.ld_qkv, j * AT_TILE, T);
        load_tile<AT_TILE>(st + AT_TILE * 128, gx2, ld_x2, j * AT_TILE, T);
        if (MODE == 2 && threadIdx.x < AT_TILE) {
            const int r = j * AT_TILE + threadIdx.x;
            s_lse[j & 1][threadIdx.x] = r < T ? p.lse[stat_vec0 + r] * AT_LOG2E : 0.f;
            s_dsum[j & 1][threadIdx.x] = r < T ? p.dsum[stat_vec0 + r] : 0.f;
        }
        cp_async_commit();
    };
    issue(0);

    for (int j = 0; j < n_tiles; ++j) {
        cp_async_wait_all();
        __syncthreads();                                       // tile j landed; everyone is done with tile j-1's buffer
        if (j + 1 < n_tiles) issue(j + 1);
        const uint32_t x1 = sbase + (j & 1) * (2 * AT_TILE * 128), x2 = x1 + AT_TILE * 128;
        const int col0 = j * AT_TILE;                          // first streamed index of this tile

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
        mma_a_xt(s, a1, x1, lane);

        uint32_t pa[4][4];
        if (MODE == 0) {
            // ---- online softmax over the streamed keys
            float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool ok = col0 + nt * 8 + 2 * t + e < T;
                    s[nt][e] = ok ? s[nt][e] * c2 : -INFINITY;
                    s[nt][2 + e] = ok ? s[nt][2 + e] * c2 : -INFINITY;
                    mx_lo = fmaxf(mx_lo, s[nt][e]);
                    mx_hi = fmaxf(mx_hi, s[nt][2 + e]);
                }
            const float mn_lo = fmaxf(m_lo, quad_max(mx_lo)), mn_hi = fmaxf(m_hi, quad_max(mx_hi));   // finite: every tile has a valid column
            const float cr_lo = ex2_approx(m_lo - mn_lo), cr_hi = ex2_approx(m_hi - mn_hi);
            m_lo = mn_lo; m_hi = mn_hi;
            float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float plo = ex2_approx(s[nt][e] - mn_lo), phi = ex2_approx(s[nt][2 + e] - mn_hi);
                    sum_lo += plo; sum_hi += phi;
                    if (p.drop_thresh) {
                        const uint64_t col = (uint64_t)(col0 + nt * 8 + 2 * t + e);
                        if (!drop_keep((uint64_t)(stat_row0 + my_row_lo) * (uint64_t)T + col, p.seed, p.drop_thresh)) plo = 0.f;
                        if (!drop_keep((uint64_t)(stat_row0 + my_row_hi) * (uint64_t)T + col, p.seed, p.drop_thresh)) phi = 0.f;
                    }
                    s[nt][e] = plo; s[nt][2 + e] = phi;
                }
            l_lo = l_lo * cr_lo + sum_lo;
            l_hi = l_hi * cr_hi + sum_hi;
#pragma unroll
            for (int dt = 0; dt < 8; ++dt) { acc1[dt][0] *= cr_lo; acc1[dt][1] *= cr_lo; acc1[dt][2] *= cr_hi; acc1[dt][3] *= cr_hi; }
            acc_to_a(s, pa);
            mma_p_x(acc1, pa, x2, lane);                       // O += P V
        } else {
            float dp[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) dp[i][e] = 0.f;
            mma_a_xt(dp, a2, x2, lane);                        // MODE 1: dO V^T   MODE 2: V dO^T
            float pd[8][4];                                    // MODE 2 only: dropped probabilities for dV
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = col0 + nt * 8 + 2 * t + (e & 1);          // streamed index
                    const int row = (e & 2) ? my_row_hi : my_row_lo;          // stationary index
                    const int cl = nt * 8 + 2 * t + (e & 1);
                    float lse2, dsum;
                    if (MODE == 1) { lse2 = (e & 2) ? lse_hi : lse_lo; dsum = (e & 2) ? d_hi : d_lo; }
                    else { lse2 = s_lse[j & 1][cl]; dsum = s_dsum[j & 1][cl]; }
                    float pr = (col < T && row < T) ? ex2_approx(fmaf(s[nt][e], c2, -lse2)) : 0.f;
                    float dpe = dp[nt][e];
                    bool keep = true;
                    if (p.drop_thresh) {
                        const uint64_t qi = (uint64_t)(MODE == 1 ? row : col), kj = (uint64_t)(MODE == 1 ? col : row);
                        keep = drop_keep(((uint64_t)stat_row0 + qi) * (uint64_t)T + kj, p.seed, p.drop_thresh);
                        dpe = keep ? dpe * p.drop_scale : 0.f;
                    }
                    if (MODE == 2) pd[nt][e] = keep ? pr * p.drop_scale : 0.f;
                    s[nt][e] = pr * (dpe - dsum);              // dS (or dS^T)
                }
            if (MODE == 2) {
                acc_to_a(pd, pa);
                mma_p_x(acc2, pa, x2, lane);                   // dV += Pd^T dO
            }
            acc_to_a(s, pa);
            mma_p_x(acc1, pa, x1, lane);                       // MODE 1: dQ += dS K    MODE 2: dK += dS^T Q
        }
    }

    // ---- epilogue
    const int wrow0 = row_base + warp * 16;
    if (MODE == 0) {
        l_lo = quad_sum(l_lo); l_hi = quad_sum(l_hi);
        store_rows(p.out + (long long)b * T * p.ld_o + h * 64, p.ld_o, wrow0, T, lane, acc1, p.drop_scale / l_lo, p.drop_scale / l_hi);
        if (t == 0) {
            if (my_row_lo < T) p.lse[stat_vec0 + my_row_lo] = (m_lo + log2f(l_lo)) * AT_LN2;
            if (my_row_hi < T) p.lse[stat_vec0 + my_row_hi] = (m_hi + log2f(l_hi)) * AT_LN2;
        }
    } else if (MODE == 1) {
        store_rows(p.dq + (long long)b * T * p.ld_dqkv + h * 64, p.ld_dqkv, wrow0, T, lane, acc1, p.alpha, p.alpha);
    } else {
        store_rows(p.dk + (long long)b * T * p.ld_dqkv + h * 64, p.ld_dqkv, wrow0, T, lane, acc1, p.alpha, p.alpha);
        store_rows(p.dv + (long long)b * T * p.ld_dqkv + h * 64, p.ld_dqkv, wrow0, T, lane, acc2, 1.f, 1.f);
    }
}

// D[(b*heads + h)*T + t] = sum_d dO[b,t,h,d] * O[b,t,h,d]; one warp per (b, t, h)
__global__ void __launch_bounds__(256) attention_rowdot_kernel(const __nv_bfloat16* __restrict__ o, long long ld_o,
                                                               const __nv_bfloat16* __restrict__ dout, long long ld_do, float* __restrict__ dsum,
                                                               int B, int T, int heads, int stat_ld) {
    const int lane = threadIdx.x & 31;
    const long long n = (long long)B * T * heads;
    for (long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); w < n; w += (long long)gridDim.x * 8) {
        const int h = (int)(w % heads);
        const long long bt = w / heads;
        const int tt = (int)(bt % T);
        const int b = (int)(bt / T);
        const uint32_t ov = *reinterpret_cast<const uint32_t*>(o + bt * ld_o + h * 64 + lane * 2);
        const uint32_t dv = *reinterpret_cast<const uint32_t*>(dout + bt * ld_do + h * 64 + lane * 2);
        const float s = warp_sum(bf16lo(ov) * bf16lo(dv) + bf16hi(ov) * bf16hi(dv));
        if (lane == 0) dsum[((long long)b * heads + h) * stat_ld + tt] = s;
    }
}

static uint32_t attn_drop_thresh(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 4294967296.0;
    uint32_t r = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    return r == 0 ? 1u : r;
}

}  // namespace pk

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define DONE() PK_CHECK_LAUNCH(); count_launch(); return 0

#define ATTN_CHECKS()                                                                                                      \
    PK_CHECK_ARG(B > 0 && T > 0 && heads > 0, "bad dims");                                                                 \
    PK_CHECK_ARG(dh == 64, "fused attention supports head dim 64");                                                        \
    PK_CHECK_ARG(ld_qkv % 8 == 0 && ld_out % 8 == 0, "row strides must be multiples of 8 elements (16 bytes)");            \
    PK_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "drop_p out of range");                                                    \
    PK_CHECK_ARG((long long)B * heads < 65536, "B * heads must be < 65536")

extern "C" int pk_attention_hmma_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out, float* lse,
                                int B, int T, int heads, int dh, float alpha, float drop_p, uint32_t seed, void* stream) {
    using namespace pk;
    ATTN_CHECKS();
    AttnParams p{};
    p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.ld_qkv = ld_qkv;
    p.out = (__nv_bfloat16*)out; p.ld_o = ld_out; p.lse = lse;
    p.B = B; p.T = T; p.heads = heads; p.alpha = alpha; p.stat_ld = (T + 63) / 64 * 64;
    p.drop_thresh = attn_drop_thresh(drop_p); p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f; p.seed = seed;
    dim3 grid((T + AT_ROWS - 1) / AT_ROWS, B * heads);
    attention_kernel<0><<<grid, AT_THREADS, 0, STREAM(stream)>>>(p);
    DONE();
}

/* dq/dk/dv share the row stride ld_dqkv (the fused [B,T,3D] gradient buffer); dsum_ws: B*heads*T floats of scratch */
extern "C" int pk_attention_hmma_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                                const void* dout, long long ld_dout, const float* lse, float* dsum_ws, void* dq, void* dk, void* dv,
                                long long ld_dqkv, int B, int T, int heads, int dh, float alpha, float drop_p, uint32_t seed, void* stream) {
    using namespace pk;
    ATTN_CHECKS();
    PK_CHECK_ARG(ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "row strides must be multiples of 8 elements (16 bytes)");
    AttnParams p{};
    p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.ld_qkv = ld_qkv;
    p.o = (const __nv_bfloat16*)out; p.ld_o = ld_out; p.dout = (const __nv_bfloat16*)dout; p.ld_do = ld_dout;
    p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv; p.ld_dqkv = ld_dqkv;
    p.lse = const_cast<float*>(lse); p.dsum = dsum_ws;
    p.B = B; p.T = T; p.heads = heads; p.alpha = alpha; p.stat_ld = (T + 63) / 64 * 64;
    p.drop_thresh = attn_drop_thresh(drop_p); p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f; p.seed = seed;
    const long long n = (long long)B * T * heads;
    const int rgrid = (int)((n + 7) / 8 < 148ll * 16 ? (n + 7) / 8 : 148ll * 16);
    attention_rowdot_kernel<<<rgrid, 256, 0, STREAM(stream)>>>(p.o, ld_out, p.dout, ld_dout, dsum_ws, B, T, heads, p.stat_ld);
    PK_CHECK_LAUNCH(); count_launch();
    dim3 grid((T + AT_ROWS - 1) / AT_ROWS, B * heads);
    attention_kernel<1><<<grid, AT_THREADS, 0, STREAM(stream)>>>(p);
    PK_CHECK_LAUNCH(); count_launch();
    attention_kernel<2><<<grid, AT_THREADS, 0, STREAM(stream)>>>(p);
    DONE();
}
