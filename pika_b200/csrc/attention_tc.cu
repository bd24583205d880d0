// Fused multi-head self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head dim 64, bf16:
//     O = dropout(softmax((Q / sqrt(d)) K^T)) V            (reference: trainer/model/modules/multi_headed_attn.py:199-223)
// and its backward, without ever writing the [B, heads, T, T] score / probability tensors to HBM.
//
// One CTA = 128 "stationary" rows of one (batch, head) x all 64-row "streamed" tiles of the other operand:
//   MODE 0  forward          stationary = queries   S = Q K^T -> P -> O += P V ;  writes O and the row log-sum-exp
//   MODE 1  backward, dQ     stationary = queries   recomputes P, dP = dO V^T, dS = P o (dP - D), dQ += dS K
//   MODE 2  backward, dK/dV  stationary = keys      works on S^T: dV += P^T dO, dK += dS^T Q
// (every output row has one owner: no atomics, bit-reproducible).
//
// Warp roles (320 threads):
//   warp 0       TMA loader: stationary tiles once, streamed tiles (and, in MODE 2, the per-query lse / D vectors) through a
//                6-stage mbarrier ring
//   warp 1       MMA issuer (one thread): stage 1  S = A_stat X1^T (and dP = A_stat' X2^T)   tcgen05.mma 128 x 64 x 64, fp32 in TMEM
//                                         stage 2  acc (+)= A_elem X   where A_elem is the bf16 tile the element-wise warps wrote to
//                                                  shared memory (P | dS | P^T, dS^T) and X the streamed tile read MN-major
//   warps 2..5   element-wise group 0: even tiles      one thread per stationary row: tcgen05.ld of its row of S (dP), softmax /
//   warps 6..9   element-wise group 1: odd tiles       dropout / dS in registers, bf16 row -> 128B-swizzled A tile -> stage 2
// The two groups ping-pong, so the exponentials of tile j+1 overlap the tensor-core work of tile j; each group owns one S / dP
// accumulator in TMEM and one set of A tiles.  The forward keeps the running (max, sum, O) per row in registers (each PV product
// lands in a fresh TMEM accumulator and is folded in one tile later), and the two groups' partial softmax states are merged at
// the end.  The backward accumulates dQ / dK / dV in TMEM over all tiles and reads them once.
//
// Dropout: counter-based mask shared with the stand-alone softmax kernels -- one 32-bit hash per PAIR of adjacent keys
// (index (row * ceil(T/2) + key/2) over the [B*heads*T, T] probability matrix), 16 bits per element.
#include <cstdlib>
#include <cstring>

#include "../../include/pika_b200.h"
#include "common.cuh"

namespace pk {
void count_launch();

constexpr int TA_BR = 128;
constexpr int TA_BC = 64;
constexpr int TA_NST = 6;
constexpr int TA_THREADS = 320;
constexpr float TA_LOG2E = 1.4426950408889634f;
constexpr float TA_LN2 = 0.6931471805599453f;

constexpr int TA_TILE_S = TA_BR * 128;            // 16 KB: 128 rows x 64 bf16
constexpr int TA_TILE_X = TA_BC * 128;            // 8 KB
constexpr int TA_OFF_STAT = 0;                    // two stationary tiles
constexpr int TA_OFF_RING = 2 * TA_TILE_S;        // NST x (X1, X2)
constexpr int TA_OFF_ABUF = TA_OFF_RING + TA_NST * 2 * TA_TILE_X;      // [2 groups][2] A tiles of 16 KB
constexpr int TA_OFF_VEC = TA_OFF_ABUF + 4 * TA_TILE_S;                // [NST][2][64] floats
constexpr int TA_OFF_SALT = TA_OFF_VEC + TA_NST * 2 * TA_BC * 4;         // [2 groups][2][64] dropout row salts of the streamed queries (MODE 2)
constexpr int TA_OFF_BAR = TA_OFF_SALT + 2 * 2 * TA_BC * 4;
constexpr int TA_SMEM = TA_OFF_BAR + 256 + 1024;
constexpr int TA_SCR_LD = 68;                     // floats per row of the merge scratch (aliases the ring)
static_assert(TA_BR * TA_SCR_LD * 4 <= TA_NST * 2 * TA_TILE_X, "merge scratch must fit in the ring");

// TMEM columns (fp32): S[2] | dP[2] | second-stage accumulators
constexpr int TA_COL_S = 0, TA_COL_DP = 128, TA_COL_ACC = 256;

struct AttnTcParams {
    CUtensorMap q, k, v, dout;                    // [B][T][heads*64] bf16, box 64 x 64 x 1, 128B swizzle
    __nv_bfloat16* out; long long ld_o;
    __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv; long long ld_dqkv;
    float* lse;                                   // [B*heads][Tpad] natural-log row log-sum-exp of the scaled scores
    float* dsum;                                  // [B*heads][Tpad] D_i = sum_d dO_id O_id
    int B, T, heads, Tpad, Tp2;
    float alpha;
    uint32_t thresh16; float drop_scale; uint32_t seed;
};

template <int MODE>
__global__ void __launch_bounds__(TA_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TA_OFF_BAR);
    uint64_t* stat_full = bars;               // 1
    uint64_t* full_bar = bars + 1;            // NST
    uint64_t* empty_bar = full_bar + TA_NST;  // NST
    uint64_t* s_ready = empty_bar + TA_NST;   // 2   stage-1 products of a group's tile are in TMEM
    uint64_t* a_ready = s_ready + 2;          // 2   the group wrote its A tile(s) and is done reading S / dP
    uint64_t* pv_done = a_ready + 2;          // 2   stage 2 of the group's tile completed (A tiles free, PV readable)
    uint64_t* final_bar = pv_done + 2;        // 1   every MMA of the CTA completed
    uint64_t* s_free = final_bar + 1;         // 2   the group has pulled S / dP of its tile into registers: the accumulator may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
    float* vec = reinterpret_cast<float*>(smem + TA_OFF_VEC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int bh = blockIdx.y, b = bh / p.heads, h = bh - b * p.heads;
    const int row_base = blockIdx.x * TA_BR;
    const int n_tiles = (T + TA_BC - 1) / TA_BC;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.q); tma_prefetch_desc(&p.k); tma_prefetch_desc(&p.v);
        if (MODE != 0) tma_prefetch_desc(&p.dout);
        mbar_init(stat_full, 1);
        for (int s = 0; s < TA_NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(&s_ready[g], 1); mbar_init(&a_ready[g], 128); mbar_init(&pv_done[g], 1); mbar_init(&s_free[g], 128); }
        mbar_init(final_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA loader (converged warp, one elected lane issues)
        const uint32_t lead = elect_one_u32();
        const CUtensorMap* ma0 = (MODE == 2) ? &p.k : &p.q;
        const CUtensorMap* ma1 = (MODE == 2) ? &p.v : &p.dout;
        const CUtensorMap* mx1 = (MODE == 2) ? &p.q : &p.k;
        const CUtensorMap* mx2 = (MODE == 2) ? &p.dout : &p.v;
        mbar_arrive_expect_tx_p(stat_full, (MODE == 0 ? 1 : 2) * TA_TILE_S, lead);
        for (int half = 0; half < 2; ++half) {
            tma_load_3d_p(smem + TA_OFF_STAT + half * TA_TILE_X, ma0, stat_full, h * 64, row_base + half * 64, b, lead);
            if (MODE != 0) tma_load_3d_p(smem + TA_OFF_STAT + TA_TILE_S + half * TA_TILE_X, ma1, stat_full, h * 64, row_base + half * 64, b, lead);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* x1 = smem + TA_OFF_RING + stage * 2 * TA_TILE_X;
            mbar_arrive_expect_tx_p(&full_bar[stage], 2 * TA_TILE_X + (MODE == 2 ? 2 * TA_BC * 4 : 0), lead);
            tma_load_3d_p(x1, mx1, &full_bar[stage], h * 64, j * TA_BC, b, lead);
            tma_load_3d_p(x1 + TA_TILE_X, mx2, &full_bar[stage], h * 64, j * TA_BC, b, lead);
            if (MODE == 2) {
                const size_t off = (size_t)bh * p.Tpad + (size_t)j * TA_BC;
                bulk_load_p(vec + (stage * 2 + 0) * TA_BC, p.lse + off, TA_BC * 4, &full_bar[stage], lead);
                bulk_load_p(vec + (stage * 2 + 1) * TA_BC, p.dsum + off, TA_BC * 4, &full_bar[stage], lead);
            }
            if (++stage == TA_NST) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (converged warp, one elected lane issues)
        const uint32_t lead = elect_one_u32();
        constexpr uint32_t idesc1 = make_idesc_bf16(TA_BR, TA_BC, false, false);   // A K-major, B K-major (both over d)
        constexpr uint32_t idesc2 = make_idesc_bf16(TA_BR, 64, false, true);       // A K-major (over the streamed index), B MN-major
        const uint32_t sbase = smem_u32(smem);
        const uint64_t kdesc = make_smem_desc_sw128(0, 16, 1024);                  // K-major tile: rows of 128 B
        const uint64_t mdesc = make_smem_desc_sw128(0, 8192, 1024);                // MN-major tile: 8-row atoms along K
        // descriptor of a tile = template + (address >> 4); a k16 step adds a compile-time constant to the 14-bit address field (tiles are
        // 1024-byte aligned and the whole buffer lies below 256 KB, so the field never carries): one 64-bit add per operand per MMA
        auto kmaj = [&](uint32_t addr, int k4) { return kdesc + (uint64_t)((addr >> 4) & 0x3FFF) + (uint64_t)(k4 * 2); };
        auto mnmaj = [&](uint32_t addr, int k4) { return mdesc + (uint64_t)((addr >> 4) & 0x3FFF) + (uint64_t)(k4 * 128); };
        mbar_wait(stat_full, 0);
        tc_fence_after();
        const uint32_t st0 = sbase + TA_OFF_STAT, st1 = st0 + TA_TILE_S;
        auto stage2 = [&](int t) {
            const int g = t & 1, s = t % TA_NST;
            mbar_wait(&a_ready[g], (t >> 1) & 1);
            tc_fence_after();
            const uint32_t x1 = sbase + TA_OFF_RING + s * 2 * TA_TILE_X, x2 = x1 + TA_TILE_X;
            const uint32_t a0 = sbase + TA_OFF_ABUF + (g * 2) * TA_TILE_S, a1 = a0 + TA_TILE_S;
#pragma unroll
            for (int k4 = 0; k4 < TA_BC / 16; ++k4) {
                if (MODE == 0) umma_bf16_p(tmem_base + TA_COL_ACC + g * 64, kmaj(a0, k4), mnmaj(x2, k4), idesc2, k4 > 0 ? 1u : 0u, lead);
                if (MODE == 1) umma_bf16_p(tmem_base + TA_COL_ACC, kmaj(a0, k4), mnmaj(x1, k4), idesc2, (t > 0 || k4 > 0) ? 1u : 0u, lead);
                if (MODE == 2) {
                    umma_bf16_p(tmem_base + TA_COL_ACC + 64, kmaj(a0, k4), mnmaj(x2, k4), idesc2, (t > 0 || k4 > 0) ? 1u : 0u, lead);   // dV += Pd^T dO
                    umma_bf16_p(tmem_base + TA_COL_ACC, kmaj(a1, k4), mnmaj(x1, k4), idesc2, (t > 0 || k4 > 0) ? 1u : 0u, lead);        // dK += dS^T Q
                }
            }
            umma_commit_p(&pv_done[g], lead);
            umma_commit_p(&empty_bar[s], lead);
        };
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_tiles; ++j) {
            const int g = j & 1;
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t x1 = sbase + TA_OFF_RING + stage * 2 * TA_TILE_X, x2 = x1 + TA_TILE_X;
            // S (and dP) of group g may be overwritten as soon as the group has pulled tile j-2 into registers (s_free), long before it
            // finishes the exponentials of that tile
            if (j >= 2) { mbar_wait(&s_free[g], ((j >> 1) - 1) & 1); tc_fence_after(); }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) umma_bf16_p(tmem_base + TA_COL_S + g * 64, kmaj(st0, k4), kmaj(x1, k4), idesc1, k4 > 0 ? 1u : 0u, lead);
            if (MODE != 0) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) umma_bf16_p(tmem_base + TA_COL_DP + g * 64, kmaj(st1, k4), kmaj(x2, k4), idesc1, k4 > 0 ? 1u : 0u, lead);
            }
            umma_commit_p(&s_ready[g], lead);
            // second stage of tile j-2, i.e. AFTER the stage-1 products of the next tile of the same group have been queued: the
            // issue order S0 S1 S2 P0 S3 P1 ... keeps a finished S waiting for each group when it comes back from its exponentials
            if (j >= 2) stage2(j - 2);
            if (++stage == TA_NST) { stage = 0; phase ^= 1; }
        }
        if (n_tiles >= 2) stage2(n_tiles - 2);
        stage2(n_tiles - 1);
        umma_commit_p(final_bar, lead);
    } else {
        // ===================================================== element-wise groups
        const int g = (warp - 2) >> 2;
        const int qd = warp & 3;                               // TMEM lane quarter of this warp
        const int r = qd * 32 + lane;                          // stationary row inside the CTA
        const int srow = row_base + r;                         // stationary index (query in MODE 0/1, key in MODE 2)
        const uint32_t tq = tmem_base + ((uint32_t)(qd * 32) << 16);
        const uint64_t stat_row0 = (uint64_t)bh * (uint64_t)T; // row offset into the dropout index space
        const float c2 = p.alpha * TA_LOG2E;
        uint8_t* a0 = smem + TA_OFF_ABUF + (g * 2) * TA_TILE_S + r * 128;
        uint8_t* a1 = a0 + TA_TILE_S;
        const int sw = r & 7;
        const bool use_drop = p.thresh16 != 0u;
        const uint32_t row_salt = drop_row_salt(stat_row0 + (uint64_t)srow, p.seed);     // MODE 0/1: the probability row is this thread's row

        float o_acc[64];                                       // MODE 0: running O of this group's tiles (scale = m_run)
        float m_run = -INFINITY, l_run = 0.f;
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 64; ++c) o_acc[c] = 0.f;
        }
        float lse2 = 0.f, dsum_r = 0.f;                        // MODE 1 row scalars
        if (MODE == 1 && srow < T) {
            lse2 = p.lse[(size_t)bh * p.Tpad + srow] * TA_LOG2E;
            dsum_r = p.dsum[(size_t)bh * p.Tpad + srow];
        }

        int it = 0;
        for (int j = g; j < n_tiles; j += 2, ++it) {
            const int col0 = j * TA_BC;
            const int nvalid = min(TA_BC, T - col0);           // streamed indices of this tile inside the sequence
            mbar_wait(&s_ready[g], it & 1);
            tc_fence_after();
            if (MODE == 0) {
                if (it > 0) {
                    // the PV product of this group's previous tile: same scale as o_acc (both relative to m_run); folded before S is
                    // pulled so that the two 32/64-register blocks are never live together
                    mbar_wait(&pv_done[g], (it - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t pv[32];
                        tmem_ld_32x32(tq + TA_COL_ACC + g * 64 + half * 32, pv);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) o_acc[half * 32 + c] += __uint_as_float(pv[c]);
                    }
                }
                uint32_t sr[64];
                tmem_ld_32x32(tq + TA_COL_S + g * 64, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
                tmem_ld_32x32(tq + TA_COL_S + g * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_free[g]);
                if (nvalid < TA_BC) {                          // last tile only (uniform): keys beyond the sequence score -inf -> probability 0
#pragma unroll
                    for (int c = 0; c < 64; ++c) if (c >= nvalid) sr[c] = 0xff800000u;
                }
                float mx = -INFINITY;
#pragma unroll
                for (int c = 0; c < 64; ++c) mx = fmaxf(mx, __uint_as_float(sr[c]));
                const float m_new = fmaxf(m_run, mx * c2);     // finite: every tile holds a valid column
                const float corr = ex2_approx(m_run - m_new);  // 0 on the group's first tile
#pragma unroll
                for (int c = 0; c < 64; ++c) o_acc[c] *= corr;
                float sum = 0.f;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    float pr[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = ch * 8 + e;
                        pr[e] = ex2_approx(fmaf(__uint_as_float(sr[c]), c2, -m_new));
                        sum += pr[e];
                    }
                    if (use_drop) {
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const uint32_t km = drop_pair(row_salt, (uint32_t)((col0 + ch * 8) >> 1) + e2, p.thresh16);
                            if (!(km & 1u)) pr[2 * e2] = 0.f;
                            if (!(km & 2u)) pr[2 * e2 + 1] = 0.f;
                        }
                    }
                    *reinterpret_cast<uint4*>(a0 + ((ch ^ sw) << 4)) =
                        make_uint4(pack_bf16x2(pr[0], pr[1]), pack_bf16x2(pr[2], pr[3]), pack_bf16x2(pr[4], pr[5]), pack_bf16x2(pr[6], pr[7]));
                }
                l_run = l_run * corr + sum;
                m_run = m_new;
            } else {
                if (it > 0) mbar_wait(&pv_done[g], (it - 1) & 1);       // the A tiles of this group are free again
                const float* lse_t = vec + ((j % TA_NST) * 2 + 0) * TA_BC;
                const float* dsum_t = lse_t + TA_BC;
                uint32_t* qsalt_t = reinterpret_cast<uint32_t*>(smem + TA_OFF_SALT) + (g * 2 + (it & 1)) * TA_BC;
                if (MODE == 2 && use_drop) {
                    // the probability rows of this tile are the streamed queries: one salt per query, shared by the group (double
                    // buffered by tile parity, so a group barrier per tile is all the ordering it needs)
                    if (r < TA_BC) qsalt_t[r] = drop_row_salt(stat_row0 + (uint64_t)(col0 + r), p.seed);
                    named_bar_sync(2 + g, 128);
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t sr[32], dr[32];
                    tmem_ld_32x32(tq + TA_COL_S + g * 64 + half * 32, sr);
                    tmem_ld_32x32(tq + TA_COL_DP + g * 64 + half * 32, dr);
                    tmem_ld_wait();
                    if (half == 1) { tc_fence_before(); mbar_arrive(&s_free[g]); }       // S / dP of this tile are in registers
                    if (nvalid < TA_BC) {                      // last tile only (uniform): streamed indices beyond the sequence get probability 0
#pragma unroll                                                 // (their dP is 0 and the lse / D pads are zero-initialised, so dS = 0 * finite)
                        for (int c = 0; c < 32; ++c) if (half * 32 + c >= nvalid) sr[c] = 0xff800000u;
                    }
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        float ds[8], pd[8];
                        uint32_t keep = 0xFFu;
                        if (use_drop) {
                            if (MODE == 1) {
#pragma unroll
                                for (int e2 = 0; e2 < 4; ++e2) {
                                    const uint32_t km = drop_pair(row_salt, (uint32_t)((col0 + half * 32 + ch * 8) >> 1) + e2, p.thresh16);
                                    if (!(km & 1u)) keep &= ~(1u << (2 * e2));
                                    if (!(km & 2u)) keep &= ~(2u << (2 * e2));
                                }
                            } else {
                                // probability row = streamed query, key = this thread's row: pairs of keys sit in adjacent lanes; a lane hashes the
                                // columns of its own parity and trades with its neighbour
#pragma unroll
                                for (int e2 = 0; e2 < 4; ++e2) {
                                    const int c_mine = half * 32 + ch * 8 + 2 * e2 + (lane & 1), c_other = c_mine ^ 1;
                                    const uint32_t km = drop_pair(qsalt_t[c_mine], (uint32_t)(srow >> 1), p.thresh16);
                                    const uint32_t ko = __shfl_xor_sync(0xffffffffu, km, 1);
                                    const uint32_t bit = (lane & 1) ? 2u : 1u;
                                    const int e_mine = c_mine & 7, e_other = c_other & 7;
                                    if (!(km & bit)) keep &= ~(1u << e_mine);
                                    if (!(ko & bit)) keep &= ~(1u << e_other);
                                }
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int cl = half * 32 + ch * 8 + e;             // streamed index inside the tile
                            const float l2 = (MODE == 1) ? lse2 : lse_t[cl] * TA_LOG2E;
                            const float dd = (MODE == 1) ? dsum_r : dsum_t[cl];
                            const float pr = ex2_approx(fmaf(__uint_as_float(sr[ch * 8 + e]), c2, -l2));
                            const bool kp = (keep >> e) & 1u;
                            const float dpe = kp ? __uint_as_float(dr[ch * 8 + e]) * p.drop_scale : 0.f;
                            ds[e] = pr * (dpe - dd);
                            if (MODE == 2) pd[e] = kp ? pr * p.drop_scale : 0.f;
                        }
                        const int chunk = half * 4 + ch;
                        if (MODE == 1) {
                            *reinterpret_cast<uint4*>(a0 + ((chunk ^ sw) << 4)) =
                                make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7]));
                        } else {
                            *reinterpret_cast<uint4*>(a0 + ((chunk ^ sw) << 4)) =
                                make_uint4(pack_bf16x2(pd[0], pd[1]), pack_bf16x2(pd[2], pd[3]), pack_bf16x2(pd[4], pd[5]), pack_bf16x2(pd[6], pd[7]));
                            *reinterpret_cast<uint4*>(a1 + ((chunk ^ sw) << 4)) =
                                make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7]));
                        }
                    }
                }
            }
            tc_fence_before();                 // the tcgen05.ld of S / dP above are complete (wait::ld) and ordered before the arrive
            fence_proxy_async_smem();          // generic-proxy writes of the A tile -> visible to the tensor core's async proxy
            mbar_arrive(&a_ready[g]);
        }

        // ---- epilogue
        mbar_wait(final_bar, 0);
        tc_fence_after();
        if (MODE == 0) {
            if (it > 0) {                      // fold the last PV product of this group (already complete: final_bar)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t pv[32];
                    tmem_ld_32x32(tq + TA_COL_ACC + g * 64 + half * 32, pv);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) o_acc[half * 32 + c] += __uint_as_float(pv[c]);
                }
            }
            float* scr = reinterpret_cast<float*>(smem + TA_OFF_RING) + r * TA_SCR_LD;
            if (g == 1) {
#pragma unroll
                for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(scr + c) = make_float4(o_acc[c], o_acc[c + 1], o_acc[c + 2], o_acc[c + 3]);
                scr[64] = m_run; scr[65] = l_run;
            }
            named_bar_sync(1, 256);
            if (g == 0) {
                const float m1 = scr[64], l1 = scr[65];
                const float mm = fmaxf(m_run, m1);             // group 0 always has tile 0: finite
                const float e0 = ex2_approx(m_run - mm), e1 = ex2_approx(m1 - mm);
                const float l = l_run * e0 + l1 * e1;
                const float inv = p.drop_scale / l;
                if (srow < T) {
                    __nv_bfloat16* orow = p.out + ((long long)b * T + srow) * p.ld_o + h * 64;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = (o_acc[ch * 8 + e] * e0 + scr[ch * 8 + e] * e1) * inv;
                        *reinterpret_cast<uint4*>(orow + ch * 8) =
                            make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                    }
                    p.lse[(size_t)bh * p.Tpad + srow] = (mm + log2f(l)) * TA_LN2;
                }
            }
        } else {
            // MODE 1: group 0 stores dQ columns 0..31, group 1 columns 32..63.  MODE 2: group 0 stores dK, group 1 dV.
            const int ncol = (MODE == 1) ? 32 : 64;
            const uint32_t tcol = TA_COL_ACC + ((MODE == 1) ? g * 32 : g * 64);
            __nv_bfloat16* base = (MODE == 1) ? p.dq : (g == 0 ? p.dk : p.dv);
            const float sc = (MODE == 1 || g == 0) ? p.alpha : 1.f;
            __nv_bfloat16* orow = base + ((long long)b * T + srow) * p.ld_dqkv + h * 64 + ((MODE == 1) ? g * 32 : 0);
#pragma unroll
            for (int part = 0; part < ncol / 32; ++part) {
                uint32_t v[32];
                tmem_ld_32x32(tq + tcol + part * 32, v);
                tmem_ld_wait();
                if (srow < T) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[ch * 8 + e]) * sc;
                        *reinterpret_cast<uint4*>(orow + part * 32 + ch * 8) =
                            make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// D[(b*heads + h)*Tpad + t] = sum_d dO[b,t,h,d] * O[b,t,h,d]; one warp per (b, t, h)
__global__ void __launch_bounds__(256) attention_rowdot_tc_kernel(const __nv_bfloat16* __restrict__ o, long long ld_o,
                                                                  const __nv_bfloat16* __restrict__ dout, long long ld_do, float* __restrict__ dsum,
                                                                  int B, int T, int heads, int Tpad) {
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
        if (lane == 0) dsum[((long long)b * heads + h) * Tpad + tt] = s;
    }
}

static int attn_maps(AttnTcParams& p, const void* q, const void* k, const void* v, long long ld_qkv, const void* dout, long long ld_dout) {
    const unsigned long long dims[3] = {(unsigned long long)p.heads * 64, (unsigned long long)p.T, (unsigned long long)p.B};
    const unsigned box[3] = {64, TA_BC, 1};
    const unsigned long long st_qkv[2] = {(unsigned long long)ld_qkv * 2, (unsigned long long)ld_qkv * 2 * p.T};
    int rc;
    if ((rc = encode_tiled_bf16_3d(&p.q, q, dims, st_qkv, box, "attention q"))) return rc;
    if ((rc = encode_tiled_bf16_3d(&p.k, k, dims, st_qkv, box, "attention k"))) return rc;
    if ((rc = encode_tiled_bf16_3d(&p.v, v, dims, st_qkv, box, "attention v"))) return rc;
    if (dout) {
        const unsigned long long st_do[2] = {(unsigned long long)ld_dout * 2, (unsigned long long)ld_dout * 2 * p.T};
        if ((rc = encode_tiled_bf16_3d(&p.dout, dout, dims, st_do, box, "attention dout"))) return rc;
    }
    return 0;
}

template <int MODE> static int launch_attn(const AttnTcParams& p, cudaStream_t st) {
    auto kern = attention_tc_kernel<MODE>;
    static bool configured = false;
    if (!configured) {
        PK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        configured = true;
    }
    dim3 grid((p.T + TA_BR - 1) / TA_BR, p.B * p.heads);
    kern<<<grid, TA_THREADS, TA_SMEM, st>>>(p);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}

}  // namespace pk

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int pk_attention_lse_stride(int T) { return (T + 63) / 64 * 64; }

#define ATTN_CHECKS()                                                                                                      \
    PK_CHECK_ARG(B > 0 && T > 0 && heads > 0, "bad dims");                                                                 \
    PK_CHECK_ARG(dh == 64, "fused attention supports head dim 64");                                                        \
    PK_CHECK_ARG(ld_qkv % 8 == 0 && ld_out % 8 == 0, "row strides must be multiples of 8 elements (16 bytes)");            \
    PK_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "drop_p out of range");                                                    \
    PK_CHECK_ARG((long long)B * heads < 65536, "B * heads must be < 65536")

extern "C" int pk_attention_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out, float* lse,
                                int B, int T, int heads, int dh, float alpha, float drop_p, uint32_t seed, void* stream) {
    using namespace pk;
    ATTN_CHECKS();
    static thread_local AttnTcParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.heads = heads; p.Tpad = pk_attention_lse_stride(T); p.Tp2 = (T + 1) / 2;
    int rc = attn_maps(p, q, k, v, ld_qkv, nullptr, 0);
    if (rc) return rc;
    p.out = (__nv_bfloat16*)out; p.ld_o = ld_out; p.lse = lse;
    p.alpha = alpha;
    p.thresh16 = drop_thresh16_of(drop_p); p.drop_scale = drop_scale16_of(p.thresh16); p.seed = seed;
    return launch_attn<0>(p, STREAM(stream));
}

/* dq/dk/dv share the row stride ld_dqkv (the fused [B,T,3D] gradient buffer); lse and dsum_ws: [B*heads][pk_attention_lse_stride(T)] */
extern "C" int pk_attention_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                                const void* dout, long long ld_dout, const float* lse, float* dsum_ws, void* dq, void* dk, void* dv,
                                long long ld_dqkv, int B, int T, int heads, int dh, float alpha, float drop_p, uint32_t seed, void* stream) {
    using namespace pk;
    ATTN_CHECKS();
    PK_CHECK_ARG(ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "row strides must be multiples of 8 elements (16 bytes)");
    static thread_local AttnTcParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.heads = heads; p.Tpad = pk_attention_lse_stride(T); p.Tp2 = (T + 1) / 2;
    int rc = attn_maps(p, q, k, v, ld_qkv, dout, ld_dout);
    if (rc) return rc;
    p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv; p.ld_dqkv = ld_dqkv;
    p.lse = const_cast<float*>(lse); p.dsum = dsum_ws;
    p.alpha = alpha;
    p.thresh16 = drop_thresh16_of(drop_p); p.drop_scale = drop_scale16_of(p.thresh16); p.seed = seed;
    const long long n = (long long)B * T * heads;
    const int rgrid = (int)((n + 7) / 8 < 148ll * 16 ? (n + 7) / 8 : 148ll * 16);
    attention_rowdot_tc_kernel<<<rgrid, 256, 0, STREAM(stream)>>>((const __nv_bfloat16*)out, ld_out, (const __nv_bfloat16*)dout, ld_dout, dsum_ws,
                                                                   B, T, heads, p.Tpad);
    PK_CHECK_LAUNCH(); count_launch();
    rc = launch_attn<1>(p, STREAM(stream));
    if (rc) return rc;
    return launch_attn<2>(p, STREAM(stream));
}
