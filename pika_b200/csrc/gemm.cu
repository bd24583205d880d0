// tcgen05 / TMEM / TMA GEMM for sm_100a: C = epilogue(alpha * sum_p A_p * B_p^T).
//
// One persistent CTA per SM, 192 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      MMA issuer     (one thread: tcgen05.mma kind::f16, fp32 accumulators in TMEM,
//                               tcgen05.commit frees smem stages / publishes the accumulator)
//   warps 2..5  epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> swizzled smem
//                               -> TMA store; double-buffered TMEM accumulators overlap it with
//                               the next tile's main loop)
// Tile 128 x BN x 64 (BN = 64 | 128 | 256).  Operands may be K-major or MN-major (wgrad / dgrad /
// P.V use the MN-major form so no transposes are ever materialised).  Up to 9 (A,B) pairs
// accumulate into one tile (TDNN taps, split-bf16 fp32-class mode), plus a batched reduction
// loop (kz) for per-utterance wgrad.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/pika_b200.h"
#include "common.cuh"

namespace pk {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int C_STAGE_BYTES = BM * 128;      // 128 rows x 128 B
constexpr int GEMM_THREADS = 192;

struct GemmParams {
    CUtensorMap a[PK_GEMM_MAX_PAIRS];
    CUtensorMap b[PK_GEMM_MAX_PAIRS];
    CUtensorMap c;
    int a_off[PK_GEMM_MAX_PAIRS];
    int b_off[PK_GEMM_MAX_PAIRS];
    int n_pairs, kz_count, num_k_blocks;
    int k_splits, iters_per_split;      // split-K over the flattened (pair, kz, k-block) iteration space
    int M, N, tiles_m, tiles_n, zb0, zb1;
    int a_sel2, a_sel3, b_sel2, b_sel3;
    int c_is_f32, c_accumulate;
    float alpha;
    const float* bias;
    int act;
    uint32_t drop_thresh;
    float drop_scale;
    uint32_t drop_seed;
    int aux_mode, aux_is_f32;
    const void* aux;
    long long aux_sm, aux_s0, aux_s1;
    float aux_scale;
    float* row_lse;                     // optional [tiles_n][M][2] per-row (max*log2e, sum 2^(x*log2e-max)) partials of the bf16 output
};

template <int BN> struct GemmCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int C_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = C_OFF + 2 * C_STAGE_BYTES;
    static constexpr int BAR_OFF = BIAS_OFF + BN * 4;
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;   // + barriers + alignment slack
    static constexpr int TMEM_COLS = 2 * BN;                  // double-buffered accumulator (>= 32, pow2)
};

PK_DEVICE int pick_sel(int sel, int zb0, int zb1, int kz) {
    return sel == PK_SEL_ZB0 ? zb0 : (sel == PK_SEL_ZB1 ? zb1 : (sel == PK_SEL_KZ ? kz : 0));
}

// EPI selects what the epilogue compiles in, so that the common case is straight-line code (the run-time-uniform
// branches of the full epilogue cost an instruction-fetch bubble per 16-byte group, profiles/r01_notes.md):
//   EPI_PLAIN  alpha, bias, ReLU only      EPI_LSE  + per-row log-sum-exp partials (bf16 C)      EPI_FULL  + dropout / aux add / aux mask
// The epilogue pulls 32 accumulator columns per tcgen05.ld (one wait per 32 columns).
enum { EPI_PLAIN = 0, EPI_LSE = 1, EPI_FULL = 2 };
template <bool A_MN, bool B_MN, int BN, bool CF32, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* empty_bar = full_bar + Cfg::STAGES;
    uint64_t* tmem_full = empty_bar + Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* bias_smem = reinterpret_cast<float*>(smem + Cfg::BIAS_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.n_pairs; ++i) {
            tma_prefetch_desc(&p.a[i]);
            tma_prefetch_desc(&p.b[i]);
        }
        tma_prefetch_desc(&p.c);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_z = p.tiles_m * p.tiles_n;
    const int num_tiles = tiles_per_z * p.zb0 * p.zb1 * p.k_splits;      // work units = tiles x K-splits
    const int k_iters_total = p.n_pairs * p.kz_count * p.num_k_blocks;
    const int kzb = p.kz_count * p.num_k_blocks;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x) {
                const int tile = unit / p.k_splits, split = unit - tile * p.k_splits;
                const int z = tile / tiles_per_z;
                const int r = tile - z * tiles_per_z;
                const int mb = r / p.tiles_n, nb = r - mb * p.tiles_n;
                const int zb1 = z / p.zb0, zb0 = z - zb1 * p.zb0;
                const int m0 = mb * BM, n0 = nb * BN;
                const int i0 = split * p.iters_per_split, i1 = min(k_iters_total, i0 + p.iters_per_split);
                if (p.n_pairs == 1 && p.kz_count == 1) {
                    // One (A, B) pair and no batched reduction (every forward / dgrad / plain wgrad GEMM): the flattened index IS the
                    // k-block and the selectors are tile constants.  The general loop below spends ~160 SASS instructions per k-block
                    // in this single thread (two integer divisions, four selector ladders) -- more than the 4 MMAs of a stage take --
                    // so the common case gets a loop that only waits, arms the barrier and issues the two TMA loads.
                    const int a2 = pick_sel(p.a_sel2, zb0, zb1, 0), a3 = pick_sel(p.a_sel3, zb0, zb1, 0);
                    const int b2 = pick_sel(p.b_sel2, zb0, zb1, 0), b3 = pick_sel(p.b_sel3, zb0, zb1, 0);
                    const int a_off = p.a_off[0], b_off = p.b_off[0];
                    for (int kb = i0; kb < i1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sb = sa + A_STAGE_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        if (A_MN) {
#pragma unroll
                            for (int c = 0; c < BM / 64; ++c)
                                tma_load_4d(sa + c * (64 * BK * 2), &p.a[0], &full_bar[stage], m0 + c * 64, kb * BK + a_off, a2, a3);
                        } else {
                            tma_load_4d(sa, &p.a[0], &full_bar[stage], kb * BK, m0 + a_off, a2, a3);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int c = 0; c < BN / 64; ++c)
                                tma_load_4d(sb + c * (64 * BK * 2), &p.b[0], &full_bar[stage], n0 + c * 64, kb * BK + b_off, b2, b3);
                        } else {
                            tma_load_4d(sb, &p.b[0], &full_bar[stage], kb * BK, n0 + b_off, b2, b3);
                        }
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    }
                    continue;
                }
                for (int i = i0; i < i1; ++i) {
                    const int pr = i / kzb;
                    const int rem = i - pr * kzb;
                    const int kz = rem / p.num_k_blocks, kb = rem - kz * p.num_k_blocks;
                    const int a2 = pick_sel(p.a_sel2, zb0, zb1, kz), a3 = pick_sel(p.a_sel3, zb0, zb1, kz);
                    const int b2 = pick_sel(p.b_sel2, zb0, zb1, kz), b3 = pick_sel(p.b_sel3, zb0, zb1, kz);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (A_MN) {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c)
                            tma_load_4d(sa + c * (64 * BK * 2), &p.a[pr], &full_bar[stage], m0 + c * 64, kb * BK + p.a_off[pr], a2, a3);
                    } else {
                        tma_load_4d(sa, &p.a[pr], &full_bar[stage], kb * BK, m0 + p.a_off[pr], a2, a3);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int c = 0; c < BN / 64; ++c)
                            tma_load_4d(sb + c * (64 * BK * 2), &p.b[pr], &full_bar[stage], n0 + c * 64, kb * BK + p.b_off[pr], b2, b3);
                    } else {
                        tma_load_4d(sb, &p.b[pr], &full_bar[stage], kb * BK, n0 + p.b_off[pr], b2, b3);
                    }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x, ++it) {
                const int split = unit % p.k_splits;
                const int k_iters = min(k_iters_total, (split + 1) * p.iters_per_split) - split * p.iters_per_split;
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 16; ++k4) {
                        // K-major: 16 elements = 32 B inside the 128 B swizzle row; atoms of 8 rows (1024 B).
                        // MN-major: 16 k-rows = two 8-row atoms (2048 B); 64-wide MN chunks 8192 B apart.
                        const uint64_t ad = A_MN ? make_smem_desc_sw128(sa + k4 * 2048, 64 * BK * 2, 1024)
                                                 : make_smem_desc_sw128(sa + k4 * 32, 16, 1024);
                        const uint64_t bd = B_MN ? make_smem_desc_sw128(sb + k4 * 2048, 64 * BK * 2, 1024)
                                                 : make_smem_desc_sw128(sb + k4 * 32, 16, 1024);
                        umma_bf16(d_tmem, ad, bd, idesc, (k > 0 || k4 > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ===================================================== epilogue (warps 2..5)
        // Compact loop over 16-byte output groups (8 bf16 / 4 f32 columns): one small tcgen05.ld per
        // group keeps the body a few dozen instructions, so it stays resident in the instruction cache.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // tile row owned by this thread
        const int et = threadIdx.x - 64;        // 0..127
        const bool store_thread = (et == 0);
        constexpr int GW = CF32 ? 4 : 8;        // columns per 16-byte group
        constexpr int CH = 8 * GW;              // columns per 128-byte staging row
        uint8_t* cst = smem + Cfg::C_OFF;
        const float relu_floor = (p.act == PK_ACT_RELU) ? 0.f : -INFINITY;
        if (p.bias == nullptr) {
            for (int j = et; j < BN; j += 128) bias_smem[j] = 0.f;
            named_bar_sync(2, 128);
        }
        int it = 0;
        uint32_t chunk_ctr = 0;
        for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x, ++it) {
            const int tile = unit / p.k_splits;
            const int z = tile / tiles_per_z;
            const int r = tile - z * tiles_per_z;
            const int mb = r / p.tiles_n, nb = r - mb * p.tiles_n;
            const int zb1 = z / p.zb0, zb0 = z - zb1 * p.zb0;
            const int m0 = mb * BM, n0 = nb * BN;
            const int acc = it & 1;
            const int m = m0 + row;
            if (p.bias != nullptr) {
                named_bar_sync(2, 128);         // previous tile's readers are done with bias_smem
                for (int j = et; j < BN; j += 128) bias_smem[j] = (n0 + j < p.N) ? p.bias[n0 + j] : 0.f;
                named_bar_sync(2, 128);
            }
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            const bool row_ok = m < p.M;
            const unsigned char* aux_row = nullptr;
            if (p.aux_mode != PK_AUX_NONE && row_ok) {
                const long long off = (long long)m * p.aux_sm + (long long)zb0 * p.aux_s0 + (long long)zb1 * p.aux_s1;
                aux_row = reinterpret_cast<const unsigned char*>(p.aux) + off * (p.aux_is_f32 ? 4 : 2);
            }
            const uint64_t lin_row = ((uint64_t)(zb1 * p.zb0 + zb0) * (uint64_t)p.M + (uint64_t)m) * (uint64_t)p.N;
            float lse_m = -INFINITY, lse_s = 0.f;   // running row max (log2 units) and sum over this tile's columns
            constexpr int n_chunks = BN / CH;
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int nc0 = n0 + ch * CH;
                if (nc0 >= p.N) break;           // uniform across the 4 epilogue warps
                uint8_t* sbuf = cst + (chunk_ctr & 1) * C_STAGE_BYTES;
                if (store_thread) tma_store_wait_read<1>();     // the buffer used two chunks ago is free
                named_bar_sync(1, 128);
                uint8_t* srow = sbuf + row * 128;
                auto do_group = [&](const uint32_t (&rr)[GW], const int gq) {
                    const int ncol = nc0 + gq * GW;
                    float x[GW];
                    const float* bsm = bias_smem + ch * CH + gq * GW;
#pragma unroll
                    for (int e = 0; e < GW; ++e) x[e] = fmaxf(fmaf(__uint_as_float(rr[e]), p.alpha, bsm[e]), relu_floor);
                    if (EPI == EPI_FULL && p.drop_thresh != 0u) {
#pragma unroll
                        for (int e = 0; e < GW; ++e)
                            x[e] = drop_keep(lin_row + (uint64_t)(ncol + e), p.drop_seed, p.drop_thresh) ? x[e] * p.drop_scale : 0.f;
                    }
                    if (EPI == EPI_FULL && aux_row != nullptr && ncol < p.N) {       // N % GW == 0 is enforced on the host when aux is used
                        float a[GW];
                        if (p.aux_is_f32) {
                            const float4* ap = reinterpret_cast<const float4*>(aux_row + (size_t)ncol * 4);
#pragma unroll
                            for (int e4 = 0; e4 < GW / 4; ++e4) {
                                const float4 t4 = ap[e4];
                                a[e4 * 4 + 0] = t4.x; a[e4 * 4 + 1] = t4.y; a[e4 * 4 + 2] = t4.z; a[e4 * 4 + 3] = t4.w;
                            }
                        } else {
                            if (GW == 8) {
                                const uint4 t4 = *reinterpret_cast<const uint4*>(aux_row + (size_t)ncol * 2);
                                a[0] = bf16lo(t4.x); a[1] = bf16hi(t4.x); a[2] = bf16lo(t4.y); a[3] = bf16hi(t4.y);
                                a[GW - 4] = bf16lo(t4.z); a[GW - 3] = bf16hi(t4.z); a[GW - 2] = bf16lo(t4.w); a[GW - 1] = bf16hi(t4.w);
                            } else {
                                const uint2 t2 = *reinterpret_cast<const uint2*>(aux_row + (size_t)ncol * 2);
                                a[0] = bf16lo(t2.x); a[1] = bf16hi(t2.x); a[2] = bf16lo(t2.y); a[3] = bf16hi(t2.y);
                            }
                        }
                        if (p.aux_mode == PK_AUX_ADD) {
#pragma unroll
                            for (int e = 0; e < GW; ++e) x[e] += a[e];
                        } else {
#pragma unroll
                            for (int e = 0; e < GW; ++e) x[e] = (a[e] != 0.f) ? x[e] * p.aux_scale : 0.f;
                        }
                    }
                    uint4 w;
                    if (CF32) {
                        w = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
                    } else {
                        w.x = pack_bf16x2(x[0], x[1]); w.y = pack_bf16x2(x[2], x[3]);
                        w.z = pack_bf16x2(x[GW - 4], x[GW - 3]); w.w = pack_bf16x2(x[GW - 2], x[GW - 1]);
                        if (EPI == EPI_LSE) {
                            // online log-sum-exp over the ROUNDED values (what the consumer of C will read); N % 8 == 0, so a
                            // group is valid or invalid as a whole: invalid groups are pushed to -inf instead of branching
                            const float kill = (ncol < p.N) ? 0.f : -INFINITY;
                            float r[8];
                            r[0] = bf16lo(w.x) + kill; r[1] = bf16hi(w.x) + kill; r[2] = bf16lo(w.y) + kill; r[3] = bf16hi(w.y) + kill;
                            r[4] = bf16lo(w.z) + kill; r[5] = bf16hi(w.z) + kill; r[6] = bf16lo(w.w) + kill; r[7] = bf16hi(w.w) + kill;
                            const float gm = fmaxf(fmaxf(fmaxf(r[0], r[1]), fmaxf(r[2], r[3])), fmaxf(fmaxf(r[4], r[5]), fmaxf(r[6], r[7])));
                            const float m_new = fmaxf(lse_m, gm * 1.4426950408889634f);   // finite: the first group of a tile is valid
                            // raw MUFU.EX2 (no denormal fix-up sequence): eight independent exponentials issue back to back
                            float ex[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) ex[e] = ex2_approx(fmaf(r[e], 1.4426950408889634f, -m_new));
                            const float acc = ((ex[0] + ex[1]) + (ex[2] + ex[3])) + ((ex[4] + ex[5]) + (ex[6] + ex[7]));
                            lse_s = fmaf(lse_s, ex2_approx(lse_m - m_new), acc);
                            lse_m = m_new;
                        }
                    }
                    *reinterpret_cast<uint4*>(srow + ((gq ^ (row & 7)) << 4)) = w;
                };
#pragma unroll 1
                for (int part = 0; part < CH / 32; ++part) {
                    uint32_t r32[32];
                    tmem_ld_32x32(t_addr + ch * CH + part * 32, r32);
                    tmem_ld_wait();
#pragma unroll
                    for (int g4 = 0; g4 < 32 / GW; ++g4) {
                        uint32_t rr[GW];
#pragma unroll
                        for (int e = 0; e < GW; ++e) rr[e] = r32[g4 * GW + e];
                        do_group(rr, part * (32 / GW) + g4);
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (store_thread) {
                    if (p.c_accumulate) tma_reduce_add_4d(&p.c, sbuf, nc0, m0, zb0, zb1);
                    else tma_store_4d(&p.c, sbuf, nc0, m0, zb0, zb1);
                    tma_store_commit();
                }
                ++chunk_ctr;
            }
            if (!CF32 && EPI == EPI_LSE && row_ok)
                *reinterpret_cast<float2*>(p.row_lse + ((size_t)nb * (size_t)p.M + (size_t)m) * 2) = make_float2(lse_m, lse_s);
            // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (store_thread) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}


// ------------------------------------------------------------------------------------------ CTA-pair variant
// cta_group::2: two CTAs on one TPC cooperate on a 256 x 256 tile.  Each CTA stages its own 128 rows of A and
// HALF of B (128 of the 256 N-rows), so shared-memory and L2->SM traffic per CTA drop by a third and the ring
// deepens from 4 to 6 stages; the leader CTA's single thread issues tcgen05.mma.cta_group::2 for both, the
// accumulator rows of each CTA live in its own TMEM, and each CTA runs its own epilogue.
struct Gemm2Cfg {
    static constexpr int BN = 256;
    static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;          // this CTA's half of B
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = 6;
    static constexpr int C_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = C_OFF + 2 * C_STAGE_BYTES;
    static constexpr int BAR_OFF = BIAS_OFF + BN * 4;
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 2 * BN;
};

template <bool A_MN, bool B_MN, bool CF32>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_2sm_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = Gemm2Cfg;
    constexpr int BN = Cfg::BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* empty_bar = full_bar + Cfg::STAGES;
    uint64_t* tmem_full = empty_bar + Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* bias_smem = reinterpret_cast<float*>(smem + Cfg::BIAS_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.n_pairs; ++i) {
            tma_prefetch_desc(&p.a[i]);
            tma_prefetch_desc(&p.b[i]);
        }
        tma_prefetch_desc(&p.c);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);              // leader's expect_tx arrive; the peer's loads are tracked by byte count only
            mbar_init(&empty_bar[s], 1);             // tcgen05.commit multicast
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);             // tcgen05.commit multicast
            mbar_init(&tmem_empty[a], 8);            // 4 epilogue warps x 2 CTAs (used in the leader only)
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_m2 = (p.M + 255) / 256;
    const int tiles_per_z = tiles_m2 * p.tiles_n;
    const int num_tiles = tiles_per_z * p.zb0 * p.zb1 * p.k_splits;
    const int k_iters_total = p.n_pairs * p.kz_count * p.num_k_blocks;
    const int kzb = p.kz_count * p.num_k_blocks;

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = pair; unit < num_tiles; unit += num_pairs) {
                const int tile = unit / p.k_splits, split = unit - tile * p.k_splits;
                const int z = tile / tiles_per_z;
                const int r = tile - z * tiles_per_z;
                const int mb = r / p.tiles_n, nb = r - mb * p.tiles_n;
                const int zb1 = z / p.zb0, zb0 = z - zb1 * p.zb0;
                const int m0 = mb * 256 + (int)rank * BM, n0 = nb * BN + (int)rank * (BN / 2);
                const int i0 = split * p.iters_per_split, i1 = min(k_iters_total, i0 + p.iters_per_split);
                for (int i = i0; i < i1; ++i) {
                    const int pr = i / kzb;
                    const int rem = i - pr * kzb;
                    const int kz = rem / p.num_k_blocks, kb = rem - kz * p.num_k_blocks;
                    const int a2 = pick_sel(p.a_sel2, zb0, zb1, kz), a3 = pick_sel(p.a_sel3, zb0, zb1, kz);
                    const int b2 = pick_sel(p.b_sel2, zb0, zb1, kz), b3 = pick_sel(p.b_sel3, zb0, zb1, kz);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    if (A_MN) {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c)
                            tma_load_4d_2sm(sa + c * (64 * BK * 2), &p.a[pr], &full_bar[stage], m0 + c * 64, kb * BK + p.a_off[pr], a2, a3);
                    } else {
                        tma_load_4d_2sm(sa, &p.a[pr], &full_bar[stage], kb * BK, m0 + p.a_off[pr], a2, a3);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int c = 0; c < BN / 128; ++c)
                            tma_load_4d_2sm(sb + c * (64 * BK * 2), &p.b[pr], &full_bar[stage], n0 + c * 64, kb * BK + p.b_off[pr], b2, b3);
                    } else {
                        tma_load_4d_2sm(sb, &p.b[pr], &full_bar[stage], kb * BK, n0 + p.b_off[pr], b2, b3);
                    }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit = pair; unit < num_tiles; unit += num_pairs, ++it) {
                const int split = unit % p.k_splits;
                const int k_iters = min(k_iters_total, (split + 1) * p.iters_per_split) - split * p.iters_per_split;
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 16; ++k4) {
                        const uint64_t ad = A_MN ? make_smem_desc_sw128(sa + k4 * 2048, 64 * BK * 2, 1024)
                                                 : make_smem_desc_sw128(sa + k4 * 32, 16, 1024);
                        const uint64_t bd = B_MN ? make_smem_desc_sw128(sb + k4 * 2048, 64 * BK * 2, 1024)
                                                 : make_smem_desc_sw128(sb + k4 * 32, 16, 1024);
                        umma_bf16_2sm(d_tmem, ad, bd, idesc, (k > 0 || k4 > 0) ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty_bar[stage], 3);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tmem_full[acc], 3);
            }
        }
    } else {
        // ===================================================== epilogue (warps 2..5)
        // Compact loop over 16-byte output groups (8 bf16 / 4 f32 columns): one small tcgen05.ld per
        // group keeps the body a few dozen instructions, so it stays resident in the instruction cache.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // tile row owned by this thread
        const int et = threadIdx.x - 64;        // 0..127
        const bool store_thread = (et == 0);
        constexpr int GW = CF32 ? 4 : 8;        // columns per 16-byte group
        constexpr int CH = 8 * GW;              // columns per 128-byte staging row
        uint8_t* cst = smem + Cfg::C_OFF;
        const float relu_floor = (p.act == PK_ACT_RELU) ? 0.f : -INFINITY;
        if (p.bias == nullptr) {
            for (int j = et; j < BN; j += 128) bias_smem[j] = 0.f;
            named_bar_sync(2, 128);
        }
        int it = 0;
        uint32_t chunk_ctr = 0;
        for (int unit = pair; unit < num_tiles; unit += num_pairs, ++it) {
            const int tile = unit / p.k_splits;
            const int z = tile / tiles_per_z;
            const int r = tile - z * tiles_per_z;
            const int mb = r / p.tiles_n, nb = r - mb * p.tiles_n;
            const int zb1 = z / p.zb0, zb0 = z - zb1 * p.zb0;
            const int m0 = mb * 256 + (int)rank * BM, n0 = nb * BN;
            const int acc = it & 1;
            const int m = m0 + row;
            if (p.bias != nullptr) {
                named_bar_sync(2, 128);         // previous tile's readers are done with bias_smem
                for (int j = et; j < BN; j += 128) bias_smem[j] = (n0 + j < p.N) ? p.bias[n0 + j] : 0.f;
                named_bar_sync(2, 128);
            }
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            const bool row_ok = m < p.M;
            const unsigned char* aux_row = nullptr;
            if (p.aux_mode != PK_AUX_NONE && row_ok) {
                const long long off = (long long)m * p.aux_sm + (long long)zb0 * p.aux_s0 + (long long)zb1 * p.aux_s1;
                aux_row = reinterpret_cast<const unsigned char*>(p.aux) + off * (p.aux_is_f32 ? 4 : 2);
            }
            const uint64_t lin_row = ((uint64_t)(zb1 * p.zb0 + zb0) * (uint64_t)p.M + (uint64_t)m) * (uint64_t)p.N;
            constexpr int n_chunks = BN / CH;
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int nc0 = n0 + ch * CH;
                if (nc0 >= p.N) break;           // uniform across the 4 epilogue warps
                uint8_t* sbuf = cst + (chunk_ctr & 1) * C_STAGE_BYTES;
                if (store_thread) tma_store_wait_read<1>();     // the buffer used two chunks ago is free
                named_bar_sync(1, 128);
                uint8_t* srow = sbuf + row * 128;
#pragma unroll 2
                for (int gq = 0; gq < 8; ++gq) {
                    uint32_t rr[GW];
                    if (CF32) tmem_ld_32x4(t_addr + ch * CH + gq * GW, *reinterpret_cast<uint32_t (*)[4]>(&rr[0]));
                    else tmem_ld_32x8(t_addr + ch * CH + gq * GW, *reinterpret_cast<uint32_t (*)[8]>(&rr[0]));
                    tmem_ld_wait();
                    const int ncol = nc0 + gq * GW;
                    float x[GW];
                    const float* bsm = bias_smem + ch * CH + gq * GW;
#pragma unroll
                    for (int e = 0; e < GW; ++e) x[e] = fmaxf(fmaf(__uint_as_float(rr[e]), p.alpha, bsm[e]), relu_floor);
                    if (p.drop_thresh != 0u) {
#pragma unroll
                        for (int e = 0; e < GW; ++e)
                            x[e] = drop_keep(lin_row + (uint64_t)(ncol + e), p.drop_seed, p.drop_thresh) ? x[e] * p.drop_scale : 0.f;
                    }
                    if (aux_row != nullptr && ncol < p.N) {       // N % GW == 0 is enforced on the host when aux is used
                        float a[GW];
                        if (p.aux_is_f32) {
                            const float4* ap = reinterpret_cast<const float4*>(aux_row + (size_t)ncol * 4);
#pragma unroll
                            for (int e4 = 0; e4 < GW / 4; ++e4) {
                                const float4 t4 = ap[e4];
                                a[e4 * 4 + 0] = t4.x; a[e4 * 4 + 1] = t4.y; a[e4 * 4 + 2] = t4.z; a[e4 * 4 + 3] = t4.w;
                            }
                        } else {
                            if (GW == 8) {
                                const uint4 t4 = *reinterpret_cast<const uint4*>(aux_row + (size_t)ncol * 2);
                                a[0] = bf16lo(t4.x); a[1] = bf16hi(t4.x); a[2] = bf16lo(t4.y); a[3] = bf16hi(t4.y);
                                a[GW - 4] = bf16lo(t4.z); a[GW - 3] = bf16hi(t4.z); a[GW - 2] = bf16lo(t4.w); a[GW - 1] = bf16hi(t4.w);
                            } else {
                                const uint2 t2 = *reinterpret_cast<const uint2*>(aux_row + (size_t)ncol * 2);
                                a[0] = bf16lo(t2.x); a[1] = bf16hi(t2.x); a[2] = bf16lo(t2.y); a[3] = bf16hi(t2.y);
                            }
                        }
                        if (p.aux_mode == PK_AUX_ADD) {
#pragma unroll
                            for (int e = 0; e < GW; ++e) x[e] += a[e];
                        } else {
#pragma unroll
                            for (int e = 0; e < GW; ++e) x[e] = (a[e] != 0.f) ? x[e] * p.aux_scale : 0.f;
                        }
                    }
                    uint4 w;
                    if (CF32) {
                        w = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
                    } else {
                        w.x = pack_bf16x2(x[0], x[1]); w.y = pack_bf16x2(x[2], x[3]);
                        w.z = pack_bf16x2(x[GW - 4], x[GW - 3]); w.w = pack_bf16x2(x[GW - 2], x[GW - 1]);
                    }
                    *reinterpret_cast<uint4*>(srow + ((gq ^ (row & 7)) << 4)) = w;
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (store_thread) {
                    if (p.c_accumulate) tma_reduce_add_4d(&p.c, sbuf, nc0, m0, zb0, zb1);
                    else tma_store_4d(&p.c, sbuf, nc0, m0, zb0, zb1);
                    tma_store_commit();
                }
                ++chunk_ctr;
            }
            // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_remote(&tmem_empty[acc], 0); }
        }
        if (store_thread) tma_store_wait<0>();
    }


    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// Build a rank-4 tiled tensor map with 128B swizzle.  box0 * elem_size must be 128 bytes.
static int make_map(CUtensorMap* out, const pk_view4& v, int is_f32, int box0, int box1, const char* what) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point not available"); return -3; }
    const int es = is_f32 ? 4 : 2;
    cuuint64_t dims[4];
    cuuint64_t strides[3];
    cuuint32_t box[4] = {(cuuint32_t)box0, (cuuint32_t)box1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        if (v.dim[i] <= 0) { set_last_error("gemm %s: dim[%d]=%lld must be > 0", what, i, (long long)v.dim[i]); return -1; }
        dims[i] = (cuuint64_t)v.dim[i];
    }
    for (int i = 0; i < 3; ++i) {
        long long sb = (long long)v.stride[i] * es;
        if (v.dim[i + 1] == 1 && (sb <= 0 || (sb % 16) != 0)) sb = 16;   // unused dimension: any legal stride
        if (sb <= 0 || (sb % 16) != 0) {
            set_last_error("gemm %s: stride[%d]=%lld elements is not a positive multiple of 16 bytes", what, i,
                           (long long)v.stride[i]);
            return -1;
        }
        strides[i] = (cuuint64_t)sb;
    }
    if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) != 0) { set_last_error("gemm %s: base pointer not 16B aligned", what); return -1; }
    if ((cuuint64_t)box[1] > 256) { set_last_error("gemm %s: box too large", what); return -1; }
    CUresult r = enc(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     const_cast<void*>(v.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("gemm %s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r); return -3; }
    return 0;
}

void count_launch();

template <bool A_MN, bool B_MN, int BN, bool CF32, int EPI>
static int launch_gemm_e(const GemmParams& gp, int grid, cudaStream_t stream) {
    auto kern = gemm_tcgen05_kernel<A_MN, B_MN, BN, CF32, EPI>;
    static bool configured = false;
    if (!configured) {
        PK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::SMEM_BYTES));
        configured = true;
    }
    kern<<<grid, GEMM_THREADS, GemmCfg<BN>::SMEM_BYTES, stream>>>(gp);
    PK_CHECK_LAUNCH();
    count_launch();
    return 0;
}

template <bool A_MN, bool B_MN, int BN, bool CF32>
static int launch_gemm(const GemmParams& gp, int grid, cudaStream_t stream) {
    if (gp.row_lse != nullptr) {
        if (CF32) { set_last_error("gemm: row_lse needs a bf16 C"); return -1; }
        return launch_gemm_e<A_MN, B_MN, BN, false, EPI_LSE>(gp, grid, stream);
    }
    if (gp.drop_thresh != 0u || gp.aux_mode != PK_AUX_NONE) return launch_gemm_e<A_MN, B_MN, BN, CF32, EPI_FULL>(gp, grid, stream);
    return launch_gemm_e<A_MN, B_MN, BN, CF32, EPI_PLAIN>(gp, grid, stream);
}

template <bool A_MN, bool B_MN, bool CF32>
static int launch_gemm_2sm(const GemmParams& gp, int pairs, cudaStream_t stream) {
    auto kern = gemm_tcgen05_2sm_kernel<A_MN, B_MN, CF32>;
    static bool configured = false;
    if (!configured) {
        PK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg::SMEM_BYTES));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Gemm2Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, gp));
    count_launch();
    return 0;
}
static int dispatch_2sm(const GemmParams& gp, int a_mn, int b_mn, int pairs, cudaStream_t stream) {
#define PK_2SM(AM, BM_) (gp.c_is_f32 ? launch_gemm_2sm<AM, BM_, true>(gp, pairs, stream) : launch_gemm_2sm<AM, BM_, false>(gp, pairs, stream))
    if (!a_mn && !b_mn) return PK_2SM(false, false);
    if (!a_mn && b_mn) return PK_2SM(false, true);
    if (a_mn && !b_mn) return PK_2SM(true, false);
    return PK_2SM(true, true);
#undef PK_2SM
}

template <int BN>
static int dispatch_major(const GemmParams& gp, int a_mn, int b_mn, int grid, cudaStream_t stream) {
    if (gp.c_is_f32) {
        if (!a_mn && !b_mn) return launch_gemm<false, false, BN, true>(gp, grid, stream);
        if (!a_mn && b_mn) return launch_gemm<false, true, BN, true>(gp, grid, stream);
        if (a_mn && !b_mn) return launch_gemm<true, false, BN, true>(gp, grid, stream);
        return launch_gemm<true, true, BN, true>(gp, grid, stream);
    }
    if (!a_mn && !b_mn) return launch_gemm<false, false, BN, false>(gp, grid, stream);
    if (!a_mn && b_mn) return launch_gemm<false, true, BN, false>(gp, grid, stream);
    if (a_mn && !b_mn) return launch_gemm<true, false, BN, false>(gp, grid, stream);
    return launch_gemm<true, true, BN, false>(gp, grid, stream);
}

}  // namespace pk

extern "C" int pk_gemm_bf16(const pk_gemm_desc* d, void* stream_v) {
    using namespace pk;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    PK_CHECK_ARG(d != nullptr, "null descriptor");
    PK_CHECK_ARG(d->n_pairs >= 1 && d->n_pairs <= PK_GEMM_MAX_PAIRS, "n_pairs out of range");
    PK_CHECK_ARG(d->kz_count >= 1, "kz_count must be >= 1");
    PK_CHECK_ARG(d->c_dtype == PK_F32 || d->c_dtype == PK_BF16, "bad c_dtype");
    PK_CHECK_ARG(!d->c_accumulate || d->c_dtype == PK_F32, "c_accumulate needs an f32 C");
    PK_CHECK_ARG(d->drop_p >= 0.f && d->drop_p < 1.f, "drop_p out of range");
    const long long N = d->c.dim[0], M = d->c.dim[1];
    PK_CHECK_ARG(M > 0 && N > 0, "empty C");
    PK_CHECK_ARG(d->aux == nullptr || d->aux_mode == PK_AUX_NONE || (N % 8 == 0), "aux epilogue needs N % 8 == 0");
    int bn = d->block_n;
    if (bn == 0) bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    PK_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "block_n must be 64, 128 or 256");

    // CTA-pair kernel for the large tiles: 256 x 256 per pair (needs M > 128 so that the second CTA has rows)
    static int use_2sm = -1;
    if (use_2sm < 0) { const char* e = getenv("PK_GEMM_2SM"); use_2sm = e ? atoi(e) : 0; }   // off by default until it beats the single-CTA kernel (profiles/r01_notes.md)
    const int want_2sm = d->two_sm < 0 ? 0 : (d->two_sm > 0 ? 1 : use_2sm);
    const bool two_sm = want_2sm && bn == 256 && M > 128;

    {   // the tensor-map encoder is a driver-API call: make sure this host thread (e.g. an autograd worker) has the primary context bound
        static thread_local bool ctx_ready = false;
        if (!ctx_ready) { PK_CHECK_CUDA(cudaFree(nullptr)); ctx_ready = true; }
    }
    static thread_local GemmParams gp;   // ~2.6 KB; filled per call, copied into the launch
    memset(&gp, 0, sizeof(gp));
    long long K = d->a_mn_major ? d->a[0].dim[1] : d->a[0].dim[0];
    for (int i = 0; i < d->n_pairs; ++i) {
        const long long ka = d->a_mn_major ? d->a[i].dim[1] : d->a[i].dim[0];
        const long long kb = d->b_mn_major ? d->b[i].dim[1] : d->b[i].dim[0];
        PK_CHECK_ARG(ka == K && kb == K, "all pairs must share the reduction extent K");
        int rc = make_map(&gp.a[i], d->a[i], 0, 64, d->a_mn_major ? 64 : BM, "A");
        if (rc) return rc;
        rc = make_map(&gp.b[i], d->b[i], 0, 64, d->b_mn_major ? 64 : (two_sm ? bn / 2 : bn), "B");   // a CTA of a pair stages half of B
        if (rc) return rc;
        gp.a_off[i] = d->a_row_off[i];
        gp.b_off[i] = d->b_row_off[i];
    }
    {
        int rc = make_map(&gp.c, d->c, d->c_dtype == PK_F32, d->c_dtype == PK_F32 ? 32 : 64, BM, "C");
        if (rc) return rc;
    }
    gp.n_pairs = d->n_pairs;
    gp.kz_count = d->kz_count;
    gp.num_k_blocks = (int)((K + BK - 1) / BK);
    gp.M = (int)M;
    gp.N = (int)N;
    gp.tiles_m = (int)((M + BM - 1) / BM);
    gp.tiles_n = (int)((N + bn - 1) / bn);
    gp.zb0 = (int)d->c.dim[2];
    gp.zb1 = (int)d->c.dim[3];
    gp.a_sel2 = d->a_sel2; gp.a_sel3 = d->a_sel3; gp.b_sel2 = d->b_sel2; gp.b_sel3 = d->b_sel3;
    gp.c_is_f32 = d->c_dtype == PK_F32;
    gp.c_accumulate = d->c_accumulate;
    gp.alpha = d->alpha;
    gp.bias = d->bias;
    gp.act = d->act;
    if (d->drop_p > 0.f) {
        double t = (double)d->drop_p * 4294967296.0;
        gp.drop_thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
        if (gp.drop_thresh == 0) gp.drop_thresh = 1;
        gp.drop_scale = 1.f / (1.f - d->drop_p);
    }
    gp.drop_seed = d->drop_seed;
    gp.aux_mode = d->aux ? d->aux_mode : PK_AUX_NONE;
    gp.aux_is_f32 = d->aux_dtype == PK_F32;
    gp.aux = d->aux;
    gp.aux_sm = d->aux_stride[0]; gp.aux_s0 = d->aux_stride[1]; gp.aux_s1 = d->aux_stride[2];
    gp.aux_scale = d->aux_scale;
    gp.row_lse = d->row_lse;
    PK_CHECK_ARG(d->row_lse == nullptr || (d->c_dtype == PK_BF16 && N % 8 == 0 && gp.zb0 == 1 && gp.zb1 == 1 && !two_sm &&
                                           gp.drop_thresh == 0u && gp.aux_mode == PK_AUX_NONE),
                 "row_lse needs a 2-D bf16 C with N % 8 == 0, no dropout / aux, on the single-CTA kernel");

    const long long out_tiles = (long long)gp.tiles_m * gp.tiles_n * gp.zb0 * gp.zb1;
    // split-K: under-filled grids with a long reduction (wgrad, the LSTM's recurrent dgrad) are cut along the
    // flattened (pair, kz, k-block) axis; partial tiles are combined with TMA reduce-add into a zeroed f32 C.
    const int k_iters_total = gp.n_pairs * gp.kz_count * gp.num_k_blocks;
    int splits = 1;
    const bool plain_epi = d->bias == nullptr && d->act == PK_ACT_NONE && d->drop_p == 0.f && gp.aux_mode == PK_AUX_NONE;
    if (d->k_splits > 0) splits = d->k_splits;
    else if (gp.c_is_f32 && plain_epi && gp.zb0 == 1 && gp.zb1 == 1 && k_iters_total >= 16) {
        static int mode = -1;                            // tuning hook: PK_GEMM_SPLIT_MODE=0 (fill two waves) | 1 (least last-wave waste)
        if (mode < 0) { const char* e = getenv("PK_GEMM_SPLIT_MODE"); mode = e ? atoi(e) : 0; }   // measured: mode 0 = 91.2 ms/step, mode 1 = 94-97 (profiles/r01_notes.md)
        const int sms = num_sms();
        if (mode == 0 || mode == 2) {
            if (out_tiles < 2 * sms) {
                splits = (int)((2 * sms + out_tiles / 2) / out_tiles);
                if (splits > k_iters_total / 8) splits = k_iters_total / 8;
                if (splits > 64) splits = 64;
                if (mode == 2 && splits >= 1 && splits + 1 <= k_iters_total / 8) {
                    // one more split when it fills the last wave noticeably better (fc2 wgrad: 188 tiles, 2 -> 3 splits = 0.85 -> 0.95)
                    auto eff = [&](long long sp) { const long long u = out_tiles * sp; return (double)u / (double)(((u + sms - 1) / sms) * sms); };
                    if (eff(splits + 1) > eff(splits) + 0.05) ++splits;
                }
            }
        } else if (mode == 1 && out_tiles < 6 * sms) {
            // pick the split count (<= 16, >= 8 k-iterations each) that wastes the fewest SM-slots in the last wave
            double best = 0.0;
            for (int s = 1; s <= 16 && s <= k_iters_total / 8; ++s) {
                const long long units = out_tiles * s;
                const double eff = (double)units / (double)(((units + sms - 1) / sms) * sms);
                if (eff > best + 0.02) { best = eff; splits = s; }
            }
        }
        if (splits < 1) splits = 1;
    }
    PK_CHECK_ARG(splits == 1 || (gp.c_is_f32 && plain_epi && gp.zb0 == 1 && gp.zb1 == 1), "split-K needs a plain f32 2-D C");
    gp.iters_per_split = (k_iters_total + splits - 1) / splits;
    gp.k_splits = (k_iters_total + gp.iters_per_split - 1) / gp.iters_per_split;
    if (gp.k_splits > 1) {
        if (!gp.c_accumulate)
            PK_CHECK_CUDA(cudaMemset2DAsync(const_cast<void*>(d->c.ptr), (size_t)d->c.stride[0] * 4, 0, (size_t)N * 4, (size_t)M, stream));
        gp.c_accumulate = 1;
    }
    const long long num_tiles = out_tiles * gp.k_splits;
    PK_CHECK_ARG(num_tiles < (1ll << 31), "too many tiles");
    if (two_sm) {
        const long long tiles_m2 = (M + 255) / 256;
        const long long units = tiles_m2 * gp.tiles_n * gp.zb0 * gp.zb1 * gp.k_splits;
        int pairs = num_sms() / 2;
        if (units < pairs) pairs = (int)units;
        return dispatch_2sm(gp, d->a_mn_major, d->b_mn_major, pairs, stream);
    }
    int grid = num_sms();
    if (num_tiles < grid) grid = (int)num_tiles;
    if (bn == 64) return dispatch_major<64>(gp, d->a_mn_major, d->b_mn_major, grid, stream);
    if (bn == 128) return dispatch_major<128>(gp, d->a_mn_major, d->b_mn_major, grid, stream);
    return dispatch_major<256>(gp, d->a_mn_major, d->b_mn_major, grid, stream);
}
