// tcgen05 / TMEM / TMA GEMM for sm_100a: C = epilogue(alpha * sum_p A_p * B_p^T).
//
// One persistent CTA per SM (ONE mode) or one CTA PAIR per TPC (TWO mode, cta_group::2):
//   warp 0        TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx; in TWO mode both CTAs load
//                                 their own 128 rows of A and HALF of B, the bytes are credited to the leader's barrier)
//   warp 1        MMA issuer     (one thread: tcgen05.mma kind::f16, fp32 accumulators in TMEM; shared-memory descriptors are
//                                 built once and advanced by adding 16-byte units; tcgen05.commit frees smem stages / publishes
//                                 the accumulator; in TWO mode only the leader CTA issues, for both SMs)
//   warps 2..5    epilogue group 0   (tcgen05.ld TMEM -> registers -> fused epilogue -> swizzled smem -> TMA store;
//   warps 6..9    epilogue group 1    TWO mode only: the second group takes the right half of the tile's columns, which hides
//                                     the latency-bound row log-sum-exp epilogue of the joint projection)
// Tile 128 x BN x 64 per CTA (BN = 64 | 128 | 256; TWO: 256 x 256 per pair).  Operands may be K-major or MN-major (wgrad /
// dgrad / P.V use the MN-major form so no transposes are ever materialised).  Up to 9 (A,B) pairs accumulate into one tile
// (TDNN taps, split-bf16 fp32-class mode), plus a batched reduction loop (kz) for per-utterance wgrad, plus split-K.
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

struct GemmParams {
    CUtensorMap a[PK_GEMM_MAX_PAIRS];
    CUtensorMap b[PK_GEMM_MAX_PAIRS];
    CUtensorMap c;
    int a_off[PK_GEMM_MAX_PAIRS];
    int b_off[PK_GEMM_MAX_PAIRS];
    int n_pairs, kz_count, num_k_blocks;
    int k_splits, iters_per_split;      // split-K over the flattened (pair, kz, k-block) iteration space
    int split_major;                    // unit order: 1 = all tiles of split 0, then split 1, ... (CTAs that run together share a k-window)
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
    float* row_lse;                     // optional [tiles_n * EG][M][2] per-row (max*log2e, sum 2^(x*log2e-max)) partials of the bf16 output
    uint64_t pol_a, pol_b, pol_c;       // L2 eviction priorities of the three streams
};

template <int BN, bool TWO> struct GemmCfg {
    static constexpr int EG = TWO ? 2 : 1;                              // epilogue groups of 4 warps
    static constexpr int B_ROWS = TWO ? BN / 2 : BN;                    // rows of B this CTA stages
    static constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = TWO ? 5 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
    static constexpr int C_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = C_OFF + EG * 2 * C_STAGE_BYTES;
    static constexpr int BAR_OFF = BIAS_OFF + BN * 4;
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;             // + barriers + alignment slack
    static constexpr int TMEM_COLS = 2 * BN;                            // double-buffered accumulator (>= 32, pow2)
    static constexpr int THREADS = 64 + EG * 128;
    static_assert(!TWO || BN == 256, "the CTA-pair kernel works on 256 x 256 tiles");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

PK_DEVICE int pick_sel(int sel, int zb0, int zb1, int kz) {
    return sel == PK_SEL_ZB0 ? zb0 : (sel == PK_SEL_ZB1 ? zb1 : (sel == PK_SEL_KZ ? kz : 0));
}

struct UnitCoord { int mb, nb, zb0, zb1, split; };
// work unit -> (output tile, K split).  tiles_mu = M tiles of this kernel flavour (256-row tiles in TWO mode).
PK_DEVICE UnitCoord decode_unit(const GemmParams& p, int unit, int tiles_mu, int out_tiles) {
    UnitCoord u;
    int tile;
    if (p.split_major) { u.split = unit / out_tiles; tile = unit - u.split * out_tiles; }
    else { tile = unit / p.k_splits; u.split = unit - tile * p.k_splits; }
    const int tiles_per_z = tiles_mu * p.tiles_n;
    const int z = tile / tiles_per_z;
    const int r = tile - z * tiles_per_z;
    u.mb = r / p.tiles_n; u.nb = r - u.mb * p.tiles_n;
    u.zb1 = z / p.zb0; u.zb0 = z - u.zb1 * p.zb0;
    return u;
}

// EPI selects what the epilogue compiles in, so that the common case is straight-line code (the run-time-uniform
// branches of the full epilogue cost an instruction-fetch bubble per 16-byte group, profiles/r01_notes.md):
//   EPI_PLAIN  alpha, bias, ReLU only      EPI_LSE  + per-row log-sum-exp partials (bf16 C)      EPI_FULL  + dropout / aux add / aux mask
// The epilogue pulls 32 accumulator columns per tcgen05.ld (one wait per 32 columns).
enum { EPI_PLAIN = 0, EPI_LSE = 1, EPI_FULL = 2 };
template <bool A_MN, bool B_MN, int BN, bool CF32, int EPI, bool TWO>
__global__ void __launch_bounds__(GemmCfg<BN, TWO>::THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BN, TWO>;
    constexpr int EG = Cfg::EG;
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
    const uint32_t rank = TWO ? cluster_ctarank() : 0u;                 // position in the CTA pair
    const int worker = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;  // persistent worker (CTA or CTA pair)
    const int n_workers = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.n_pairs; ++i) {
            tma_prefetch_desc(&p.a[i]);
            tma_prefetch_desc(&p.b[i]);
        }
        tma_prefetch_desc(&p.c);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);      // one expect_tx arrive (the leader's in TWO mode; the peer's loads count by bytes only)
            mbar_init(&empty_bar[s], 1);     // tcgen05.commit (multicast to both CTAs in TWO mode)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4 * EG * (TWO ? 2 : 1));          // every epilogue warp (of both CTAs) releases the accumulator
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        if (TWO) { tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if (TWO) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_mu = TWO ? (p.M + 2 * BM - 1) / (2 * BM) : p.tiles_m;
    const int out_tiles = tiles_mu * p.tiles_n * p.zb0 * p.zb1;
    const int num_units = out_tiles * p.k_splits;                       // work units = tiles x K-splits
    const int k_iters_total = p.n_pairs * p.kz_count * p.num_k_blocks;

    if (warp == 0) {
        // ===================================================== TMA producer (converged warp; one elected lane issues, common.cuh)
        {
            const uint32_t lead = elect_one_u32();
            int stage = 0;
            uint32_t phase = 0;
            const int kzb = p.kz_count * p.num_k_blocks;
            for (int unit = worker; unit < num_units; unit += n_workers) {
                const UnitCoord u = decode_unit(p, unit, tiles_mu, out_tiles);
                const int m0 = TWO ? u.mb * (2 * BM) + (int)rank * BM : u.mb * BM;
                const int n0 = TWO ? u.nb * BN + (int)rank * (BN / 2) : u.nb * BN;   // this CTA's rows of B
                const int i0 = u.split * p.iters_per_split, i1 = min(k_iters_total, i0 + p.iters_per_split);
                // position in the flattened (pair, kz, k-block) space: decoded once per unit, then advanced by counters (a division per
                // k-block in this single thread costs more than the four MMAs of a stage take, profiles/r01_notes.md)
                int pr = i0 / kzb;
                int rem = i0 - pr * kzb;
                int kz = rem / p.num_k_blocks, kb = rem - kz * p.num_k_blocks;
                int a2 = pick_sel(p.a_sel2, u.zb0, u.zb1, kz), a3 = pick_sel(p.a_sel3, u.zb0, u.zb1, kz);
                int b2 = pick_sel(p.b_sel2, u.zb0, u.zb1, kz), b3 = pick_sel(p.b_sel3, u.zb0, u.zb1, kz);
                const CUtensorMap* ma = &p.a[pr];
                const CUtensorMap* mbp = &p.b[pr];
                int a_off = p.a_off[pr], b_off = p.b_off[pr];
                for (int i = i0; i < i1; ++i) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    if (!TWO || rank == 0) mbar_arrive_expect_tx_p(&full_bar[stage], (TWO ? 2 : 1) * Cfg::STAGE_BYTES, lead);
                    const int k0 = kb * BK;
                    if (A_MN) {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c) {
                            if (TWO) tma_load_4d_2sm_hint_p(sa + c * (64 * BK * 2), ma, &full_bar[stage], m0 + c * 64, k0 + a_off, a2, a3, p.pol_a, lead);
                            else tma_load_4d_hint_p(sa + c * (64 * BK * 2), ma, &full_bar[stage], m0 + c * 64, k0 + a_off, a2, a3, p.pol_a, lead);
                        }
                    } else {
                        if (TWO) tma_load_4d_2sm_hint_p(sa, ma, &full_bar[stage], k0, m0 + a_off, a2, a3, p.pol_a, lead);
                        else tma_load_4d_hint_p(sa, ma, &full_bar[stage], k0, m0 + a_off, a2, a3, p.pol_a, lead);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int c = 0; c < Cfg::B_ROWS / 64; ++c) {
                            if (TWO) tma_load_4d_2sm_hint_p(sb + c * (64 * BK * 2), mbp, &full_bar[stage], n0 + c * 64, k0 + b_off, b2, b3, p.pol_b, lead);
                            else tma_load_4d_hint_p(sb + c * (64 * BK * 2), mbp, &full_bar[stage], n0 + c * 64, k0 + b_off, b2, b3, p.pol_b, lead);
                        }
                    } else {
                        if (TWO) tma_load_4d_2sm_hint_p(sb, mbp, &full_bar[stage], k0, n0 + b_off, b2, b3, p.pol_b, lead);
                        else tma_load_4d_hint_p(sb, mbp, &full_bar[stage], k0, n0 + b_off, b2, b3, p.pol_b, lead);
                    }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    if (++kb == p.num_k_blocks) {
                        kb = 0;
                        if (++kz == p.kz_count) {
                            kz = 0;
                            if (++pr < p.n_pairs) { ma = &p.a[pr]; mbp = &p.b[pr]; a_off = p.a_off[pr]; b_off = p.b_off[pr]; }
                        }
                        a2 = pick_sel(p.a_sel2, u.zb0, u.zb1, kz); a3 = pick_sel(p.a_sel3, u.zb0, u.zb1, kz);
                        b2 = pick_sel(p.b_sel2, u.zb0, u.zb1, kz); b3 = pick_sel(p.b_sel3, u.zb0, u.zb1, kz);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (the leader CTA's in TWO mode; converged warp, one lane issues)
        if (rank == 0) {
            const uint32_t lead = elect_one_u32();
            constexpr uint32_t idesc = make_idesc_bf16(TWO ? 2 * BM : BM, BN, A_MN, B_MN);
            // K-major: 16 elements = 32 B inside the 128 B swizzle row; atoms of 8 rows (1024 B).
            // MN-major: 16 k-rows = two 8-row atoms (2048 B); 64-wide MN chunks 8192 B apart.
            // The descriptors are built once; a stage / k16 step only adds to the 14-bit start-address field (16-byte units,
            // never carries out of it: the whole ring lies below 256 KB).
            const uint32_t base16 = (smem_u32(smem) >> 4) & 0x3FFF;
            const uint64_t a_desc0 = (A_MN ? make_smem_desc_sw128(0, 64 * BK * 2, 1024) : make_smem_desc_sw128(0, 16, 1024)) + base16;
            const uint64_t b_desc0 = (B_MN ? make_smem_desc_sw128(0, 64 * BK * 2, 1024) : make_smem_desc_sw128(0, 16, 1024)) + base16 +
                                     (A_STAGE_BYTES >> 4);
            constexpr uint32_t A_K16 = (A_MN ? 2048 : 32) >> 4, B_K16 = (B_MN ? 2048 : 32) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t stage16 = 0;                    // stage * STAGE_BYTES / 16
            int it = 0;
            for (int unit = worker; unit < num_units; unit += n_workers, ++it) {
                const int split = p.split_major ? unit / out_tiles : unit % p.k_splits;
                const int k_iters = min(k_iters_total, (split + 1) * p.iters_per_split) - split * p.iters_per_split;
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t ad = a_desc0 + stage16, bd = b_desc0 + stage16;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 16; ++k4) {
                        if (TWO) umma_bf16_2sm_p(d_tmem, ad + k4 * A_K16, bd + k4 * B_K16, idesc, (k > 0 || k4 > 0) ? 1u : 0u, lead);
                        else umma_bf16_p(d_tmem, ad + k4 * A_K16, bd + k4 * B_K16, idesc, (k > 0 || k4 > 0) ? 1u : 0u, lead);
                    }
                    if (TWO) umma_commit_2sm_p(&empty_bar[stage], 3, lead); else umma_commit_p(&empty_bar[stage], lead);
                    stage16 += Cfg::STAGE_BYTES >> 4;
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; stage16 = 0; }
                }
                if (TWO) umma_commit_2sm_p(&tmem_full[acc], 3, lead); else umma_commit_p(&tmem_full[acc], lead);
            }
        }
    } else {
        // ===================================================== epilogue (EG groups of 4 warps)
        // Compact loop over 16-byte output groups (8 bf16 / 4 f32 columns) keeps the body resident in the instruction cache.
        const int g = (warp - 2) >> 2;          // epilogue group: columns [g * BN/EG, (g+1) * BN/EG) of the tile
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // tile row owned by this thread
        const int et = threadIdx.x - 64 - g * 128;   // 0..127 inside the group
        const uint32_t store_pred = (et == 0) ? 1u : 0u;    // the group's first thread owns the TMA stores (predicated, no divergent region)
        const int bar_stage = 1 + 2 * g, bar_bias = 2 + 2 * g;
        constexpr int GW = CF32 ? 4 : 8;        // columns per 16-byte group
        constexpr int CH = 8 * GW;              // columns per 128-byte staging row
        constexpr int GCOLS = BN / EG;          // columns of the tile this group handles
        uint8_t* cst = smem + Cfg::C_OFF + g * 2 * C_STAGE_BYTES;
        float* bias_g = bias_smem + g * GCOLS;
        const float relu_floor = (p.act == PK_ACT_RELU) ? 0.f : -INFINITY;
        if (p.bias == nullptr) {
            for (int j = et; j < GCOLS; j += 128) bias_g[j] = 0.f;
            named_bar_sync(bar_bias, 128);
        }
        int it = 0;
        uint32_t chunk_ctr = 0;
        for (int unit = worker; unit < num_units; unit += n_workers, ++it) {
            const UnitCoord u = decode_unit(p, unit, tiles_mu, out_tiles);
            const int zb0 = u.zb0, zb1 = u.zb1;
            const int m0 = TWO ? u.mb * (2 * BM) + (int)rank * BM : u.mb * BM;
            const int n0 = u.nb * BN + g * GCOLS;
            const int acc = it & 1;
            const int m = m0 + row;
            if (p.bias != nullptr) {
                named_bar_sync(bar_bias, 128);  // previous tile's readers are done with bias_smem
                for (int j = et; j < GCOLS; j += 128) bias_g[j] = (n0 + j < p.N) ? p.bias[n0 + j] : 0.f;
                named_bar_sync(bar_bias, 128);
            }
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + g * GCOLS;
            const bool row_ok = m < p.M;
            const unsigned char* aux_row = nullptr;
            if (EPI == EPI_FULL && p.aux_mode != PK_AUX_NONE && row_ok) {
                const long long off = (long long)m * p.aux_sm + (long long)zb0 * p.aux_s0 + (long long)zb1 * p.aux_s1;
                aux_row = reinterpret_cast<const unsigned char*>(p.aux) + off * (p.aux_is_f32 ? 4 : 2);
            }
            const uint64_t lin_row = ((uint64_t)(zb1 * p.zb0 + zb0) * (uint64_t)p.M + (uint64_t)m) * (uint64_t)p.N;
            float lse_m = -INFINITY, lse_s = 0.f;   // running row max (log2 units) and sum over this group's columns
            constexpr int n_chunks = GCOLS / CH;
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int nc0 = n0 + ch * CH;
                if (nc0 >= p.N || m0 >= p.M) break;   // uniform across the 4 warps of the group
                uint8_t* sbuf = cst + (chunk_ctr & 1) * C_STAGE_BYTES;
                tma_store_wait_read_p<1>(store_pred);           // the buffer used two chunks ago is free
                named_bar_sync(bar_stage, 128);
                uint8_t* srow = sbuf + row * 128;
                auto do_group = [&](const uint32_t (&rr)[GW], const int gq) {
                    const int ncol = nc0 + gq * GW;
                    float x[GW];
                    const float* bsm = bias_g + ch * CH + gq * GW;
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
                            const float m_new = fmaxf(lse_m, gm * 1.4426950408889634f);   // finite: the first group of a chunk set is valid
                            // raw MUFU.EX2 (no denormal fix-up sequence): eight independent exponentials issue back to back
                            float ex[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) ex[e] = ex2_approx(fmaf(r[e], 1.4426950408889634f, -m_new));
                            const float acc8 = ((ex[0] + ex[1]) + (ex[2] + ex[3])) + ((ex[4] + ex[5]) + (ex[6] + ex[7]));
                            lse_s = fmaf(lse_s, ex2_approx(lse_m - m_new), acc8);
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
                named_bar_sync(bar_stage, 128);
                if (p.c_accumulate) tma_reduce_add_4d_p(&p.c, sbuf, nc0, m0, zb0, zb1, store_pred);
                else tma_store_4d_hint_p(&p.c, sbuf, nc0, m0, zb0, zb1, p.pol_c, store_pred);
                tma_store_commit_p(store_pred);
                ++chunk_ctr;
            }
            // a group whose columns lie entirely beyond N contributes an empty partial (max = -inf, sum = 0), which the merge ignores
            if (!CF32 && EPI == EPI_LSE && row_ok)
                *reinterpret_cast<float2*>(p.row_lse + ((size_t)(u.nb * EG + g) * (size_t)p.M + (size_t)m) * 2) = make_float2(lse_m, lse_s);
            // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_remote(&tmem_empty[acc], 0); }
        }
        if (store_pred) tma_store_wait<0>();
    }

    tc_fence_before();
    if (TWO) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (TWO) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

// rank-3 bf16 map with 128B swizzle for the attention kernels: dims / strides as cuTensorMapEncodeTiled takes them
int encode_tiled_bf16_3d(CUtensorMap* out, const void* ptr, const unsigned long long (&dims)[3], const unsigned long long (&strides_bytes)[2],
                         const unsigned (&box)[3], const char* what) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point not available"); return -3; }
    {
        static thread_local bool ctx_ready = false;
        if (!ctx_ready) { PK_CHECK_CUDA(cudaFree(nullptr)); ctx_ready = true; }
    }
    cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
    cuuint64_t st[2] = {strides_bytes[0], strides_bytes[1]};
    cuuint32_t bx[3] = {box[0], box[1], box[2]};
    cuuint32_t es[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (st[0] % 16) != 0 || (st[1] % 16) != 0) {
        set_last_error("%s: base pointer / strides must be 16-byte aligned", what);
        return -1;
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r); return -3; }
    return 0;
}

template <bool A_MN, bool B_MN, int BN, bool CF32, int EPI, bool TWO>
static int launch_gemm_e(const GemmParams& gp, int workers, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, TWO>;
    auto kern = gemm_tcgen05_kernel<A_MN, B_MN, BN, CF32, EPI, TWO>;
    static bool configured = false;
    if (!configured) {
        PK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    if (TWO) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * workers);
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, gp));
    } else {
        kern<<<workers, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(gp);
        PK_CHECK_LAUNCH();
    }
    count_launch();
    return 0;
}

template <bool A_MN, bool B_MN, int BN, bool CF32, bool TWO>
static int launch_gemm(const GemmParams& gp, int workers, cudaStream_t stream) {
    if (gp.row_lse != nullptr) {
        // the row log-sum-exp epilogue exists for the joint projection's layout only: K-major operands, bf16 C, 256-wide tiles
        if constexpr (!A_MN && !B_MN && BN == 256 && !CF32) return launch_gemm_e<false, false, 256, false, EPI_LSE, TWO>(gp, workers, stream);
        set_last_error("gemm: row_lse needs K-major operands, a bf16 C and block_n = 256");
        return -1;
    }
    if (gp.drop_thresh != 0u || gp.aux_mode != PK_AUX_NONE) return launch_gemm_e<A_MN, B_MN, BN, CF32, EPI_FULL, TWO>(gp, workers, stream);
    return launch_gemm_e<A_MN, B_MN, BN, CF32, EPI_PLAIN, TWO>(gp, workers, stream);
}

template <int BN, bool TWO>
static int dispatch_major(const GemmParams& gp, int a_mn, int b_mn, int workers, cudaStream_t stream) {
    if (gp.c_is_f32) {
        if (!a_mn && !b_mn) return launch_gemm<false, false, BN, true, TWO>(gp, workers, stream);
        if (!a_mn && b_mn) return launch_gemm<false, true, BN, true, TWO>(gp, workers, stream);
        if (a_mn && !b_mn) return launch_gemm<true, false, BN, true, TWO>(gp, workers, stream);
        return launch_gemm<true, true, BN, true, TWO>(gp, workers, stream);
    }
    if (!a_mn && !b_mn) return launch_gemm<false, false, BN, false, TWO>(gp, workers, stream);
    if (!a_mn && b_mn) return launch_gemm<false, true, BN, false, TWO>(gp, workers, stream);
    if (a_mn && !b_mn) return launch_gemm<true, false, BN, false, TWO>(gp, workers, stream);
    return launch_gemm<true, true, BN, false, TWO>(gp, workers, stream);
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static uint64_t policy_of(int code) { return code == 1 ? kL2EvictFirst : (code == 2 ? kL2EvictLast : kL2EvictNormal); }

// Kernel flavour for a problem: tile width and whether the CTA-pair (cta_group::2) kernel runs it.  One place, shared by the
// launch and by pk_gemm_row_lse_parts (the caller sizes the partials buffer from it).
struct GemmPlan { int bn; bool two; };
static GemmPlan plan_gemm(long long M, long long N, int block_n, int two_sm_req) {
    GemmPlan pl;
    pl.bn = block_n ? block_n : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
    static int use_2sm = -1;                 // PK_GEMM_2SM=0 restores the single-CTA kernel everywhere
    if (use_2sm < 0) use_2sm = env_int("PK_GEMM_2SM", 1);
    static int min_tiles = -1;               // tuning hook: smallest number of 256 x 256 tiles handed to the pair kernel
    if (min_tiles < 0) min_tiles = env_int("PK_GEMM_2SM_MIN_TILES", 1);
    const long long tiles2 = ((M + 255) / 256) * ((N + 255) / 256);
    const int want = two_sm_req < 0 ? 0 : (two_sm_req > 0 ? 1 : (use_2sm && tiles2 >= min_tiles));
    pl.two = want && pl.bn == 256 && M > 128;
    return pl;
}

}  // namespace pk

extern "C" int pk_gemm_row_lse_parts(long long M, long long N, int block_n, int two_sm) {
    const pk::GemmPlan pl = pk::plan_gemm(M, N, block_n ? block_n : 256, two_sm);
    return (int)((N + pl.bn - 1) / pl.bn) * (pl.two ? 2 : 1);
}

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
    PK_CHECK_ARG(d->block_n == 0 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256, "block_n must be 64, 128 or 256");
    const GemmPlan plan = plan_gemm(M, N, d->row_lse ? (d->block_n ? d->block_n : 256) : d->block_n, d->two_sm);
    const int bn = plan.bn;
    const bool two_sm = plan.two;       // CTA-pair kernel: 256 x 256 per pair (needs M > 128 so that the second CTA has rows)

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
    PK_CHECK_ARG(d->row_lse == nullptr || (d->c_dtype == PK_BF16 && N % 8 == 0 && gp.zb0 == 1 && gp.zb1 == 1 && bn == 256 &&
                                           gp.drop_thresh == 0u && gp.aux_mode == PK_AUX_NONE),
                 "row_lse needs a 2-D bf16 C with N % 8 == 0, block_n = 256, no dropout / aux");
    {   // L2 eviction priorities by stream size (measured, profiles/r02_gemm_lab.txt): an operand that fits in L2 many times over
        // (a weight matrix) is kept with evict_last -- +14 % on the joint dgrad; the same hint on a multi-GB streamed operand (the
        // wgrad's activations) costs 15 %; a C stream much larger than L2 leaves first.  PK_GEMM_L2_HINTS=0 -> all normal.
        static int hints = -1;
        if (hints < 0) hints = env_int("PK_GEMM_L2_HINTS", 1);
        const long long K_ = (long long)gp.num_k_blocks * BK * gp.kz_count;
        const long long a_bytes = M * K_ * 2 * gp.n_pairs, b_bytes = N * K_ * 2 * gp.n_pairs;
        const long long c_bytes = M * N * (gp.c_is_f32 ? 4 : 2) * gp.zb0 * gp.zb1;
        const long long small = 32ll << 20, big = 256ll << 20;
        gp.pol_a = policy_of(hints && a_bytes <= small && gp.zb0 * gp.zb1 == 1 ? 2 : 0);
        gp.pol_b = policy_of(hints && b_bytes <= small && gp.zb0 * gp.zb1 == 1 ? 2 : 0);
        gp.pol_c = policy_of(hints && c_bytes >= big ? 1 : 0);
    }

    // work in units of this flavour's tiles and workers (256-row tiles on CTA pairs, or 128-row tiles on single CTAs)
    const long long out_tiles = (two_sm ? (M + 255) / 256 : (long long)gp.tiles_m) * gp.tiles_n * gp.zb0 * gp.zb1;
    const int workers_max = two_sm ? num_sms() / 2 : num_sms();
    // split-K: under-filled grids with a long reduction (wgrad, the LSTM's recurrent dgrad) are cut along the
    // flattened (pair, kz, k-block) axis; partial tiles are combined with TMA reduce-add into a zeroed f32 C.
    const int k_iters_total = gp.n_pairs * gp.kz_count * gp.num_k_blocks;
    int splits = 1;
    const bool plain_epi = d->bias == nullptr && d->act == PK_ACT_NONE && d->drop_p == 0.f && gp.aux_mode == PK_AUX_NONE;
    if (d->k_splits > 0) splits = d->k_splits;
    else if (gp.c_is_f32 && plain_epi && gp.zb0 == 1 && gp.zb1 == 1 && k_iters_total >= 16) {
        // PK_GEMM_SPLIT_MODE: 1 (default) = the split count (>= 8 k-blocks each, <= PK_GEMM_SPLIT_MAX = 16) that wastes the least of the last
        // wave, preferring fewer splits on ties: fc2 wgrad 1138 -> 1269 TFLOP/s stand-alone, 71.0 -> 69.5 ms per step same box
        // (profiles/r02_gemm_lab.txt) | 0 = fill about two waves (the round-1 rule)
        static int mode = -1, smax = -1;
        if (mode < 0) { mode = env_int("PK_GEMM_SPLIT_MODE", 1); smax = env_int("PK_GEMM_SPLIT_MAX", 16); }
        const int w = workers_max;
        if (mode == 0) {
            if (out_tiles < 2 * w) {
                splits = (int)((2 * w + out_tiles / 2) / out_tiles);
                if (splits > k_iters_total / 8) splits = k_iters_total / 8;
                if (splits > 64) splits = 64;
            }
        } else if (out_tiles < 6 * w) {
            double best = 0.0;
            for (int sp = 1; sp <= smax && sp <= k_iters_total / 8; ++sp) {
                const long long units = out_tiles * sp;
                const double eff = (double)units / (double)(((units + w - 1) / w) * w);
                if (eff > best + 0.02) { best = eff; splits = sp; }
            }
        }
        if (splits < 1) splits = 1;
    }
    PK_CHECK_ARG(splits == 1 || (gp.c_is_f32 && plain_epi && gp.zb0 == 1 && gp.zb1 == 1), "split-K needs a plain f32 2-D C");
    {
        static int sm = -1;
        if (sm < 0) sm = env_int("PK_GEMM_SPLIT_MAJOR", 1);
        gp.split_major = sm;
    }
    gp.iters_per_split = (k_iters_total + splits - 1) / splits;
    gp.k_splits = (k_iters_total + gp.iters_per_split - 1) / gp.iters_per_split;
    if (gp.k_splits > 1) {
        if (!gp.c_accumulate)
            PK_CHECK_CUDA(cudaMemset2DAsync(const_cast<void*>(d->c.ptr), (size_t)d->c.stride[0] * 4, 0, (size_t)N * 4, (size_t)M, stream));
        gp.c_accumulate = 1;
    }
    const long long num_tiles = out_tiles * gp.k_splits;
    PK_CHECK_ARG(num_tiles < (1ll << 31), "too many tiles");
    int grid = workers_max;
    if (num_tiles < grid) grid = (int)num_tiles;
    if (two_sm) return dispatch_major<256, true>(gp, d->a_mn_major, d->b_mn_major, grid, stream);
    if (bn == 64) return dispatch_major<64, false>(gp, d->a_mn_major, d->b_mn_major, grid, stream);
    if (bn == 128) return dispatch_major<128, false>(gp, d->a_mn_major, d->b_mn_major, grid, stream);
    return dispatch_major<256, false>(gp, d->a_mn_major, d->b_mn_major, grid, stream);
}
