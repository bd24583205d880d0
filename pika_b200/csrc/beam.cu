// Device-resident batched beam search for the transducer (decode + MBR N-best).
//
// Replaces the per-utterance Python bookkeeping of decoder/beam_transducer.py:82-187
// (BeamMergeTransducer.advance: score add, EOS / duplicate-hypothesis kill, top-k over beam*V,
// finish rule, partial-hypothesis update) and the gather / state-reorder steps of
// decoder/transducer_decoder.py:127-150,188-202, for ALL utterances of the batch in one launch
// per step, with no host round trip except a single "all done" flag.
//
// Row layout: row = b * K + k (utterance-major).  blank = blk, EOS = -1 as in the reference.
#include "../../include/pika_b200.h"
#include "common.cuh"

namespace pk {
void count_launch();

constexpr float kKill = -1e20f;

// The beam loop is replayed from a CUDA graph, so nothing that changes from step to step may be a kernel argument: the step index
// lives in device memory (advanced by beam_step_inc_kernel at the end of every step) and every kernel turns into a no-op once all
// utterances are done (``while not all(b.done() for b in beam)``, decoder/transducer_decoder.py:123) or the history buffers are full.
struct StepCtx {
    const int* step;          // [2]: step[0] = current step, step[1] = 1 while the loop is live (latched between steps, so that every
};                            //      kernel and every CTA of one step sees the same value even while beam_advance counts utterances down)
PK_DEVICE bool step_active(const StepCtx& c) { return c.step[1] != 0; }
// end of a step: advance the counter and latch the loop condition for the next step
__global__ void beam_step_end_kernel(int* step, const int* not_done, int max_steps) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && step[1] != 0) {
        const int s = step[0] + 1;
        step[0] = s;
        step[1] = (*not_done > 0 && s < max_steps) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------ step prologue
// t_idx += (tok == blk); enc_hid[row] = enc[b, t_idx[row]]; x_emb[row] = embed[tok] (zeros unless tok > blk)
template <typename T>
__global__ void beam_prepare_kernel(const int* __restrict__ next_ys, StepCtx ctx, int* __restrict__ t_idx, const T* __restrict__ enc, int Tenc, int H,
                                    T* __restrict__ enc_hid, const float* __restrict__ embed, int E, T* __restrict__ x_emb, int ld_x,
                                    int K, int blk, int rows) {
    const int row = blockIdx.x;
    if (row >= rows || !step_active(ctx)) return;
    const int b = row / K;
    const int tk = next_ys[(long long)ctx.step[0] * rows + row];
    __shared__ int s_t;
    if (threadIdx.x == 0) {
        int t = t_idx[row] + (tk == blk ? 1 : 0);
        t_idx[row] = t;
        s_t = min(max(t, 0), Tenc - 1);
    }
    __syncthreads();
    const T* src = enc + ((long long)b * Tenc + s_t) * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) enc_hid[(long long)row * H + c] = src[c];
    for (int c = threadIdx.x; c < ld_x; c += blockDim.x)
        x_emb[(long long)row * ld_x + c] = from_f32<T>((tk > blk && c < E) ? embed[(long long)tk * E + c] : 0.f);
}

// LSTM cell on rows whose current token is a real label (tok > blk); other rows keep (h, c).
template <typename T>
__global__ void beam_lstm_cell_kernel(const float* __restrict__ gates, const int* __restrict__ next_ys, StepCtx ctx, int blk, T* __restrict__ h,
                                      float* __restrict__ c, int rows, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * H || !step_active(ctx)) return;
    const int r = i / H, j = i - r * H;
    if (!(next_ys[(long long)ctx.step[0] * rows + r] > blk)) return;
    const float* g = gates + (long long)r * 4 * H;
    const float gi = 1.f / (1.f + expf(-g[j])), gf = 1.f / (1.f + expf(-g[H + j]));
    const float gg = tanhf(g[2 * H + j]), go = 1.f / (1.f + expf(-g[3 * H + j]));
    const float cn = gf * c[i] + gi * gg;
    c[i] = cn;
    h[i] = from_f32<T>(go * tanhf(cn));
}

// h = tanh(a[:, :H]) * sigmoid(a[:, H:])   (gated joint on [rows, 2H] pre-activations)
template <typename T>
__global__ void beam_gate_kernel(const float* __restrict__ a, T* __restrict__ h, int rows, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * H) return;
    const int r = i / H, j = i - r * H;
    const float x = a[(long long)r * 2 * H + j], g = a[(long long)r * 2 * H + H + j];
    h[i] = from_f32<T>(tanhf(x) * (1.f / (1.f + expf(-g))));
}

// ------------------------------------------------------------------------------------ beam advance
struct BeamState {
    float* scores;      // [B,K]
    int* next_ys;       // [S+1, B, K]  (entry 0 = initial blk)
    int* prev_ks;       // [S, B, K]
    int* hyp_tok;       // [2, B, K, L] partial (non-blank) hypotheses, ping-pong by step parity
    int* hyp_len;       // [2, B, K]
    float* fin_score;   // [B, cap]
    int* fin_step;      // [B, cap]
    int* fin_k;         // [B, cap]
    int* fin_count;     // [B]
    int* eos_top;       // [B]
    int* done;          // [B]
    int* not_done_total;  // [1] number of utterances not yet done (written every step)
};

// ------------------------------------------------------------------------------------ on-the-fly FST shallow fusion
// decoder/sorted_matcher.py:24-111 over a flattened (CSR) arc table: arcs of a state are sorted by ilabel.  Costs are
// accumulated in double like the reference's Python floats; the per-beam state sets keep INSERTION ORDER, which the
// reference's strict-< update with the reward subtracted only from the stored value makes observable
// (decoder/beam_transducer.py:146-149).
constexpr int LM_MAX_DISAMBIG = 4;
struct LmArgs {
    const int* arc_off;        // [n_states + 1]; nullptr = no LM
    const int* arc_il;         // [n_arcs] input labels (token + 1)
    const double* arc_w;       // [n_arcs] costs
    const int* arc_ns;         // [n_arcs] next states
    const double* finals;      // [n_states] final cost, +inf = not final
    int backoff_id, n_disambig;
    int disambig[LM_MAX_DISAMBIG];
    float scale;               // lm_scorer_scale
    double scale_d, reward;    // lm_scorer_scale and args.nonblk_reward as Python floats
    int* set_state;            // [2][B][K][MS] state sets, ping-pong by step parity
    double* set_cost;          // [2][B][K][MS]
    int* set_n;                // [2][B][K]
    float* lm_scores;          // [B][K]
    int MS;
    int* err;                  // set to 1 when a state set overflows MS
};

PK_DEVICE int fst_search(const LmArgs& f, int state, int ilabel) {          // SortedMatcher.search (:24-50): first arc with this ilabel
    int lo = f.arc_off[state], hi = f.arc_off[state + 1];
    const int end = hi;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (f.arc_il[mid] >= ilabel) hi = mid; else lo = mid + 1;
    }
    return (lo < end && f.arc_il[lo] == ilabel) ? lo : -1;
}
// get_scores (:70-85) -> emit(cost, next_state) in the reference's order: the state itself, then its disambiguation arcs, each
// followed down its back-off chain (:52-68)
template <typename F> PK_DEVICE void fst_get_scores(const LmArgs& f, int state, int ilabel, F&& emit) {
    for (int i = -1; i < f.n_disambig; ++i) {
        double bf = 0.0;
        int cur = state;
        if (i >= 0) {
            const int a = fst_search(f, state, f.disambig[i]);
            if (a < 0) continue;
            bf = f.arc_w[a]; cur = f.arc_ns[a];
        }
        while (true) {
            const int a = fst_search(f, cur, ilabel);
            if (a >= 0) emit(bf + f.arc_w[a], f.arc_ns[a]);
            const int bo = fst_search(f, cur, f.backoff_id);
            if (bo < 0) break;
            bf += f.arc_w[bo]; cur = f.arc_ns[bo];
        }
    }
}
// min over final_score(state) (:87-111) of (base + cost); +inf when no final state is reachable
PK_DEVICE double fst_final_min(const LmArgs& f, int state, double base) {
    double best = INFINITY;
    for (int i = -1; i < f.n_disambig; ++i) {
        double sc = 0.0;
        int cur = state;
        if (i >= 0) {
            const int a = fst_search(f, state, f.disambig[i]);
            if (a < 0) continue;
            sc = f.arc_w[a]; cur = f.arc_ns[a];
        }
        while (true) {
            const double fc = f.finals[cur];
            if (isinf(fc)) {
                const int bo = fst_search(f, cur, f.backoff_id);
                if (bo < 0) { sc = INFINITY; break; }
                sc += f.arc_w[bo]; cur = f.arc_ns[bo];
            } else { sc += fc; break; }
        }
        best = fmin(best, base + sc);
    }
    return best;
}

template <int K>
PK_DEVICE void topk_insert(float (&v)[K], int (&ix)[K], float x, int id) {
    // keeps v descending; ties keep the earlier (smaller-index) element first
    if (!(x > v[K - 1])) return;
    v[K - 1] = x; ix[K - 1] = id;
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        if (v[j] > v[j - 1]) {
            const float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv;
            const int ti = ix[j]; ix[j] = ix[j - 1]; ix[j - 1] = ti;
        }
    }
}

constexpr int ADV_THREADS = 512;                    // 16 warps (128 registers each): the candidate scan is latency-bound, one CTA per utterance
constexpr int ADV_LOADS = 12;                       // independent score loads in flight per thread

template <int K>
__global__ void __launch_bounds__(ADV_THREADS) beam_advance_kernel(const float* __restrict__ logits, int ldv, const float* __restrict__ row_lse,
                                                                   float sm_scale, const int* __restrict__ t_idx,
                                                                   const int* __restrict__ num_frames, const int* __restrict__ max_len,
                                                                   BeamState st, int B, int V, int L, int cap, StepCtx ctx, int blk,
                                                                   int n_best, int beam_prune, LmArgs lm) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    if (!step_active(ctx)) return;                 // uniform across the grid (latched by beam_step_end_kernel)
    const int step = ctx.step[0];
    const bool use_lm = lm.arc_off != nullptr;
    __shared__ float s_rowscore[K];
    __shared__ float s_lmterm[K];            // lm_scorer_scale * lm_scores[k] (fp32 product, as the reference's tensor expression forms it)
    __shared__ int s_kill[K];
    __shared__ float s_wv[ADV_THREADS / 32][K];     // per-warp top-K (sorted: value desc, index asc)
    __shared__ int s_wi[ADV_THREADS / 32][K];
    __shared__ float s_best[K];
    __shared__ int s_besti[K];
    __shared__ float s_lse[K];
    __shared__ int s_oldlen[K];

    const int par_old = step & 1, par_new = par_old ^ 1;
    const int* cur_tok = st.next_ys + ((long long)step * B + b) * K;
    const int* old_hyp = st.hyp_tok + (((long long)par_old * B + b) * K) * L;
    const int* old_len = st.hyp_len + ((long long)par_old * B + b) * K;
    int* new_hyp = st.hyp_tok + (((long long)par_new * B + b) * K) * L;
    int* new_len = st.hyp_len + ((long long)par_new * B + b) * K;

    const int* hyp_src = old_hyp;
    // ---- 1. which beam rows may have children.  Warp i decides row i; the duplicate test (first row with the same partial hypothesis
    // wins, decoder/beam_transducer.py:105-114) compares the two token rows 32 elements at a time.
    {
        const int wi = tid >> 5, ln = tid & 31;
        if (wi < K) {
            int kill = 0;
            if (step > 0) {
                if (cur_tok[wi] == -1) kill = 1;                                // finished beams have no children
                else if (beam_prune && old_len[wi] > 0) {
                    const int n = old_len[wi];
                    for (int j = 0; j < wi && !kill; ++j) {
                        if (cur_tok[j] == -1 || old_len[j] != n) continue;
                        bool same = true;
                        for (int q0 = 0; q0 < n && same; q0 += 32) {
                            const int q = q0 + ln;
                            const bool eq = q >= n || hyp_src[(long long)j * L + q] == hyp_src[(long long)wi * L + q];
                            same = __all_sync(0xffffffffu, eq);
                        }
                        if (same) kill = 1;
                    }
                }
            } else if (wi > 0) {
                kill = 2;                                                       // first step: only row 0 is expanded
            }
            if (ln == 0) {
                s_kill[wi] = kill;
                s_lse[wi] = row_lse[b * K + wi];
                s_oldlen[wi] = old_len[wi];
                s_rowscore[wi] = st.scores[b * K + wi];
                s_lmterm[wi] = use_lm ? lm.scale * lm.lm_scores[b * K + wi] : 0.f;
            }
        }
    }
    __syncthreads();

    // ---- 2. per-thread top-K over the K*V candidates (flat index = k*V + v).  The scores sit in L2 (the joint GEMM just wrote them): a
    // thread's candidates of one row arrive as ADV_LOADS / 4 independent 16-byte loads, and the next batch is already in flight while this
    // one is visited (the first version, one scalar load per iteration, waited an L2 round trip 375 times: 150 us per step at beam 16 x
    // V = 6000).  word_probs[k, v] = log_softmax(sm_scale * logits)[k, v] = logits[k, v] * sm_scale - row_lse[k] is formed here with the
    // expression pk_log_softmax uses, so the candidate scores are the ones a separate log-softmax pass would have written (not run).
    // Tried on top of this and measured no better (profiles/r02_notes.md): a two-pass threshold filter (the list insertion is executed by
    // a warp whenever ANY lane inserts), two CTAs per utterance, hypotheses staged in shared memory for the duplicate test.
    float lv[K];
    int li[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { lv[j] = -INFINITY; li[j] = 0x7fffffff; }
    const float* wp = logits + (long long)b * K * ldv;
    // a batch = ADV_LOADS candidates per thread as ADV_LOADS / 4 16-byte loads (4 consecutive labels each): (row it / nb, batch it % nb)
    const int nb = (V + ADV_LOADS * ADV_THREADS - 1) / (ADV_LOADS * ADV_THREADS);
    const bool vec_ok = (ldv % 4) == 0 && (reinterpret_cast<uintptr_t>(wp) & 15) == 0;
    auto fetch = [&](int k, int bi, float (&xs)[ADV_LOADS]) {
        if (k >= K || s_kill[k] != 0) return;
        const float* wr = wp + (long long)k * ldv;
#pragma unroll
        for (int u4 = 0; u4 < ADV_LOADS / 4; ++u4) {
            const int v = (bi * (ADV_LOADS / 4) + u4) * 4 * ADV_THREADS + 4 * tid;
            if (vec_ok && v + 3 < V) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(wr + v));
                xs[4 * u4] = q.x; xs[4 * u4 + 1] = q.y; xs[4 * u4 + 2] = q.z; xs[4 * u4 + 3] = q.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) xs[4 * u4 + e] = (v + e < V) ? __ldg(wr + v + e) : -INFINITY;
            }
        }
    };
    auto scan = [&](auto&& visit) {                                            // visit(x, flat index) for every candidate of this thread,
        float cur[ADV_LOADS], n1[ADV_LOADS];                                    // in increasing flat index (ties keep their order)
        int k = 0, bi = 0, k1 = 0, b1 = 0;
        auto next = [&](int& kk, int& bb) { if (++bb == nb) { bb = 0; ++kk; } };
        fetch(k, bi, cur);
        k1 = k; b1 = bi; next(k1, b1); fetch(k1, b1, n1);
        while (k < K) {
            const int kill = s_kill[k];
            if (kill == 0) {
                const float add = (step > 0) ? s_rowscore[k] : 0.f;
                const float lmt = (step > 0 && use_lm) ? s_lmterm[k] : 0.f;
                const float lse_k = s_lse[k];
#pragma unroll
                for (int u = 0; u < ADV_LOADS; ++u) {
                    const int v = (bi * (ADV_LOADS / 4) + (u >> 2)) * 4 * ADV_THREADS + 4 * tid + (u & 3);
                    if (v >= V) continue;
                    float x = cur[u] * sm_scale - lse_k;
                    if (step > 0) {
                        x = x + add;                                         // word_probs + scores (+ lm term, in this order: beam_transducer.py:94-97)
                        if (use_lm) x += lmt;
                    }
                    visit(x, k * V + v);
                }
            } else if (kill == 1) {                                             // a finished / duplicate row: V candidates at -1e20
#pragma unroll
                for (int u = 0; u < ADV_LOADS; ++u) {
                    const int v = (bi * (ADV_LOADS / 4) + (u >> 2)) * 4 * ADV_THREADS + 4 * tid + (u & 3);
                    if (v < V) visit(kKill, k * V + v);
                }
            }                                                                   // kill == 2 (first step, rows > 0): no candidates
#pragma unroll
            for (int u = 0; u < ADV_LOADS; ++u) cur[u] = n1[u];
            k = k1; bi = b1;
            next(k1, b1);
            fetch(k1, b1, n1);
        }
    };
    scan([&](float x, int id) { topk_insert<K>(lv, li, x, id); });
    // ---- 3. merge: K rounds of warp arg-max over the heads of the 32 sorted lists of a warp (value desc, index asc; the winner shifts
    // its list), then the same over the 32 warp lists by warp 0 -- registers and shuffles only, two block barriers in all
    auto warp_merge = [&](float (&v)[K], int (&ix)[K], float* out_v, int* out_i) {
        const int ln = tid & 31;
        for (int round = 0; round < K; ++round) {
            float bv = v[0];
            int bi = ix[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
            }
            if (ln == 0) { out_v[round] = bv; out_i[round] = bi; }
            if (bi != 0x7fffffff && ix[0] == bi) {                             // candidate indices are unique: exactly one lane consumed
#pragma unroll
                for (int j = 0; j < K - 1; ++j) { v[j] = v[j + 1]; ix[j] = ix[j + 1]; }
                v[K - 1] = -INFINITY; ix[K - 1] = 0x7fffffff;
            }
        }
    };
    warp_merge(lv, li, s_wv[tid >> 5], s_wi[tid >> 5]);
    __syncthreads();
    if (tid < 32) {
        const bool has = tid < ADV_THREADS / 32;
#pragma unroll
        for (int j = 0; j < K; ++j) { lv[j] = has ? s_wv[has ? tid : 0][j] : -INFINITY; li[j] = has ? s_wi[has ? tid : 0][j] : 0x7fffffff; }
        warp_merge(lv, li, s_best, s_besti);
    }
    __syncthreads();

    // ---- 4. new beam: back-pointers, tokens, scores, finish rule, partial hypotheses
    int* out_tok = st.next_ys + ((long long)(step + 1) * B + b) * K;
    int* out_prev = st.prev_ks + ((long long)step * B + b) * K;
    const int nf = num_frames[b];
    const int len_after = step + 2;                                             // len(self.next_ys) after the append
    __shared__ int s_fin[K];
    __shared__ float s_final[K];             // score recorded for a finishing beam
    if (tid < K) {
        const int id = s_besti[tid];
        const int pk = id / V, y = id - pk * V;
        float sc = s_best[tid];
        if (use_lm) sc -= s_lmterm[pk];                                        // self.scores -= lm_scale * lm_scores[prev_k]  (:131-132)
        out_prev[tid] = pk;
        const bool fin = (y == blk && t_idx[b * K + pk] == nf - 1) || (len_after > max_len[b]);
        if (use_lm) {
            // ---- next FST state set of this beam (:135-159)
            const int MS = lm.MS;
            const long long o_old = (((long long)par_old * B + b) * K + pk) * MS, o_new = (((long long)par_new * B + b) * K + tid) * MS;
            const int n_old = (step > 0) ? lm.set_n[((long long)par_old * B + b) * K + pk] : 1;
            int n_new = 0;
            int* ns_state = lm.set_state + o_new;
            double* ns_cost = lm.set_cost + o_new;
            for (int s_i = 0; s_i < n_old; ++s_i) {
                const int state = (step > 0) ? lm.set_state[o_old + s_i] : 0;       // initial set {0: 0.0} (:64-66)
                const double c0 = (step > 0) ? lm.set_cost[o_old + s_i] : 0.0;
                if (y == blk) {
                    if (n_new < MS) { ns_state[n_new] = state; ns_cost[n_new] = c0; ++n_new; } else *lm.err = 1;
                    continue;
                }
                fst_get_scores(lm, state, y + 1, [&](double cost, int nxt) {
                    const double nc = c0 + cost;
                    int idx = -1;
                    for (int q = 0; q < n_new; ++q) if (ns_state[q] == nxt) { idx = q; break; }
                    if (idx < 0) {
                        if (n_new >= MS) { *lm.err = 1; return; }
                        idx = n_new++; ns_state[idx] = nxt; ns_cost[idx] = nc - lm.reward;      // first visit: inf > nc
                    } else if (nc < ns_cost[idx]) {
                        ns_cost[idx] = nc - lm.reward;                                          // (:147-149)
                    }
                });
            }
            lm.set_n[((long long)par_new * B + b) * K + tid] = n_new;
            double mn = INFINITY;
            for (int q = 0; q < n_new; ++q) mn = fmin(mn, ns_cost[q]);
            lm.lm_scores[b * K + tid] = n_new ? (float)(-mn) : -1e20f;                          // (:155-158)
            if (fin) {                                                                          // (:167-176)
                double fmn = INFINITY;
                for (int q = 0; q < n_new; ++q) fmn = fmin(fmn, fst_final_min(lm, ns_state[q], ns_cost[q]));
                sc += (float)(lm.scale_d * (-fmn));          // `s += lm_scale * final_lm_score` on the 0-d view of self.scores[i]
            }
        }
        st.scores[b * K + tid] = sc;
        s_final[tid] = sc;
        out_tok[tid] = fin ? -1 : y;
        s_fin[tid] = fin ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) {
        int cnt = st.fin_count[b];
        for (int i = 0; i < K; ++i) {
            if (s_fin[i] && cnt < cap) {
                st.fin_score[(long long)b * cap + cnt] = s_final[i];
                st.fin_step[(long long)b * cap + cnt] = step + 1;
                st.fin_k[(long long)b * cap + cnt] = i;
                ++cnt;
            }
        }
        st.fin_count[b] = cnt;
        int et = st.eos_top[b];
        if (s_fin[0]) et = 1;
        st.eos_top[b] = et;
        const int dn = (et && cnt >= n_best) ? 1 : 0;
        if (dn && !st.done[b]) { st.done[b] = 1; atomicSub(st.not_done_total, 1); }
    }
    // partial hypotheses: new[k] = finished ? old[k] : old[prev_k] (+ y if y != blk)
    for (int k = 0; k < K; ++k) {
        const int pk = s_besti[k] / V, y = s_besti[k] - (s_besti[k] / V) * V;
        const int src = s_fin[k] ? k : pk;
        const int n = s_oldlen[src];
        for (int q = tid; q < n; q += ADV_THREADS) new_hyp[(long long)k * L + q] = hyp_src[(long long)src * L + q];
        if (tid == 0) {
            int nn = n;
            if (!s_fin[k] && y != blk && nn < L) { new_hyp[(long long)k * L + nn] = y; ++nn; }
            new_len[k] = nn;
        }
    }
}

// dec_states / t_idx reordering by the new back-pointers (TransducerDecoder._beam_update)
template <typename T>
__global__ void beam_reorder_kernel(const int* __restrict__ prev_ks, StepCtx ctx, const T* __restrict__ h_in, const float* __restrict__ c_in,
                                    const int* __restrict__ t_in, T* __restrict__ h_out, float* __restrict__ c_out, int* __restrict__ t_out,
                                    int K, int H, int layers, int rows) {
    const int row = blockIdx.x;
    const int b = row / K;
    if (!step_active(ctx)) return;
    const int src = b * K + prev_ks[(long long)ctx.step[0] * rows + row];
    for (int l = 0; l < layers; ++l) {
        const long long o = ((long long)l * rows + row) * H, s = ((long long)l * rows + src) * H;
        for (int c = threadIdx.x; c < H; c += blockDim.x) { h_out[o + c] = h_in[s + c]; c_out[o + c] = c_in[s + c]; }
    }
    if (threadIdx.x == 0) t_out[row] = t_in[src];
}
}  // namespace pk

using namespace pk;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int pk_beam_prepare(const int* next_ys, const int* step_ctx, int* t_idx, const void* enc, int dtype, int Tenc, int H, void* enc_hid,
                               const float* embed, int E, void* x_emb, int ld_x, int K, int blk, int rows, void* stream) {
    const StepCtx ctx{step_ctx};
    if (dtype == PK_BF16)
        beam_prepare_kernel<__nv_bfloat16><<<rows, 128, 0, ST(stream)>>>(next_ys, ctx, t_idx, (const __nv_bfloat16*)enc, Tenc, H, (__nv_bfloat16*)enc_hid,
                                                                        embed, E, (__nv_bfloat16*)x_emb, ld_x, K, blk, rows);
    else
        beam_prepare_kernel<float><<<rows, 128, 0, ST(stream)>>>(next_ys, ctx, t_idx, (const float*)enc, Tenc, H, (float*)enc_hid, embed, E,
                                                                (float*)x_emb, ld_x, K, blk, rows);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
extern "C" int pk_beam_lstm_cell(const float* gates, const int* next_ys, const int* step_ctx, int blk, void* h, int dtype, float* c, int rows, int H,
                                 void* stream) {
    const int n = rows * H;
    const StepCtx ctx{step_ctx};
    if (dtype == PK_BF16) beam_lstm_cell_kernel<__nv_bfloat16><<<(n + 255) / 256, 256, 0, ST(stream)>>>(gates, next_ys, ctx, blk, (__nv_bfloat16*)h, c, rows, H);
    else beam_lstm_cell_kernel<float><<<(n + 255) / 256, 256, 0, ST(stream)>>>(gates, next_ys, ctx, blk, (float*)h, c, rows, H);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
extern "C" int pk_beam_gate(const float* a, void* h, int dtype, int rows, int H, void* stream) {
    const int n = rows * H;
    if (dtype == PK_BF16) beam_gate_kernel<__nv_bfloat16><<<(n + 255) / 256, 256, 0, ST(stream)>>>(a, (__nv_bfloat16*)h, rows, H);
    else beam_gate_kernel<float><<<(n + 255) / 256, 256, 0, ST(stream)>>>(a, (float*)h, rows, H);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
static int beam_advance_impl(const float* logits, int ldv, const float* row_lse, float sm_scale, const int* t_idx, const int* num_frames,
                             const int* max_len, float* scores,
                             int* next_ys, int* prev_ks, int* hyp_tok, int* hyp_len, float* fin_score, int* fin_step, int* fin_k,
                             int* fin_count, int* eos_top, int* done, int* not_done_total, int B, int K, int V, int L, int cap,
                             const int* step_ctx, int blk, int n_best, int beam_prune, const LmArgs& lm, void* stream) {
    const StepCtx step{step_ctx};
    BeamState st{scores, next_ys, prev_ks, hyp_tok, hyp_len, fin_score, fin_step, fin_k, fin_count, eos_top, done, not_done_total};
    PK_CHECK_ARG(V >= K, "vocabulary smaller than the beam");
    PK_CHECK_ARG(ldv >= V && row_lse != nullptr, "logits pitch smaller than V, or no row log-sum-exp");
    // (two CTAs per utterance -- a cluster, each scanning every other row, top-K handed over through distributed shared memory -- measured
    //  slower than one: 54.5 vs 49 us; the scan was bound by the divergent list insertion, not by the number of SMs)
    switch (K) {
#define CASE(KK) case KK: beam_advance_kernel<KK><<<B, ADV_THREADS, 0, ST(stream)>>>(logits, ldv, row_lse, sm_scale, t_idx, num_frames, max_len, st, B, V, L, cap, step, blk, n_best, beam_prune, lm); break;
        CASE(1) CASE(2) CASE(4) CASE(8) CASE(16)
#undef CASE
        default: PK_CHECK_ARG(false, "beam size must be 1, 2, 4, 8 or 16");
    }
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
extern "C" int pk_beam_advance(const float* logits, int ldv, const float* row_lse, float sm_scale, const int* t_idx, const int* num_frames, const int* max_len, float* scores,
                               int* next_ys, int* prev_ks, int* hyp_tok, int* hyp_len, float* fin_score, int* fin_step, int* fin_k,
                               int* fin_count, int* eos_top, int* done, int* not_done_total, int B, int K, int V, int L, int cap,
                               const int* step, int blk, int n_best, int beam_prune, void* stream) {
    LmArgs lm{};
    return beam_advance_impl(logits, ldv, row_lse, sm_scale, t_idx, num_frames, max_len, scores, next_ys, prev_ks, hyp_tok, hyp_len, fin_score, fin_step, fin_k,
                             fin_count, eos_top, done, not_done_total, B, K, V, L, cap, step, blk, n_best, beam_prune, lm, stream);
}
extern "C" int pk_beam_advance_lm(const float* logits, int ldv, const float* row_lse, float sm_scale, const int* t_idx, const int* num_frames, const int* max_len, float* scores,
                                  int* next_ys, int* prev_ks, int* hyp_tok, int* hyp_len, float* fin_score, int* fin_step, int* fin_k,
                                  int* fin_count, int* eos_top, int* done, int* not_done_total, int B, int K, int V, int L, int cap,
                                  const int* step, int blk, int n_best, int beam_prune, const pk_lm_fst* fst, double lm_scale, double nonblk_reward,
                                  int* set_state, double* set_cost, int* set_n, float* lm_scores, int max_states, int* err_flag,
                                  void* stream) {
    PK_CHECK_ARG(fst != nullptr && fst->arc_off != nullptr, "null FST");
    PK_CHECK_ARG(fst->n_disambig >= 0 && fst->n_disambig <= LM_MAX_DISAMBIG, "at most 4 disambiguation labels");
    PK_CHECK_ARG(max_states >= 1, "max_states must be >= 1");
    LmArgs lm{};
    lm.arc_off = fst->arc_off; lm.arc_il = fst->arc_ilabel; lm.arc_w = fst->arc_weight; lm.arc_ns = fst->arc_next; lm.finals = fst->finals;
    lm.backoff_id = fst->backoff_id; lm.n_disambig = fst->n_disambig;
    for (int i = 0; i < fst->n_disambig; ++i) lm.disambig[i] = fst->disambig_ids[i];
    lm.scale = (float)lm_scale; lm.scale_d = lm_scale; lm.reward = nonblk_reward;
    lm.set_state = set_state; lm.set_cost = set_cost; lm.set_n = set_n; lm.lm_scores = lm_scores; lm.MS = max_states; lm.err = err_flag;
    return beam_advance_impl(logits, ldv, row_lse, sm_scale, t_idx, num_frames, max_len, scores, next_ys, prev_ks, hyp_tok, hyp_len, fin_score, fin_step, fin_k,
                             fin_count, eos_top, done, not_done_total, B, K, V, L, cap, step, blk, n_best, beam_prune, lm, stream);
}
extern "C" int pk_beam_reorder(const int* prev_ks, const int* step_ctx, const void* h_in, const float* c_in, const int* t_in, void* h_out, float* c_out,
                               int* t_out, int dtype, int K, int H, int layers, int rows, void* stream) {
    const StepCtx ctx{step_ctx};
    if (dtype == PK_BF16)
        beam_reorder_kernel<__nv_bfloat16><<<rows, 128, 0, ST(stream)>>>(prev_ks, ctx, (const __nv_bfloat16*)h_in, c_in, t_in, (__nv_bfloat16*)h_out, c_out, t_out, K, H, layers, rows);
    else
        beam_reorder_kernel<float><<<rows, 128, 0, ST(stream)>>>(prev_ks, ctx, (const float*)h_in, c_in, t_in, (float*)h_out, c_out, t_out, K, H, layers, rows);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
/* end of a beam step: step_ctx[0] += 1 and step_ctx[1] = (utterances not done > 0 && step < max_steps), both only while the loop is live */
extern "C" int pk_beam_step_end(int* step_ctx, const int* not_done, int max_steps, void* stream) {
    beam_step_end_kernel<<<1, 32, 0, ST(stream)>>>(step_ctx, not_done, max_steps);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
