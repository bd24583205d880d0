/* Oracle (TEST INFRASTRUCTURE ONLY): RNN-T loss + gradient w.r.t. the joint logits, plain C.
 *
 * Restates, for sizes the numpy oracle (oracle/rnnt.py) is too slow for, the path
 *   trainer/model/transducer.py:110-111   out = F.log_softmax(out, dim=-1)
 *   trainer/train_transducer_bmuf_otfaug.py:97-99   loss = RNNTLoss.apply(...).sum()
 * (the loss arithmetic itself is the un-vendored warp_rnnt; recurrences in oracle/rnnt.py).
 * Double precision throughout; inputs are fp32 logits [B, T, U1, ldv] (first V of ldv valid).
 * Build: see oracle/Makefile.  Never linked into the product library.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double lse2(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}

/* returns 0 on success. dlogits may be NULL (loss only). */
int oracle_rnnt_loss(const float* logits, const int* labels, const int* frame_lens,
                     const int* label_lens, int B, int Tmax, int U1max, int V, int ldv,
                     int ld_labels, double* costs, float* dlogits) {
    size_t node_stride = (size_t)ldv;
    double* lse = (double*)malloc(sizeof(double) * (size_t)Tmax * U1max);
    double* alpha = (double*)malloc(sizeof(double) * (size_t)Tmax * U1max);
    double* beta = (double*)malloc(sizeof(double) * (size_t)Tmax * U1max);
    if (!lse || !alpha || !beta) return -1;
    for (int n = 0; n < B; ++n) {
        const int T = frame_lens[n], U = label_lens[n];
        const float* z = logits + (size_t)n * Tmax * U1max * node_stride;
        const int* y = labels + (size_t)n * ld_labels;
        for (int t = 0; t < T; ++t)
            for (int u = 0; u <= U; ++u) {
                const float* r = z + ((size_t)t * U1max + u) * node_stride;
                double m = r[0];
                for (int v = 1; v < V; ++v) if (r[v] > m) m = r[v];
                double s = 0;
                for (int v = 0; v < V; ++v) s += exp((double)r[v] - m);
                lse[t * U1max + u] = m + log(s);
            }
#define LPB(t, u) ((double)z[((size_t)(t) * U1max + (u)) * node_stride + 0] - lse[(t) * U1max + (u)])
#define LPL(t, u) ((double)z[((size_t)(t) * U1max + (u)) * node_stride + y[u]] - lse[(t) * U1max + (u)])
        for (int t = 0; t < T; ++t)
            for (int u = 0; u <= U; ++u) {
                if (t == 0 && u == 0) { alpha[0] = 0; continue; }
                double a = t > 0 ? alpha[(t - 1) * U1max + u] + LPB(t - 1, u) : -INFINITY;
                double b = u > 0 ? alpha[t * U1max + u - 1] + LPL(t, u - 1) : -INFINITY;
                alpha[t * U1max + u] = lse2(a, b);
            }
        for (int t = T - 1; t >= 0; --t)
            for (int u = U; u >= 0; --u) {
                if (t == T - 1 && u == U) { beta[t * U1max + u] = LPB(t, u); continue; }
                double a = t < T - 1 ? beta[(t + 1) * U1max + u] + LPB(t, u) : -INFINITY;
                double b = u < U ? beta[t * U1max + u + 1] + LPL(t, u) : -INFINITY;
                beta[t * U1max + u] = lse2(a, b);
            }
        const double ll = beta[0];
        costs[n] = -ll;
        if (!dlogits) continue;
        float* dz = dlogits + (size_t)n * Tmax * U1max * node_stride;
        memset(dz, 0, sizeof(float) * (size_t)Tmax * U1max * node_stride);
        for (int t = 0; t < T; ++t)
            for (int u = 0; u <= U; ++u) {
                double a = alpha[t * U1max + u];
                double bn = (t < T - 1) ? beta[(t + 1) * U1max + u] : (u == U ? 0.0 : -INFINITY);
                double gb = -exp(a + bn + LPB(t, u) - ll);          /* d cost / d lp[blank] */
                double gl = 0;
                if (u < U) gl = -exp(a + beta[t * U1max + u + 1] + LPL(t, u) - ll);
                if (!(gb == gb)) gb = 0;
                if (!(gl == gl)) gl = 0;
                double gs = gb + gl;
                const float* r = z + ((size_t)t * U1max + u) * node_stride;
                float* d = dz + ((size_t)t * U1max + u) * node_stride;
                double l = lse[t * U1max + u];
                for (int v = 0; v < V; ++v) d[v] = (float)(-exp((double)r[v] - l) * gs);
                d[0] += (float)gb;
                if (u < U) d[y[u]] += (float)gl;
            }
#undef LPB
#undef LPL
    }
    free(lse); free(alpha); free(beta);
    return 0;
}
