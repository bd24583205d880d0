// On-the-fly front end on the GPU: speed perturbation + RMS gain + int16 quantisation, Kaldi-compatible
// fbank, splice + padding + CMN/CMVN + SpecAugment.  One batch = a handful of launches.
//
//   loader/audio.py:28-36,207-262,551-603      AudioSegment.{change_speed, normalize, gain_db, rms_db, _convert_*}
//   loader/otf_utt_loader.py:195-201,218-234   augmentation chain + PyKaldi Fbank.compute_features (egs/fbank.conf)
//   loader/otf_utt_loader.py:28-46,262-270     splice +-ctx with edge replication; pad with the last valid frame
//   trainer/train_transducer_bmuf_otfaug.py:86-93 + utils/spec_augment.py:10-20   CMN (padded axis), CMVN, SpecAugment
//
// Integer path (augmented int16 samples): the speed-perturbed branch follows numpy's float64 arithmetic
// (no FMA contraction) and is bit-exact up to the summation order of the mean square (1e-16 relative);
// the rate == 1.0 branch stays in float32 like numpy, where the mean square is accumulated in float64
// here versus numpy's float32 pairwise sum, so individual samples may differ by 1 LSB.
#include <math.h>

#include "../../include/pika_b200.h"
#include "common.cuh"

namespace pk {
void count_launch();

constexpr int FB_FRAME = 400, FB_SHIFT = 160, FB_NFFT = 512, FB_BINS = 256;
constexpr int AUG_THREADS = 256;

// ------------------------------------------------------------------------------------ augmentation
// pass A: resample (float64, numpy.interp semantics) and per-CTA partial sums of squares
__global__ void __launch_bounds__(AUG_THREADS) aug_resample_kernel(const short* __restrict__ pcm, long long ld_pcm,
                                                                   const int* __restrict__ n_samples, const float* __restrict__ rate,
                                                                   const int* __restrict__ new_len, double* __restrict__ resampled,
                                                                   long long ld_res, double* __restrict__ partial, int parts) {
    const int b = blockIdx.y;
    const int N = n_samples[b], L = new_len[b];
    const short* src = pcm + (long long)b * ld_pcm;
    double* dst = resampled + (long long)b * ld_res;
    const bool unit = (rate[b] == 1.0f);
    const double step = (L > 1) ? (double)N / (double)(L - 1) : 0.0;      // numpy.linspace(0, N, L)
    double acc = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
        double v;
        if (unit) {
            const float s = (float)src[j] * (1.0f / 32768.0f);
            v = (double)s;
            const float sq = s * s;                                      // numpy: float32 array ** 2
            acc += (double)sq;
        } else {
            const double x = (j == L - 1) ? (double)N : __dmul_rn((double)j, step);
            if (x >= (double)(N - 1)) {
                v = (double)((float)src[N - 1] * (1.0f / 32768.0f));     // right of the last knot: fp[-1]
            } else {
                const int i = (int)x;                                    // floor, x >= 0
                const double f0 = (double)((float)src[i] * (1.0f / 32768.0f));
                const double f1 = (double)((float)src[i + 1] * (1.0f / 32768.0f));
                const double slope = __dsub_rn(f1, f0);                  // (f1-f0)/(1.0)
                v = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, (double)i)), f0);
            }
            acc += __dmul_rn(v, v);
        }
        dst[j] = v;
    }
    // deterministic block reduction (fixed tree), one partial per CTA
    __shared__ double red[AUG_THREADS];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = AUG_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(long long)b * parts + blockIdx.x] = red[0];
}

// pass B: gain from the mean square, apply, quantise to int16 (stored as float for the fbank kernel)
__global__ void __launch_bounds__(AUG_THREADS) aug_gain_kernel(const double* __restrict__ resampled, long long ld_res,
                                                               const double* __restrict__ partial, int parts,
                                                               const float* __restrict__ rate, const int* __restrict__ new_len,
                                                               const float* __restrict__ target_db, float* __restrict__ wave,
                                                               short* __restrict__ wave_i16, long long ld_wave, int* __restrict__ err_flag) {
    const int b = blockIdx.y;
    const int L = new_len[b];
    __shared__ double s_gain;
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < parts; ++i) tot += partial[(long long)b * parts + i];
        double ms = (L > 0) ? tot / (double)L : 0.0;
        const bool unit = (rate[b] == 1.0f);
        if (unit) ms = (double)(float)ms;                                // numpy float32 mean
        ms = fmax(1e-20, ms);
        const double rms_db = 10.0 * log10(ms);
        double gain_db = (double)target_db[b] - rms_db;
        if (gain_db > 300.0) { atomicExch(err_flag, 1); gain_db = 300.0; }  // reference raises ValueError
        s_gain = pow(10.0, gain_db / 20.0);
    }
    __syncthreads();
    const bool unit = (rate[b] == 1.0f);
    const double gain = s_gain;
    const float gain32 = (float)gain;
    const double* src = resampled + (long long)b * ld_res;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
        int q;
        if (unit) {
            float v = __fmul_rn((float)src[j], gain32);
            v = __fmul_rn(v, 32768.0f);
            v = fminf(fmaxf(v, -32768.0f), 32767.0f);
            q = (int)v;                                                  // C cast: truncation toward zero
        } else {
            double v = __dmul_rn(src[j], gain);
            v = __dmul_rn(v, 32768.0);
            v = fmin(fmax(v, -32768.0), 32767.0);
            q = (int)v;
        }
        wave[(long long)b * ld_wave + j] = (float)q;
        if (wave_i16) wave_i16[(long long)b * ld_wave + j] = (short)q;
    }
}

// ------------------------------------------------------------------------------------ fbank
struct FbankTables {
    const float* window;    // [400] Hamming
    const float2* twiddle;  // [256] exp(-2 pi i k / 512)
    const float* mel_w;     // [n_mel][256]
    const int* mel_lo;      // [n_mel] first non-zero bin
    const int* mel_hi;      // [n_mel] one past the last non-zero bin
};

// standard normal from two counter-based hashes (Box-Muller); one draw per (utterance, frame, sample), as Kaldi's Dither() draws
// one RandGauss() per sample of every extracted window (feat/feature-window.cc: Dither) -- its RNG stream itself is not reproducible
PK_DEVICE float dither_gauss(uint64_t idx, uint32_t seed) {
    const uint32_t h1 = hash_u32(idx * 2, seed), h2 = hash_u32(idx * 2 + 1, seed ^ 0x6A09E667u);
    const float u1 = ((float)h1 + 1.0f) * 2.3283064365386963e-10f;          // (0, 1]
    const float u2 = (float)h2 * 2.3283064365386963e-10f;
    return sqrtf(-2.f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

__global__ void __launch_bounds__(256) fbank_kernel(const float* __restrict__ wave, long long ld_wave, const int* __restrict__ n_frames,
                                                    FbankTables tb, int n_mel, float preemph, float* __restrict__ feats,
                                                    long long ld_b, int t_max, float dither, uint32_t dither_seed) {
    const int t = blockIdx.x, b = blockIdx.y;
    if (t >= n_frames[b]) return;
    __shared__ float2 buf[FB_NFFT];
    __shared__ float frame[FB_FRAME];
    __shared__ float red[8];
    __shared__ float power[FB_BINS];
    const int tid = threadIdx.x;
    const float* src = wave + (long long)b * ld_wave + (long long)t * FB_SHIFT;
    float part = 0.f;
    for (int i = tid; i < FB_FRAME; i += 256) {
        float v = src[i];
        if (dither != 0.f) v += dither * dither_gauss(((uint64_t)b * (uint64_t)t_max + (uint64_t)t) * FB_FRAME + (uint64_t)i, dither_seed);
        frame[i] = v; part += v;
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) mean += red[w];
    mean *= (1.0f / FB_FRAME);
    // DC removal, pre-emphasis (x[i] -= c*x[i-1], x[0] -= c*x[0]), window, zero-pad, bit-reversed placement
    for (int i = tid; i < FB_NFFT; i += 256) {
        float v = 0.f;
        if (i < FB_FRAME) {
            const float cur = frame[i] - mean;
            const float prev = (i > 0 ? frame[i - 1] : frame[0]) - mean;
            v = (cur - preemph * prev) * tb.window[i];
        }
        const int r = __brev((unsigned)i) >> (32 - 9);
        buf[r] = make_float2(v, 0.f);
    }
    __syncthreads();
    // 9 radix-2 stages, 256 butterflies each (one per thread)
#pragma unroll
    for (int s = 1; s <= 9; ++s) {
        const int half = 1 << (s - 1);
        const int grp = tid >> (s - 1), pos = tid & (half - 1);
        const int i0 = grp * (half << 1) + pos, i1 = i0 + half;
        const float2 w = tb.twiddle[pos << (9 - s)];
        const float2 a = buf[i0], c = buf[i1];
        const float2 wc = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
        buf[i0] = make_float2(a.x + wc.x, a.y + wc.y);
        buf[i1] = make_float2(a.x - wc.x, a.y - wc.y);
        __syncthreads();
    }
    power[tid] = buf[tid].x * buf[tid].x + buf[tid].y * buf[tid].y;
    __syncthreads();
    if (tid < n_mel) {
        float e = 0.f;
        const float* w = tb.mel_w + (long long)tid * FB_BINS;
        for (int k = tb.mel_lo[tid]; k < tb.mel_hi[tid]; ++k) e += w[k] * power[k];
        feats[(long long)b * ld_b + (long long)t * n_mel + tid] = logf(fmaxf(e, 1.1920928955078125e-07f));
    }
}

// ------------------------------------------------------------------------------------ splice / CMN / CMVN / SpecAugment
PK_DEVICE float spliced_value(const float* __restrict__ fb, int n_frames, int n_mel, int lctx, int t, int col) {
    // frame t of the padded batch: rows >= n_frames replicate the last valid SPLICED frame
    const int tt = min(t, n_frames - 1);
    const int k = col / n_mel, c = col - k * n_mel;
    int src = tt + k - lctx;
    src = max(0, min(n_frames - 1, src));
    return fb[(long long)src * n_mel + c];
}
__global__ void splice_colsum_kernel(const float* __restrict__ feats, long long ld_b, const int* __restrict__ n_frames, int n_mel,
                                     int lctx, int D, int t_max, float* __restrict__ sums) {
    const int b = blockIdx.y, col = threadIdx.x;
    if (col >= D || n_frames[b] <= 0) return;
    const int t0 = blockIdx.x * 64, t1 = min(t_max, t0 + 64);
    const float* fb = feats + (long long)b * ld_b;
    float s = 0.f;
    for (int t = t0; t < t1; ++t) s += spliced_value(fb, n_frames[b], n_mel, lctx, t, col);
    atomicAdd(&sums[(long long)b * D + col], s);
}
template <typename T>
__global__ void splice_finalize_kernel(const float* __restrict__ feats, long long ld_b, const int* __restrict__ n_frames, int n_mel,
                                       int lctx, int D, int t_max, const float* __restrict__ sums, int cmn,
                                       const float* __restrict__ offset, const float* __restrict__ scale, int f0, int fs, int t0m,
                                       int ts, T* __restrict__ out) {
    const int b = blockIdx.y;
    const long long total = (long long)t_max * D;
    const float* fb = feats + (long long)b * ld_b;
    const int nf = n_frames[b];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / D), col = (int)(i - (long long)t * D);
        float v = nf > 0 ? spliced_value(fb, nf, n_mel, lctx, t, col) : 0.f;
        if (cmn) v -= sums[(long long)b * D + col] / (float)t_max;
        if (offset) { v += offset[col]; v *= scale[col]; }
        if ((fs > 0 && col >= f0 && col < f0 + fs) || (ts > 0 && t >= t0m && t < t0m + ts)) v = 0.f;
        out[(long long)b * total + i] = from_f32<T>(v);
    }
}
}  // namespace pk

using namespace pk;

/* Workspace layout of pk_frontend_fwd: resampled f64 [B, n_max] | partial f64 [B, parts] | wave f32 [B, n_max] |
 * feats f32 [B, t_max, n_mel] | sums f32 [B, D] | err int */
static const int kAugParts = 64;
extern "C" long long pk_frontend_workspace_bytes(int B, int n_max, int t_max, int n_mel, int D) {
    long long b = 0;
    b += (long long)B * n_max * 8 + (long long)B * kAugParts * 8;
    b += (long long)B * n_max * 4;
    b += (long long)B * t_max * n_mel * 4;
    b += (long long)B * D * 4 + 256;
    return b + 1024;
}

extern "C" int pk_frontend_fwd(const short* pcm, long long ld_pcm, const int* n_samples, const float* rate, const int* new_len,
                               const float* target_db, const int* n_frames, int B, int n_max, int t_max, int n_mel, int lctx,
                               int rctx, const float* window, const float* twiddle, const float* mel_w, const int* mel_lo,
                               const int* mel_hi, float preemph, int cmn, const float* offset, const float* scale, int f0, int fs,
                               int t0, int ts, void* out, int out_dtype, short* wave_i16_out, void* workspace,
                               long long workspace_bytes, int* err_flag, float dither, unsigned int dither_seed, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int D = n_mel * (lctx + 1 + rctx);
    PK_CHECK_ARG(B > 0 && n_max >= FB_FRAME && t_max > 0 && n_mel > 0 && n_mel <= 256 && D <= 1024, "bad frontend dims");
    PK_CHECK_ARG(workspace_bytes >= pk_frontend_workspace_bytes(B, n_max, t_max, n_mel, D), "frontend workspace too small");
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    double* resampled = reinterpret_cast<double*>(w); w += (long long)B * n_max * 8;
    double* partial = reinterpret_cast<double*>(w);   w += (long long)B * kAugParts * 8;
    float* wave = reinterpret_cast<float*>(w);        w += (long long)B * n_max * 4;
    float* feats = reinterpret_cast<float*>(w);       w += (long long)B * t_max * n_mel * 4;
    float* sums = reinterpret_cast<float*>(w);
    dim3 ga(kAugParts, B);
    aug_resample_kernel<<<ga, AUG_THREADS, 0, st>>>(pcm, ld_pcm, n_samples, rate, new_len, resampled, n_max, partial, kAugParts);
    PK_CHECK_LAUNCH(); count_launch();
    aug_gain_kernel<<<ga, AUG_THREADS, 0, st>>>(resampled, n_max, partial, kAugParts, rate, new_len, target_db, wave, wave_i16_out, n_max,
                                               err_flag);
    PK_CHECK_LAUNCH(); count_launch();
    FbankTables tb{window, reinterpret_cast<const float2*>(twiddle), mel_w, mel_lo, mel_hi};
    fbank_kernel<<<dim3(t_max, B), 256, 0, st>>>(wave, n_max, n_frames, tb, n_mel, preemph, feats, (long long)t_max * n_mel, t_max, dither,
                                                 dither_seed);
    PK_CHECK_LAUNCH(); count_launch();
    if (cmn) {
        PK_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * B * D, st));
        splice_colsum_kernel<<<dim3((t_max + 63) / 64, B), ((D + 31) / 32) * 32, 0, st>>>(feats, (long long)t_max * n_mel, n_frames, n_mel, lctx,
                                                                                        D, t_max, sums);
        PK_CHECK_LAUNCH(); count_launch();
    }
    const int gx = (int)(((long long)t_max * D + 255) / 256);
    if (out_dtype == PK_BF16)
        splice_finalize_kernel<__nv_bfloat16><<<dim3(gx, B), 256, 0, st>>>(feats, (long long)t_max * n_mel, n_frames, n_mel, lctx, D, t_max,
                                                                          sums, cmn, offset, scale, f0, fs, t0, ts,
                                                                          reinterpret_cast<__nv_bfloat16*>(out));
    else
        splice_finalize_kernel<float><<<dim3(gx, B), 256, 0, st>>>(feats, (long long)t_max * n_mel, n_frames, n_mel, lctx, D, t_max, sums, cmn,
                                                                  offset, scale, f0, fs, t0, ts, reinterpret_cast<float*>(out));
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}

/* feats-only entry (fbank of already-augmented int16-scaled samples), used by parity tests and by
 * utils/compute_global_cmvn-style tooling: wave f32 [B, ld_wave] -> feats f32 [B, t_max, n_mel]. */
extern "C" int pk_fbank(const float* wave, long long ld_wave, const int* n_frames, int B, int t_max, int n_mel, const float* window,
                        const float* twiddle, const float* mel_w, const int* mel_lo, const int* mel_hi, float preemph, float* feats,
                        float dither, unsigned int dither_seed, void* stream) {
    FbankTables tb{window, reinterpret_cast<const float2*>(twiddle), mel_w, mel_lo, mel_hi};
    fbank_kernel<<<dim3(t_max, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(wave, ld_wave, n_frames, tb, n_mel, preemph, feats,
                                                                                  (long long)t_max * n_mel, t_max, dither, dither_seed);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
