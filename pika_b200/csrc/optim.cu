// Optimiser-side kernels on the flat fp32 parameter vector: inf-norm gradient clipping,
// Nesterov SGD, and the BMUF block-momentum update.  Pure HBM streams (float4, grid-stride).
//
//   trainer/train_transducer_bmuf_otfaug.py:105-110  clip_grad_norm_(inf) + SGD(nesterov).step()
//   trainer/bmuf.py:76-100                            BmufTrainer.update_and_sync
#include "../../include/pika_b200.h"
#include "common.cuh"

namespace pk {
void count_launch();

__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out_bits,
                                                     int* __restrict__ nan_flag) {
    float m = 0.f;
    bool bad = false;
    const long long n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x4[i];
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        bad |= !(v.x == v.x) | !(v.y == v.y) | !(v.z == v.z) | !(v.w == v.w);
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        m = fmaxf(m, fabsf(x[i]));
        bad |= !(x[i] == x[i]);
    }
    m = warp_max(m);
    const bool warp_bad = __any_sync(0xffffffffu, bad);          // all 32 lanes reach this point
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out_bits, __float_as_uint(m));     // non-negative floats order like their bit patterns
        if (nan_flag && warp_bad) atomicExch(nan_flag, 1);
    }
}

// g' = g * min(1, max_norm / (absmax + 1e-6));  buf = first ? g' : mom*buf + g';  p -= lr * (g' + mom*buf)
__global__ void __launch_bounds__(256) sgd_nesterov_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                                           long long n, float lr, float mom, float max_norm,
                                                           const unsigned int* __restrict__ absmax_bits,
                                                           const int* __restrict__ nan_flag, int first) {
    float coef = 1.f;
    if (max_norm > 0.f && absmax_bits) {
        const float tot = __uint_as_float(*absmax_bits);
        coef = fminf(1.f, max_norm / (tot + 1e-6f));
        // fmaxf drops NaNs, torch.max does not: clip_grad_norm_(inf) of a gradient holding a NaN gives a NaN coefficient,
        // i.e. every parameter turns NaN (and the next BMUF sync stops the run); reproduced through the flag absmax raised
        if (nan_flag && *nan_flag) coef = __int_as_float(0x7fc00000);
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        const float b = first ? gi : mom * buf[i] + gi;
        buf[i] = b;
        p[i] -= lr * (gi + mom * b);
    }
}

// delta = global - local  (the vector that is sum-reduced across ranks)
__global__ void __launch_bounds__(256) bmuf_delta_kernel(const float* __restrict__ glob, const float* __restrict__ local,
                                                         float* __restrict__ delta, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        delta[i] = glob[i] - local[i];
}
// d = dsum / N; dprev = bm*dprev + blr*(1-bm)*d; glob -= (1+bm)*dprev; local = glob
__global__ void __launch_bounds__(256) bmuf_update_kernel(float* __restrict__ glob, float* __restrict__ local,
                                                          float* __restrict__ dprev, const float* __restrict__ dsum, long long n,
                                                          float inv_world, float bm, float blr) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = dsum[i] * inv_world;
        const float dp = bm * dprev[i] + blr * (1.f - bm) * d;
        dprev[i] = dp;
        const float gnew = glob[i] - (1.f + bm) * dp;
        glob[i] = gnew;
        local[i] = gnew;
    }
}
}  // namespace pk

using namespace pk;
static int og(long long n) {
    long long g = (n + 1023) / 1024, cap = (long long)num_sms() * 8;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

/* out_bits[0] must be zeroed by the caller (cudaMemsetAsync) -- done here. nan_flag (int*, may be NULL) is set to 1 on NaN. */
extern "C" int pk_absmax(const float* x, long long n, float* out, int* nan_flag, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    PK_CHECK_CUDA(cudaMemsetAsync(out, 0, 4, st));
    absmax_kernel<<<og(n), 256, 0, st>>>(x, n, reinterpret_cast<unsigned int*>(out), nan_flag);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
extern "C" int pk_sgd_nesterov_clip(float* p, const float* g, float* buf, long long n, float lr, float momentum, float max_norm,
                                    const float* absmax, const int* nan_flag, int first, void* stream) {
    sgd_nesterov_kernel<<<og(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, buf, n, lr, momentum, max_norm,
                                                                                 reinterpret_cast<const unsigned int*>(absmax), nan_flag, first);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
extern "C" int pk_bmuf_delta(const float* glob, const float* local, float* delta, long long n, void* stream) {
    bmuf_delta_kernel<<<og(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(glob, local, delta, n);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
extern "C" int pk_bmuf_update(float* glob, float* local, float* delta_prev, const float* delta_sum, long long n, int world,
                              float block_momentum, float block_lr, void* stream) {
    PK_CHECK_ARG(world >= 1, "world must be >= 1");
    bmuf_update_kernel<<<og(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(glob, local, delta_prev, delta_sum, n, 1.f / world,
                                                                                block_momentum, block_lr);
    PK_CHECK_LAUNCH(); count_launch();
    return 0;
}
