// ls_adam.cu -- fused AdamUniform step (sm_100a).  Replaces largesteps/optimize.py:17-41 (8 eager torch kernels
// + a max reduction per parameter): two streaming passes over (param, grad, g1, g2).
//   pass 1: g1 = b1 g1 + (1-b1) g ; g2 = b2 g2 + (1-b2) g^2 ; gmax = max(g2)        (optimize.py:35-36)
//   pass 2: p -= lr * (g1/c1) / (1e-8 + sqrt(gmax/c2))                                 (optimize.py:37-41)
// max(sqrt(g2/c2)) == sqrt(max(g2)/c2) in fp32 (both maps are monotone), so the scalar normaliser of optimize.py:40 is
// reproduced to the last bit or one ulp (torch divides by a Python scalar as multiply-by-reciprocal; here it is a division);
// the max itself is order independent (deterministic).  A NaN in the moments makes the normaliser NaN, as torch's max() does:
// divergence poisons every parameter and is visible, instead of being dropped by fmaxf.
#include "ls_common.cuh"

namespace {
constexpr int AT = 256;

__global__ void __launch_bounds__(AT) k_adam_moments(const float *__restrict__ grad, float *__restrict__ g1,
                                                     float *__restrict__ g2, int64_t n, float b1, float b2,
                                                     float omb1, float omb2, unsigned int *__restrict__ gmax_bits) {
    float m = 0.f;
    bool nan_seen = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i];
        // g1.mul_(b1).add_(grad, alpha=1-b1): the in-place mul rounds, torch's CUDA add(alpha) contracts to an fma
        const float a = fmaf(omb1, g, __fmul_rn(g1[i], b1));
        const float b = fmaf(omb2, __fmul_rn(g, g), __fmul_rn(g2[i], b2));
        g1[i] = a;
        g2[i] = b;
        m = fmaxf(m, b);
        nan_seen |= (b != b);
    }
    if (__any_sync(0xffffffffu, nan_seen) && (threadIdx.x & 31) == 0) atomicOr(gmax_bits + 1, 1u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float wm[AT / 32];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < AT / 32 ? wm[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (threadIdx.x == 0) atomicMax(gmax_bits, __float_as_uint(v));   // g2 >= 0: uint order == float order
    }
}

__global__ void __launch_bounds__(AT) k_adam_apply(float *__restrict__ param, const float *__restrict__ g1, int64_t n,
                                                   float lr, float c1, float c2,
                                                   const unsigned int *__restrict__ gmax_bits) {
    const float gmax = gmax_bits[1] ? __int_as_float(0x7fc00000) : __uint_as_float(*gmax_bits);
    const float denom = __fadd_rn(1e-8f, __fsqrt_rn(__fdiv_rn(gmax, c2)));   // 1e-8 + m2.sqrt().max()
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float m1 = __fdiv_rn(g1[i], c1);
        const float gr = __fdiv_rn(m1, denom);
        param[i] = fmaf(-lr, gr, param[i]);                                  // p.data.sub_(gr, alpha=lr)
    }
}
}  // namespace

extern "C" int ls_adam_uniform_step(float *param, const float *grad, float *g1, float *g2, int64_t n, float lr,
                                    float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float c1,
                                    float c2, void *scratch, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LS_REQUIRE(n >= 0, "negative size");
    if (n == 0) return LS_OK;
    LS_REQUIRE(param && grad && g1 && g2 && scratch, "NULL pointer");
    LsDevInfo di;
    int rc = ls_dev_info(&di);
    if (rc) return rc;
    int64_t g = (n + AT - 1) / AT;
    if (g > (int64_t)di.sm_count * 8) g = (int64_t)di.sm_count * 8;
    LS_CUDA_TRY(cudaMemsetAsync(scratch, 0, 16, stream));
    k_adam_moments<<<(unsigned)g, AT, 0, stream>>>(grad, g1, g2, n, beta1, beta2, one_minus_beta1, one_minus_beta2,
                                                   (unsigned int *)scratch);
    LS_LAUNCH_CHECK();
    k_adam_apply<<<(unsigned)g, AT, 0, stream>>>(param, g1, n, lr, c1, c2, (const unsigned int *)scratch);
    LS_LAUNCH_CHECK();
    return LS_OK;
}
