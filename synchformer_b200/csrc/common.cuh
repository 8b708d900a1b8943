// Shared helpers for the synchformer_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/synchformer_b200.h"

namespace sfb {

constexpr int kD = 768;  // embedding width of every stream on the path

void set_error(const char *fmt, ...);

#define SFB_CHECK_ARG(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            sfb::set_error(__VA_ARGS__);  \
            return SFB_E_INVALID;         \
        }                                 \
    } while (0)

#define SFB_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            sfb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SFB_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

#define SFB_CHECK_LAUNCH() SFB_CHECK_CUDA(cudaGetLastError())

int num_sms();

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Exact-erf GELU for kernel epilogues: erfc via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on erf, i.e. fp32 round-off level;
// measured max |gelu error| 4.7e-7 over [-12, 12]), branch-free, 2 MUFU + ~12 FMA-pipe instructions instead of erff's ~30.
//   gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2),   erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1 / (1 + p z)
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float ax = fabsf(x);
    const float z = ax * 0.70710678118654752f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(t, poly, 1.421413741f);
    poly = fmaf(t, poly, -0.284496736f);
    poly = fmaf(t, poly, 0.254829592f);
    poly *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
    return fmaxf(x, 0.0f) - 0.5f * ax * poly * e;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162 *>(&u);
    return __bfloat1622float2(t);
}

}  // namespace sfb
