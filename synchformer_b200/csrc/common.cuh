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
