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

int num_sms();           // SM count of the CURRENT device (cached per device ordinal)
// One-time-per-device setup (cudaFuncSetAttribute, constant tables ...): `once.first()` is true exactly once for each device
// ordinal that becomes current, so a process that drives several GPUs sets every device up.
struct PerDeviceOnce {
    bool done[64] = {};
    bool first();
    void reset_current();            // setup failed: try again on the next call
};
int encode_tmap_bf16_2d(void *map, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
int encode_tmap_f32_2d(void *map, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols, bool swizzle128);

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Exact-erf GELU for kernel epilogues, two elements per call on the packed fp32x2 pipe of sm_100 (FFMA2 / FMUL2).
//   gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2)
//   erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1 / (1 + p z)        (Abramowitz-Stegun 7.1.26,
//   |erf error| <= 1.5e-7, i.e. fp32 round-off level; measured max |gelu error| 3.3e-7 over [-12, 12]).
// With zs = |x| sqrt(log2(e) / 2): exp(-z^2) = 2^(-zs^2), p and the 0.5 are folded into the constants.  Branch-free;
// per PAIR of elements: 4 MUFU (2 rcp, 2 ex2) + ~11 packed FMA-pipe instructions, against ~60 for two erff() calls.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu_erf_fast2(float &x0, float &x1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const uint64_t ax = pack_f32x2(a0, a1);
    const uint64_t zs = mul_f32x2(ax, pack_f32x2(0.84932178f, 0.84932178f));
    const uint64_t den = fma_f32x2(zs, pack_f32x2(0.27273747f, 0.27273747f), pack_f32x2(1.0f, 1.0f));
    float d0, d1;
    unpack_f32x2(den, d0, d1);
    const uint64_t t = pack_f32x2(rcp_approx(d0), rcp_approx(d1));
    uint64_t poly = fma_f32x2(t, pack_f32x2(0.53070271f, 0.53070271f), pack_f32x2(-0.72657603f, -0.72657603f));
    poly = fma_f32x2(t, poly, pack_f32x2(0.71070689f, 0.71070689f));
    poly = fma_f32x2(t, poly, pack_f32x2(-0.14224836f, -0.14224836f));
    poly = fma_f32x2(t, poly, pack_f32x2(0.12741479f, 0.12741479f));
    poly = mul_f32x2(poly, t);
    float s0, s1;
    unpack_f32x2(mul_f32x2(zs, zs), s0, s1);
    const uint64_t e = pack_f32x2(ex2_approx(-s0), ex2_approx(-s1));
    float h0, h1;
    unpack_f32x2(mul_f32x2(mul_f32x2(ax, poly), e), h0, h1);
    x0 = fmaxf(x0, 0.0f) - h0;
    x1 = fmaxf(x1, 0.0f) - h1;
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
