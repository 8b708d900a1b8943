// K4: LayerNorm over D = 768, one warp per row, fp32 statistics (two-pass in registers), bf16 or fp32 output.
// HBM-bound: reads 3072 B and writes 1536 B (bf16) per row; each lane moves 6 x float4 in, 6 x 8 B out,
// fully coalesced.  Optional row gather (drop CLS / aux tokens) and an optional fused second LayerNorm.
// Replaces nn.LayerNorm at vit_helper.py:366-375 (norm1/2/3), motionformer.py:231, modeling_ast.py:291-292,543,
// modules/transformer.py:84-85 and the norm1/norm2 of nn.TransformerEncoderLayer (motionformer.py:329).
#include "common.cuh"

namespace sfb {

template <bool kOutF32, bool kDouble>
__global__ void __launch_bounds__(256) layernorm768_kernel(const float *__restrict__ x, int64_t ldx, void *__restrict__ out, int64_t ldo,
                                                           const float *__restrict__ g1, const float *__restrict__ b1, float eps1,
                                                           const float *__restrict__ g2, const float *__restrict__ b2, float eps2,
                                                           int rows, int group, int group_stride, int offset) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= rows) return;
    const int64_t in_row = static_cast<int64_t>(r / group) * group_stride + offset + (r % group);
    const float4 *xp = reinterpret_cast<const float4 *>(x + in_row * ldx);
    float4 v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = __ldg(xp + lane + 32 * j);

    auto normalise = [&](const float *__restrict__ g, const float *__restrict__ b, float eps) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        const float mean = warp_sum(s) * (1.0f / kD);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            v[j].x -= mean, v[j].y -= mean, v[j].z -= mean, v[j].w -= mean;
            q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float4 gg = __ldg(reinterpret_cast<const float4 *>(g) + lane + 32 * j);
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(b) + lane + 32 * j);
            v[j].x = v[j].x * rstd * gg.x + bb.x;
            v[j].y = v[j].y * rstd * gg.y + bb.y;
            v[j].z = v[j].z * rstd * gg.z + bb.z;
            v[j].w = v[j].w * rstd * gg.w + bb.w;
        }
    };
    normalise(g1, b1, eps1);
    if (kDouble) normalise(g2, b2, eps2);

    if (kOutF32) {
        float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + static_cast<int64_t>(r) * ldo);
#pragma unroll
        for (int j = 0; j < 6; ++j) op[lane + 32 * j] = v[j];
    } else {
        uint2 *op = reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(out) + static_cast<int64_t>(r) * ldo);
#pragma unroll
        for (int j = 0; j < 6; ++j) op[lane + 32 * j] = make_uint2(pack_bf16x2(v[j].x, v[j].y), pack_bf16x2(v[j].z, v[j].w));
    }
}

// LN_FOLD inputs for rows that no EMIT_LN epilogue produced: bf16 copy + (sum, sum of squares) per row; same data movement as a LayerNorm
__global__ void __launch_bounds__(256) rowstats_cast_kernel(const float *__restrict__ x, int64_t ldx, __nv_bfloat16 *__restrict__ xb,
                                                            float *__restrict__ stats, int rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= rows) return;
    const float4 *xp = reinterpret_cast<const float4 *>(x + static_cast<int64_t>(r) * ldx);
    uint2 *op = reinterpret_cast<uint2 *>(xb + static_cast<int64_t>(r) * kD);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float4 v = __ldg(xp + lane + 32 * j);
        s1 += (v.x + v.y) + (v.z + v.w);
        s2 += fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
        op[lane + 32 * j] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    s1 = warp_sum(s1), s2 = warp_sum(s2);
    if (lane == 0) *reinterpret_cast<float2 *>(stats + static_cast<int64_t>(r) * 2) = make_float2(s1, s2);
}

}  // namespace sfb

extern "C" int sfb_rowstats_cast(const float *x, int64_t ldx, void *xb, float *stats, int rows, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(x && xb && stats && rows > 0, "sfb_rowstats_cast: bad arguments");
    SFB_CHECK_ARG(ldx % 4 == 0 && ldx >= kD && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(xb) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(stats) & 7) == 0, "sfb_rowstats_cast: alignment");
    rowstats_cast_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, reinterpret_cast<__nv_bfloat16 *>(xb), stats, rows);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_layernorm(const float *x, int64_t ldx, void *out, int64_t ldo, int out_f32, const float *gamma, const float *beta,
                             float eps, const float *gamma2, const float *beta2, float eps2, int rows, int group, int group_stride,
                             int offset, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(x && out && gamma && beta, "sfb_layernorm: null pointer");
    SFB_CHECK_ARG(rows > 0 && group > 0, "sfb_layernorm: rows=%d group=%d", rows, group);
    SFB_CHECK_ARG(ldx % 4 == 0 && ldo % 4 == 0 && ldx >= kD && ldo >= kD, "sfb_layernorm: ldx/ldo must be >= 768 and multiples of 4");
    SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0,
                  "sfb_layernorm: pointers must be 16-byte aligned");
    SFB_CHECK_ARG((gamma2 == nullptr) == (beta2 == nullptr), "sfb_layernorm: gamma2/beta2 must be given together");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (rows + 7) / 8;
#define SFB_LN_LAUNCH(F32, DBL) \
    layernorm768_kernel<F32, DBL><<<grid, 256, 0, st>>>(x, ldx, out, ldo, gamma, beta, eps, gamma2, beta2, eps2, rows, group, group_stride, offset)
    if (gamma2) {
        if (out_f32) SFB_LN_LAUNCH(true, true); else SFB_LN_LAUNCH(false, true);
    } else {
        if (out_f32) SFB_LN_LAUNCH(true, false); else SFB_LN_LAUNCH(false, false);
    }
#undef SFB_LN_LAUNCH
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
