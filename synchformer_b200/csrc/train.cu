// N3: training kernels of the synchronisation module (vproj / aproj + GlobalTransformer), i.e. everything the
// backward of modules/transformer.py:58-97 and sync_model.py:150-173 needs besides the GEMMs (which reuse sfb_gemm_bf16 on
// transposed operands) and the attention kernels (attention_train.cu):
//   sfb_dropout           nn.Dropout forward / backward with a counter-based Philox4x32-10 mask (transformer.py:47-48,74,92; sync_model.py:137)
//   sfb_gelu_fwd / _bwd   nn.GELU on the bf16 pre-activation and its derivative (transformer.py:89)
//   sfb_layernorm_bwd     nn.LayerNorm backward over D = 768, fp32 (transformer.py:84-85; sync_model.py:126-127,143)
//   sfb_transpose_bf16    operand transposes for dW = dY^T X and dX = dY W on the K-major tcgen05 GEMM
//   sfb_colsum            bias gradients / batch reductions (column sums with a deterministic two-stage reduction)
//   sfb_sync_head_bwd     backward of ln_f on token 0 + Linear(768 -> n_cls) (sync_model.py:169-172)
// All of them are HBM- or latency-bound passes over (B*T, 768..3072) matrices that are < 1 % of a training step (the frozen
// encoders' forward is > 99 % of its FLOPs), so they are written for coalescing and determinism (no atomics), not for tensor cores.
#include "common.cuh"
#include "philox.cuh"

namespace sfb {

static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ----------------------------------------------------------------------------------------------------------
// dropout: out[e] = (res ? res[e] : 0) + in[e] * (keep(e) ? 1 / (1 - p) : 0), four consecutive elements per thread = one Philox call
// ----------------------------------------------------------------------------------------------------------
template <bool kOutBf16>
__global__ void __launch_bounds__(256) dropout_kernel(const float4 *__restrict__ in, const float4 *__restrict__ res, void *__restrict__ out,
                                                      int64_t n_vec, DropParams dp) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_vec) return;
    float4 v = __ldg(in + idx);
    if (dp.thr != 0) {
        const uint4 r = philox4x32_10(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), dp.site, 0u, dp.seed_lo, dp.seed_hi);
        v.x = r.x >= dp.thr ? v.x * dp.inv_keep : 0.f;
        v.y = r.y >= dp.thr ? v.y * dp.inv_keep : 0.f;
        v.z = r.z >= dp.thr ? v.z * dp.inv_keep : 0.f;
        v.w = r.w >= dp.thr ? v.w * dp.inv_keep : 0.f;
    }
    if (res != nullptr) {
        const float4 rr = __ldg(res + idx);
        v.x += rr.x, v.y += rr.y, v.z += rr.z, v.w += rr.w;
    }
    if (kOutBf16)
        reinterpret_cast<uint2 *>(out)[idx] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    else
        reinterpret_cast<float4 *>(out)[idx] = v;
}

// ----------------------------------------------------------------------------------------------------------
// GELU on bf16 (8 elements per thread).  forward: y = gelu(x);  backward: dx = dy * (Phi(x) + x phi(x))
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, int64_t n_vec) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_vec) return;
    const uint4 u = __ldg(x + idx);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        o[k] = pack_bf16x2(gelu_erf(f.x), gelu_erf(f.y));
    }
    y[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x, uint4 *__restrict__ dx,
                                                       int64_t n_vec) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_vec) return;
    const uint4 ug = __ldg(dy + idx), ux = __ldg(x + idx);
    const uint32_t g[4] = {ug.x, ug.y, ug.z, ug.w}, w[4] = {ux.x, ux.y, ux.z, ux.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 fg = unpack_bf16x2(g[k]), fx = unpack_bf16x2(w[k]);
        o[k] = pack_bf16x2(fg.x * gelu_grad(fx.x), fg.y * gelu_grad(fx.y));
    }
    dx[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ----------------------------------------------------------------------------------------------------------
// bf16 transpose with zero padding: out[c][r] = in[r][c] (r < R, c < C), out[c][r] = 0 for R <= r < ld_out
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16 *__restrict__ in, int64_t ld_in, int R, int C,
                                                             __nv_bfloat16 *__restrict__ out, int64_t ld_out) {
    __shared__ __nv_bfloat16 tile[32][34];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < R && c < C) ? in[static_cast<int64_t>(r) * ld_in + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (c < C && r < ld_out) out[static_cast<int64_t>(c) * ld_out + r] = tile[tx][ty + 8 * k];
    }
}

// ----------------------------------------------------------------------------------------------------------
// column sums: out[c] = sum_r in[r][c].  Stage 1: CTA (64 columns, row slice) -> partial[slice][c]; stage 2 adds the slices.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 load_pair(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ float2 load_pair(const __nv_bfloat16 *p) { return unpack_bf16x2(*reinterpret_cast<const uint32_t *>(p)); }

template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T *__restrict__ in, int64_t ld, int M, int N, float *__restrict__ partial) {
    __shared__ float2 red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 64 + tx * 2;
    float2 acc = make_float2(0.f, 0.f);
    if (col < N) {
        for (int r = blockIdx.y * 8 + ty; r < M; r += 8 * gridDim.y) {
            const float2 v = load_pair(in + static_cast<int64_t>(r) * ld + col);
            acc.x += v.x, acc.y += v.y;
        }
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) acc.x += red[k][tx].x, acc.y += red[k][tx].y;
        *reinterpret_cast<float2 *>(partial + static_cast<int64_t>(blockIdx.y) * N + col) = acc;
    }
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float *__restrict__ partial, int n_part, int64_t n_cols, float *__restrict__ out) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (c >= n_cols) return;
    float s = 0.f;
    for (int k = 0; k < n_part; ++k) s += partial[static_cast<int64_t>(k) * n_cols + c];
    out[c] = s;
}

// ----------------------------------------------------------------------------------------------------------
// LayerNorm backward over D = 768, one warp per row (grid-stride), fp32:
//   xhat = (x - mean) rstd,  g = dy * gamma,  dx = rstd (g - mean(g) - xhat mean(g xhat)),  dgamma += dy xhat,  dbeta += dy
// dy row r is read at (r / group) * group_stride + offset + r % group (the token gather of sync_model.py:164); x / dx rows are r.
// Per-CTA partial dgamma / dbeta go to partial[blockIdx.x][2][768]; reduce_partials_kernel adds them up.
// ----------------------------------------------------------------------------------------------------------
template <bool kAccumulate>
__global__ void __launch_bounds__(256) layernorm768_bwd_kernel(const float *__restrict__ dy, int64_t lddy, int group, int group_stride, int offset,
                                                               const float *__restrict__ x, int64_t ldx, const float *__restrict__ gamma,
                                                               float eps, float *__restrict__ dx, int64_t lddx, float *__restrict__ partial,
                                                               int rows) {
    __shared__ float4 red[8][192];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 gm[6], dg[6], db[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        gm[j] = __ldg(reinterpret_cast<const float4 *>(gamma) + lane + 32 * j);
        dg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
        const int64_t dy_row = static_cast<int64_t>(r / group) * group_stride + offset + (r % group);
        const float4 *xp = reinterpret_cast<const float4 *>(x + static_cast<int64_t>(r) * ldx);
        const float4 *gp = reinterpret_cast<const float4 *>(dy + dy_row * lddy);
        float4 v[6], g[6];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            v[j] = __ldg(xp + lane + 32 * j);
            g[j] = __ldg(gp + lane + 32 * j);
            s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
        const float mean = warp_sum(s) * (1.0f / kD);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            v[j].x -= mean, v[j].y -= mean, v[j].z -= mean, v[j].w -= mean;
            q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            v[j].x *= rstd, v[j].y *= rstd, v[j].z *= rstd, v[j].w *= rstd;                      // xhat
            dg[j].x += g[j].x * v[j].x, dg[j].y += g[j].y * v[j].y, dg[j].z += g[j].z * v[j].z, dg[j].w += g[j].w * v[j].w;
            db[j].x += g[j].x, db[j].y += g[j].y, db[j].z += g[j].z, db[j].w += g[j].w;
            g[j].x *= gm[j].x, g[j].y *= gm[j].y, g[j].z *= gm[j].z, g[j].w *= gm[j].w;          // dy * gamma
            sg += (g[j].x + g[j].y) + (g[j].z + g[j].w);
            sgx += (g[j].x * v[j].x + g[j].y * v[j].y) + (g[j].z * v[j].z + g[j].w * v[j].w);
        }
        const float mg = warp_sum(sg) * (1.0f / kD), mgx = warp_sum(sgx) * (1.0f / kD);
        float4 *op = reinterpret_cast<float4 *>(dx + static_cast<int64_t>(r) * lddx);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            float4 o = make_float4(rstd * (g[j].x - mg - v[j].x * mgx), rstd * (g[j].y - mg - v[j].y * mgx),
                                   rstd * (g[j].z - mg - v[j].z * mgx), rstd * (g[j].w - mg - v[j].w * mgx));
            if (kAccumulate) {
                const float4 old = op[lane + 32 * j];
                o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
            }
            op[lane + 32 * j] = o;
        }
    }
    // CTA reduction of the per-warp column partials (dgamma first, then dbeta, through the same 24 KB buffer)
    float4 *pout = reinterpret_cast<float4 *>(partial + static_cast<int64_t>(blockIdx.x) * 2 * kD);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int j = 0; j < 6; ++j) red[warp][lane + 32 * j] = pass == 0 ? dg[j] : db[j];
        __syncthreads();
        if (threadIdx.x < 192) {
            float4 s = red[0][threadIdx.x];
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                const float4 t = red[k][threadIdx.x];
                s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
            }
            pout[pass * 192 + threadIdx.x] = s;
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------------------
// head backward (sync_model.py:169-172): logits[b] = W LN_f(x[b, 0]) + bias.  One CTA per clip:
//   y = LN_f(x[b,0]);  dyn = dlogits[b] W;  dx[b,0] = LN backward of dyn;  scratch[b] = { y, xhat, dyn } for the parameter gradients
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sync_head_bwd_rows_kernel(const float *__restrict__ x, int T, const float *__restrict__ g, const float *__restrict__ bt,
                                                                 float eps, const float *__restrict__ W, const float *__restrict__ dlogits, int n_cls,
                                                                 float *__restrict__ dx, float *__restrict__ scratch) {
    __shared__ float red[3][8];
    __shared__ float dl[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x;
    const float *row = x + b * T * kD;
    if (threadIdx.x < n_cls) dl[threadIdx.x] = dlogits[b * n_cls + threadIdx.x];
    // each thread owns columns threadIdx.x + 256 k, k = 0..2
    float v[3], s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = row[threadIdx.x + 256 * k], s += v[k];
    s = warp_sum(s);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) mean += red[0][k];
    mean *= (1.0f / kD);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] -= mean, q += v[k] * v[k];
    q = warp_sum(q);
    if (lane == 0) red[1][warp] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) var += red[1][k];
    const float rstd = rsqrtf(var * (1.0f / kD) + eps);
    float gg[3], sg = 0.f, sgx = 0.f;
    float *sc = scratch + b * 3 * kD;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int c = threadIdx.x + 256 * k;
        v[k] *= rstd;                                                       // xhat
        float dyn = 0.f;
        for (int o = 0; o < n_cls; ++o) dyn = fmaf(dl[o], __ldg(W + static_cast<int64_t>(o) * kD + c), dyn);
        sc[c] = v[k] * g[c] + bt[c];
        sc[kD + c] = v[k];
        sc[2 * kD + c] = dyn;
        gg[k] = dyn * g[c];
        sg += gg[k], sgx += gg[k] * v[k];
    }
    sg = warp_sum(sg), sgx = warp_sum(sgx);
    __syncthreads();                                                        // red[0] / red[1] fully consumed above
    if (lane == 0) red[0][warp] = sg, red[2][warp] = sgx;
    __syncthreads();
    float mg = 0.f, mgx = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) mg += red[0][k], mgx += red[2][k];
    mg *= (1.0f / kD), mgx *= (1.0f / kD);
    float *drow = dx + b * T * kD;
#pragma unroll
    for (int k = 0; k < 3; ++k) drow[threadIdx.x + 256 * k] = rstd * (gg[k] - mg - v[k] * mgx);
}

// parameter gradients of the head: thread per column c of 768 (blocks 0..2) -> dgamma, dbeta, dW[:, c]; block 3 -> dbias
__global__ void __launch_bounds__(256) sync_head_bwd_params_kernel(const float *__restrict__ scratch, const float *__restrict__ dlogits, int B, int n_cls,
                                                                   float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ dW,
                                                                   float *__restrict__ dbias) {
    if (blockIdx.x == 3) {
        for (int o = threadIdx.x; o < n_cls; o += 256) {
            float s = 0.f;
            for (int b = 0; b < B; ++b) s += dlogits[static_cast<int64_t>(b) * n_cls + o];
            dbias[o] = s;
        }
        return;
    }
    const int c = blockIdx.x * 256 + threadIdx.x;
    float sgam = 0.f, sbet = 0.f;
    for (int b = 0; b < B; ++b) {
        const float *sc = scratch + static_cast<int64_t>(b) * 3 * kD;
        const float dyn = sc[2 * kD + c];
        sgam += dyn * sc[kD + c];
        sbet += dyn;
    }
    dgamma[c] = sgam, dbeta[c] = sbet;
    for (int o = 0; o < n_cls; ++o) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s = fmaf(dlogits[static_cast<int64_t>(b) * n_cls + o], scratch[static_cast<int64_t>(b) * 3 * kD + c], s);
        dW[static_cast<int64_t>(o) * kD + c] = s;
    }
}

static inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace sfb

extern "C" int sfb_dropout(const float *in, const float *residual, void *out, int out_bf16, int64_t n, float p, uint64_t seed, uint32_t site,
                           void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(in && out && n > 0 && n % 4 == 0, "sfb_dropout: bad arguments (n=%lld must be a positive multiple of 4)", (long long)n);
    SFB_CHECK_ARG(al16(in) && al16(out) && al16(residual), "sfb_dropout: pointers must be 16-byte aligned");
    SFB_CHECK_ARG(p >= 0.f && p < 1.f, "sfb_dropout: p=%f outside [0, 1)", p);
    const DropParams dp = make_drop_params(p, seed, site);
    const int64_t n_vec = n / 4;
    SFB_CHECK_ARG(blocks_for(n_vec) < (1u << 31), "sfb_dropout: too many elements for one launch");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out_bf16)
        dropout_kernel<true><<<blocks_for(n_vec), 256, 0, st>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<const float4 *>(residual), out, n_vec, dp);
    else
        dropout_kernel<false><<<blocks_for(n_vec), 256, 0, st>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<const float4 *>(residual), out, n_vec, dp);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_gelu_fwd(const void *x, void *y, int64_t n, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(x && y && n > 0 && n % 8 == 0 && al16(x) && al16(y), "sfb_gelu_fwd: bad arguments (n %% 8 == 0, 16-byte aligned)");
    gelu_fwd_kernel<<<blocks_for(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4 *>(x), reinterpret_cast<uint4 *>(y), n / 8);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_gelu_bwd(const void *dy, const void *x, void *dx, int64_t n, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(dy && x && dx && n > 0 && n % 8 == 0 && al16(dy) && al16(x) && al16(dx), "sfb_gelu_bwd: bad arguments (n %% 8 == 0, 16-byte aligned)");
    gelu_bwd_kernel<<<blocks_for(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint4 *>(dy), reinterpret_cast<const uint4 *>(x), reinterpret_cast<uint4 *>(dx), n / 8);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_transpose_bf16(const void *in, int64_t ld_in, int R, int C, void *out, int64_t ld_out, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(in && out && R > 0 && C > 0 && ld_in >= C && ld_out >= R, "sfb_transpose_bf16: bad shape R=%d C=%d ld_in=%lld ld_out=%lld", R, C,
                  (long long)ld_in, (long long)ld_out);
    const dim3 grid(static_cast<unsigned>((ld_out + 31) / 32), static_cast<unsigned>((C + 31) / 32));
    SFB_CHECK_ARG(grid.y <= 65535u, "sfb_transpose_bf16: C=%d too large", C);
    transpose_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16 *>(in), ld_in, R, C,
                                                                                    reinterpret_cast<__nv_bfloat16 *>(out), ld_out);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_colsum(const void *in, int in_bf16, int64_t ld, int M, int N, float *out, float *workspace, int64_t workspace_floats,
                          void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(in && out && M > 0 && N > 0 && N % 2 == 0 && ld >= N && ld % 2 == 0, "sfb_colsum: bad shape M=%d N=%d ld=%lld (N, ld even)", M, N, (long long)ld);
    SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "sfb_colsum: pointers must be 8-byte aligned");
    int parts = (M + 63) / 64;                                              // >= 8 rows per thread before another slice is worth it
    if (parts > 64) parts = 64;
    if (workspace == nullptr || workspace_floats < 2 * static_cast<int64_t>(N)) parts = 1;
    else if (static_cast<int64_t>(parts) * N > workspace_floats) parts = static_cast<int>(workspace_floats / N);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid(static_cast<unsigned>((N + 63) / 64), static_cast<unsigned>(parts));
    float *dst = parts == 1 ? out : workspace;
    if (in_bf16)
        colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(in), ld, M, N, dst);
    else
        colsum_partial_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float *>(in), ld, M, N, dst);
    SFB_CHECK_LAUNCH();
    if (parts > 1) {
        reduce_partials_kernel<<<blocks_for(N), 256, 0, st>>>(workspace, parts, N, out);
        SFB_CHECK_LAUNCH();
    }
    return SFB_OK;
}

extern "C" int sfb_layernorm_bwd_workspace_floats(int rows) {
    int grid = (rows + 7) / 8;
    const int cap = 2 * sfb::num_sms();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    return grid * 2 * sfb::kD;
}

extern "C" int sfb_layernorm_bwd(const float *dy, int64_t lddy, int group, int group_stride, int offset, const float *x, int64_t ldx,
                                 const float *gamma, float eps, float *dx, int64_t lddx, int accumulate, float *dgamma, float *dbeta,
                                 float *workspace, int64_t workspace_floats, int rows, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(dy && x && gamma && dx && dgamma && dbeta && workspace, "sfb_layernorm_bwd: null pointer");
    SFB_CHECK_ARG(rows > 0 && group > 0, "sfb_layernorm_bwd: rows=%d group=%d", rows, group);
    SFB_CHECK_ARG(lddy % 4 == 0 && ldx % 4 == 0 && lddx % 4 == 0 && lddy >= kD && ldx >= kD && lddx >= kD, "sfb_layernorm_bwd: leading dimensions must be >= 768 and multiples of 4");
    SFB_CHECK_ARG(al16(dy) && al16(x) && al16(gamma) && al16(dx) && al16(workspace), "sfb_layernorm_bwd: pointers must be 16-byte aligned");
    SFB_CHECK_ARG(dgamma + kD == dbeta, "sfb_layernorm_bwd: dgamma and dbeta must be adjacent (dbeta == dgamma + 768)");
    const int need = sfb_layernorm_bwd_workspace_floats(rows);
    SFB_CHECK_ARG(workspace_floats >= need, "sfb_layernorm_bwd: workspace of %lld floats, need %d", (long long)workspace_floats, need);
    const int grid = need / (2 * kD);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (accumulate)
        layernorm768_bwd_kernel<true><<<grid, 256, 0, st>>>(dy, lddy, group, group_stride, offset, x, ldx, gamma, eps, dx, lddx, workspace, rows);
    else
        layernorm768_bwd_kernel<false><<<grid, 256, 0, st>>>(dy, lddy, group, group_stride, offset, x, ldx, gamma, eps, dx, lddx, workspace, rows);
    SFB_CHECK_LAUNCH();
    reduce_partials_kernel<<<blocks_for(2 * kD), 256, 0, st>>>(workspace, grid, 2 * kD, dgamma);     // [dgamma | dbeta]
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_sync_head_bwd(const float *x, int T, const float *ln_w, const float *ln_b, float eps, const float *W, const float *dlogits,
                                 int B, int n_cls, float *dx, float *dln_w, float *dln_b, float *dW, float *dbias, float *scratch,
                                 void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(x && ln_w && ln_b && W && dlogits && dx && dln_w && dln_b && dW && dbias && scratch, "sfb_sync_head_bwd: null pointer");
    SFB_CHECK_ARG(B > 0 && T > 0 && n_cls > 0 && n_cls <= 64, "sfb_sync_head_bwd: B=%d T=%d n_cls=%d (n_cls <= 64)", B, T, n_cls);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SFB_CHECK_CUDA(cudaMemsetAsync(dx, 0, static_cast<size_t>(B) * T * kD * sizeof(float), st));     // only the token-0 rows receive gradient
    sync_head_bwd_rows_kernel<<<B, 256, 0, st>>>(x, T, ln_w, ln_b, eps, W, dlogits, n_cls, dx, scratch);
    SFB_CHECK_LAUNCH();
    sync_head_bwd_params_kernel<<<4, 256, 0, st>>>(scratch, dlogits, B, n_cls, dln_w, dln_b, dW, dbias);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
