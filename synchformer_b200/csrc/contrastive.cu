// N1: the tail of the stage-I contrastive step, AVCLIP.forward after the two towers
//   open_clip/model.py:536-545  encode_stream: (time-average pooling, motionformer.py:405-409) -> identity bridge -> F.normalize(dim=-1)
//   open_clip/model.py:492-497  optional all-gather of the normalised features (global negatives; done by torch.distributed between the
//                               kernels below)
//   open_clip/model.py:507-527  sim_v2a = v a_all^T / scale, sim_a2v = a v_all^T / scale, targets = eye(n, N) (the reference's targets: the
//                               positive of local row i is column i of the GATHERED features, whatever the rank), loss = (CE + CE) / 2
// and their backward.  Sizes are plumbing-scale (n = 128 local segments, N <= 1024 gathered, D = 768: 0.4 GFLOP against 52 TFLOP of encoder
// work per step), so these are deterministic CUDA-core kernels without atomics: a CTA per similarity row forward, a CTA per output row
// backward; nothing in this file is worth a tensor-core tile.
#include <math.h>

#include "common.cuh"

namespace sfb {
namespace contrastive {

constexpr int kThreads = 256;
constexpr int kMaxKeys = 8192;      // similarity row kept in shared memory (32 KB)

__device__ __forceinline__ float block_sum(float v, float *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) t += red[k];
    return t;
}
__device__ __forceinline__ float block_max(float v, float *red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = -INFINITY;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) t = fmaxf(t, red[k]);
    return t;
}

// out[i] = mean_t x[i, t, :]   (AveragePooling 'BS T D -> BS D')
__global__ void __launch_bounds__(kThreads) mean_tokens_kernel(const float *__restrict__ x, float *__restrict__ out, int T, int D) {
    const int64_t i = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += kThreads) {
        float s = 0.f;
        for (int t = 0; t < T; ++t) s += x[(i * T + t) * D + d];
        out[i * D + d] = s / T;
    }
}
// dx[i, t, :] = dout[i] / T
__global__ void __launch_bounds__(kThreads) mean_tokens_bwd_kernel(const float *__restrict__ dout, float *__restrict__ dx, int T, int D) {
    const int64_t i = blockIdx.x;
    for (int e = threadIdx.x; e < T * D; e += kThreads) dx[i * T * D + e] = dout[i * D + e % D] / T;
}

// xn = x / max(|x|, 1e-12)   (F.normalize), inv_norm kept for the backward
__global__ void __launch_bounds__(kThreads) l2_normalize_kernel(const float *__restrict__ x, float *__restrict__ xn, float *__restrict__ inv_norm, int D) {
    __shared__ float red[kThreads / 32];
    const int64_t i = blockIdx.x;
    float s = 0.f;
    for (int d = threadIdx.x; d < D; d += kThreads) s = fmaf(x[i * D + d], x[i * D + d], s);
    const float inv = 1.0f / fmaxf(sqrtf(block_sum(s, red)), 1e-12f);
    for (int d = threadIdx.x; d < D; d += kThreads) xn[i * D + d] = x[i * D + d] * inv;
    if (threadIdx.x == 0) inv_norm[i] = inv;
}
// dx = (dxn - xn (xn . dxn)) * inv_norm
__global__ void __launch_bounds__(kThreads) l2_normalize_bwd_kernel(const float *__restrict__ xn, const float *__restrict__ inv_norm,
                                                                    const float *__restrict__ dxn, float *__restrict__ dx, int D) {
    __shared__ float red[kThreads / 32];
    const int64_t i = blockIdx.x;
    float s = 0.f;
    for (int d = threadIdx.x; d < D; d += kThreads) s = fmaf(xn[i * D + d], dxn[i * D + d], s);
    const float dot = block_sum(s, red), inv = inv_norm[i];
    for (int d = threadIdx.x; d < D; d += kThreads) dx[i * D + d] = (dxn[i * D + d] - xn[i * D + d] * dot) * inv;
}

// CTA (i, dir): similarity row i of direction dir (0: v -> a_all, 1: a -> v_all), its softmax, G = d loss / d sim, the row's loss and
// its contribution to d loss / d scale.  q (n, D) local rows, keys (N, D) gathered rows.
__global__ void __launch_bounds__(kThreads) sim_rows_kernel(const float *__restrict__ vn, const float *__restrict__ an, const float *__restrict__ vn_all,
                                                            const float *__restrict__ an_all, int n, int N, int D, const float *__restrict__ scale,
                                                            float *__restrict__ G, float *__restrict__ row_loss, float *__restrict__ row_dscale) {
    extern __shared__ float sm[];            // q[D] | s[N]
    __shared__ float red[kThreads / 32];
    const int i = blockIdx.x, dir = blockIdx.y;
    const float inv_scale = 1.0f / scale[0];
    const float *q = (dir == 0 ? vn : an) + static_cast<int64_t>(i) * D;
    const float *keys = dir == 0 ? an_all : vn_all;
    float *qs = sm, *s = sm + D;
    for (int d = threadIdx.x; d < D; d += kThreads) qs[d] = q[d];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < N; j += kThreads / 32) {
        const float *k = keys + static_cast<int64_t>(j) * D;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc = fmaf(qs[d], k[d], acc);
        acc = warp_sum(acc);
        if (lane == 0) s[j] = acc * inv_scale;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < N; j += kThreads) mx = fmaxf(mx, s[j]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int j = threadIdx.x; j < N; j += kThreads) sum += expf(s[j] - mx);
    sum = block_sum(sum, red);
    const float inv_sum = 1.0f / sum, w = 0.5f / n;          // (CE_v2a + CE_a2v) / 2, each a mean over n rows
    float *g = G + (static_cast<int64_t>(dir) * n + i) * N;
    float ds = 0.f;
    for (int j = threadIdx.x; j < N; j += kThreads) {
        const float gij = (expf(s[j] - mx) * inv_sum - (j == i ? 1.0f : 0.0f)) * w;
        g[j] = gij;
        ds = fmaf(gij, s[j], ds);
    }
    ds = block_sum(ds, red);
    if (threadIdx.x == 0) {
        row_loss[dir * n + i] = (mx + logf(sum) - s[i]) * w;
        row_dscale[dir * n + i] = ds;
    }
}

// loss = sum of the 2n row losses (already weighted); dscale = -(sum_ij G_ij sim_ij) / scale     [sim = dot / scale]
__global__ void __launch_bounds__(kThreads) finalize_kernel(const float *__restrict__ row_loss, const float *__restrict__ row_dscale, int rows,
                                                            const float *__restrict__ scale, float *__restrict__ loss, float *__restrict__ dscale) {
    __shared__ float red[kThreads / 32];
    float a = 0.f, b = 0.f;
    for (int r = threadIdx.x; r < rows; r += kThreads) a += row_loss[r], b += row_dscale[r];
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0) loss[0] = a, dscale[0] = -b / scale[0];
}

// CTA (i, dir): gradient of the LOCAL (query) rows:  dq[i] = up * sum_j G[dir, i, j] keys[j] / scale
__global__ void __launch_bounds__(kThreads) grad_query_kernel(const float *__restrict__ G, const float *__restrict__ vn_all, const float *__restrict__ an_all,
                                                              int n, int N, int D, const float *__restrict__ scale, const float *__restrict__ upstream,
                                                              float *__restrict__ d_vn, float *__restrict__ d_an, const float *__restrict__ dscale,
                                                              float *__restrict__ dscale_out) {
    const int i = blockIdx.x, dir = blockIdx.y;
    const float coef = upstream[0] / scale[0];
    if (i == 0 && dir == 0 && threadIdx.x == 0) dscale_out[0] = dscale[0] * upstream[0];
    const float *keys = dir == 0 ? an_all : vn_all;
    const float *g = G + (static_cast<int64_t>(dir) * n + i) * N;
    float *dq = (dir == 0 ? d_vn : d_an) + static_cast<int64_t>(i) * D;
    for (int d = threadIdx.x; d < D; d += kThreads) {
        float acc = 0.f;
        for (int j = 0; j < N; ++j) acc = fmaf(g[j], keys[static_cast<int64_t>(j) * D + d], acc);
        dq[d] = acc * coef;
    }
}
// CTA (j, dir): gradient of the GATHERED (key) rows:  dk[j] = up * sum_i G[dir, i, j] q[i] / scale   (dir 0: keys are a_all, queries v)
// accumulate != 0 (no gathering, keys == queries): d_vn_all / d_an_all ARE the query gradients grad_query_kernel wrote; add to them
__global__ void __launch_bounds__(kThreads) grad_key_kernel(const float *__restrict__ G, const float *__restrict__ vn, const float *__restrict__ an,
                                                            int n, int N, int D, const float *__restrict__ scale, const float *__restrict__ upstream,
                                                            float *__restrict__ d_vn_all, float *__restrict__ d_an_all, int accumulate) {
    const int j = blockIdx.x, dir = blockIdx.y;
    const float coef = upstream[0] / scale[0];
    const float *q = dir == 0 ? vn : an;
    const float *g = G + static_cast<int64_t>(dir) * n * N + j;
    float *dk = (dir == 0 ? d_an_all : d_vn_all) + static_cast<int64_t>(j) * D;
    for (int d = threadIdx.x; d < D; d += kThreads) {
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc = fmaf(g[static_cast<int64_t>(i) * N], q[static_cast<int64_t>(i) * D + d], acc);
        dk[d] = accumulate ? dk[d] + acc * coef : acc * coef;
    }
}

}  // namespace contrastive
}  // namespace sfb

using namespace sfb;
using namespace sfb::contrastive;

extern "C" int sfb_mean_tokens(const float *x, float *out, int n, int T, int D, void *stream) {
    SFB_CHECK_ARG(x && out && n > 0 && T > 0 && D > 0, "sfb_mean_tokens: bad arguments");
    mean_tokens_kernel<<<n, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, T, D);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_mean_tokens_bwd(const float *dout, float *dx, int n, int T, int D, void *stream) {
    SFB_CHECK_ARG(dout && dx && n > 0 && T > 0 && D > 0, "sfb_mean_tokens_bwd: bad arguments");
    mean_tokens_bwd_kernel<<<n, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dout, dx, T, D);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_l2_normalize(const float *x, float *xn, float *inv_norm, int n, int D, void *stream) {
    SFB_CHECK_ARG(x && xn && inv_norm && n > 0 && D > 0, "sfb_l2_normalize: bad arguments");
    l2_normalize_kernel<<<n, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, xn, inv_norm, D);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_l2_normalize_bwd(const float *xn, const float *inv_norm, const float *dxn, float *dx, int n, int D, void *stream) {
    SFB_CHECK_ARG(xn && inv_norm && dxn && dx && n > 0 && D > 0, "sfb_l2_normalize_bwd: bad arguments");
    l2_normalize_bwd_kernel<<<n, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xn, inv_norm, dxn, dx, D);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_contrastive_loss(const float *vn, const float *an, const float *vn_all, const float *an_all, int n, int N, int D, const float *scale,
                                    float *loss, float *dscale, float *G, float *row_ws, void *stream) {
    SFB_CHECK_ARG(vn && an && vn_all && an_all && scale && loss && dscale && G && row_ws, "sfb_contrastive_loss: null pointer");
    SFB_CHECK_ARG(n > 0 && N >= n && N <= kMaxKeys && D > 0 && D <= 4096, "sfb_contrastive_loss: bad shape n=%d N=%d D=%d", n, N, D);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int smem = (D + N) * static_cast<int>(sizeof(float));
    static PerDeviceOnce once;
    if (once.first()) SFB_CHECK_CUDA(cudaFuncSetAttribute(sim_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (4096 + kMaxKeys) * 4));
    sim_rows_kernel<<<dim3(n, 2), kThreads, smem, st>>>(vn, an, vn_all, an_all, n, N, D, scale, G, row_ws, row_ws + 2 * n);
    SFB_CHECK_LAUNCH();
    finalize_kernel<<<1, kThreads, 0, st>>>(row_ws, row_ws + 2 * n, 2 * n, scale, loss, dscale);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_contrastive_loss_bwd(const float *vn, const float *an, const float *vn_all, const float *an_all, const float *G, int n, int N, int D,
                                        const float *scale, const float *upstream, const float *dscale, float *d_vn, float *d_an, float *d_vn_all,
                                        float *d_an_all, float *dscale_out, void *stream) {
    SFB_CHECK_ARG(vn && an && vn_all && an_all && G && scale && upstream && dscale && d_vn && d_an && dscale_out, "sfb_contrastive_loss_bwd: null pointer");
    SFB_CHECK_ARG(n > 0 && N >= n && D > 0, "sfb_contrastive_loss_bwd: bad shape");
    const bool accumulate = d_vn_all == nullptr && d_an_all == nullptr;
    SFB_CHECK_ARG(accumulate ? N == n : (d_vn_all != nullptr && d_an_all != nullptr),
                  "sfb_contrastive_loss_bwd: pass both key-gradient buffers, or neither when N == n (keys are the queries)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    grad_query_kernel<<<dim3(n, 2), kThreads, 0, st>>>(G, vn_all, an_all, n, N, D, scale, upstream, d_vn, d_an, dscale, dscale_out);
    SFB_CHECK_LAUNCH();
    grad_key_kernel<<<dim3(N, 2), kThreads, 0, st>>>(G, vn, an, n, N, D, scale, upstream, accumulate ? d_vn : d_vn_all, accumulate ? d_an : d_an_all,
                                                     accumulate ? 1 : 0);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
