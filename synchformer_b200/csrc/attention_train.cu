// N3: training-mode self-attention of the synchronisation transformer (modules/transformer.py:58-76) and its backward.
//
//   forward   P = softmax(scale Q K^T) (fp32),  Pd = dropout(P),  O = Pd V            -> O (bf16), lse (fp32, log2 units)
//   backward  D_i = dO_i . O_i,  dPd = dO V^T,  dS = P o (dPd o mask / (1 - p) - D),
//             dQ = scale dS K,  dK = scale dS^T Q,  dV = Pd^T dO                      -> dqkv (bf16), same fused layout as qkv
//
// Problems are (clip, head) pairs of T = 2 + 14 S <= 487 tokens (198 for 14 segments) with head_dim 96 (or 64): 256 problems of
// 198 x 198 x 96 at a batch of 32 clips, 0.1 % of the FLOPs of a training step (the frozen encoders' forward dominates), so these run
// in fp32 on the CUDA cores with K / V (or Q / dO) of one problem staged in shared memory, one warp per query row (dQ pass) or key
// row (dK / dV pass).  P is recomputed from Q, K and the saved log-sum-exp; dropout masks are regenerated from the Philox counter
// (philox.cuh), so nothing of size T x T is ever stored, and there are no atomics: results are deterministic.
// Layout: qkv / dqkv (B*T, 3*Dm) bf16 rows = [q | k | v], each head-major (h d); O / dO (B*T, Dm) bf16.
#include "common.cuh"
#include "philox.cuh"

namespace sfb {
namespace attn_train {

constexpr int kWarps = 8;
constexpr int kRowsPerCta = 32;

template <int HD>
struct Smem {
    static constexpr int HDP = HD + 2;     // padded row: stride of HDP / 2 = odd number of 32-bit words -> conflict-free row-per-lane reads
};

// stage `rows` rows of HD bf16 (global row stride ld elements) into shared memory rows of HDP elements
template <int HD>
__device__ __forceinline__ void stage_rows(const __nv_bfloat16 *__restrict__ g, int64_t ld, int rows, __nv_bfloat16 *s) {
    constexpr int HDP = Smem<HD>::HDP, CH = HD / 8;
    for (int c = threadIdx.x; c < rows * CH; c += blockDim.x) {
        const int r = c / CH, k = c % CH;
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(g + static_cast<int64_t>(r) * ld) + k);
        uint32_t *d = reinterpret_cast<uint32_t *>(s + r * HDP + k * 8);
        d[0] = u.x, d[1] = u.y, d[2] = u.z, d[3] = u.w;
    }
}

// dot product of an fp32 vector in shared memory (broadcast reads) with one padded bf16 row
template <int HD>
__device__ __forceinline__ float dot_row(const float *__restrict__ vec, const __nv_bfloat16 *__restrict__ row) {
    const uint32_t *rw = reinterpret_cast<const uint32_t *>(row);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int w = 0; w < HD / 2; ++w) {
        const float2 kk = unpack_bf16x2(rw[w]);
        const float2 qq = *reinterpret_cast<const float2 *>(vec + 2 * w);
        s0 = fmaf(qq.x, kk.x, s0);
        s1 = fmaf(qq.y, kk.y, s1);
    }
    return s0 + s1;
}

// ---------------------------------------------------------------------------------------------------------------------------
// forward: grid (ceil(T / 32), B * n_heads), 8 warps, warp per query row
// shared: K[T][HDP] V[T][HDP] bf16 | per warp: p[Tp] fp32, q[HD] fp32
// ---------------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kWarps * 32) attn_train_fwd_kernel(const __nv_bfloat16 *__restrict__ qkv, int64_t ld, int Dm, __nv_bfloat16 *__restrict__ out,
                                                                     int64_t ldo, float *__restrict__ lse, int T, int n_heads, float scale_log2,
                                                                     DropParams dp) {
    constexpr int HDP = Smem<HD>::HDP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tp = (T + 3) & ~3;
    __nv_bfloat16 *Ks = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *Vs = Ks + static_cast<size_t>(T) * HDP;
    float *wbase = reinterpret_cast<float *>(smem_raw + ((static_cast<size_t>(2) * T * HDP * 2 + 15) & ~static_cast<size_t>(15)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *prow = wbase + static_cast<size_t>(warp) * (Tp + HD);
    float *qrow = prow + Tp;

    const int bh = blockIdx.y, b = bh / n_heads, h = bh % n_heads;
    const __nv_bfloat16 *base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
    stage_rows<HD>(base + Dm, ld, T, Ks);
    stage_rows<HD>(base + 2 * Dm, ld, T, Vs);
    __syncthreads();

    const int i_end = min(T, (static_cast<int>(blockIdx.x) + 1) * kRowsPerCta);
    for (int i = blockIdx.x * kRowsPerCta + warp; i < i_end; i += kWarps) {
        const __nv_bfloat16 *qg = base + static_cast<int64_t>(i) * ld;
        for (int d = lane; d < HD; d += 32) qrow[d] = __bfloat162float(qg[d]) * scale_log2;
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j < T; j += 32) {
            const float s = dot_row<HD>(qrow, Ks + j * HDP);
            prow[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) {
            const float p = exp2f(prow[j] - mx);
            prow[j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        const uint64_t e0 = (static_cast<uint64_t>(bh) * T + i) * T;
        for (int j = lane; j < T; j += 32) prow[j] = prow[j] * inv * drop_scale(dp, e0 + j);
        if (lane == 0) lse[static_cast<int64_t>(bh) * T + i] = mx + log2f(sum);
        __syncwarp();
        float acc[HD / 32];
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) acc[c] = 0.f;
        for (int j = 0; j < T; ++j) {
            const float pj = prow[j];
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) acc[c] = fmaf(pj, __bfloat162float(Vs[j * HDP + lane + 32 * c]), acc[c]);
        }
        __nv_bfloat16 *og = out + (static_cast<int64_t>(b) * T + i) * ldo + h * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) og[lane + 32 * c] = __float2bfloat16(acc[c]);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// backward, dQ pass: grid (ceil(T / 32), B * n_heads), warp per query row; also writes delta_i = dO_i . O_i for the dK / dV pass
// shared: K V | per warp: ds[Tp], q[HD], do[HD] fp32
// ---------------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kWarps * 32) attn_train_bwd_dq_kernel(const __nv_bfloat16 *__restrict__ qkv, int64_t ld, int Dm,
                                                                        const __nv_bfloat16 *__restrict__ o, const __nv_bfloat16 *__restrict__ d_o, int64_t ldo,
                                                                        const float *__restrict__ lse, float *__restrict__ delta,
                                                                        __nv_bfloat16 *__restrict__ dqkv, int T, int n_heads, float scale, float scale_log2,
                                                                        DropParams dp) {
    constexpr int HDP = Smem<HD>::HDP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tp = (T + 3) & ~3;
    __nv_bfloat16 *Ks = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *Vs = Ks + static_cast<size_t>(T) * HDP;
    float *wbase = reinterpret_cast<float *>(smem_raw + ((static_cast<size_t>(2) * T * HDP * 2 + 15) & ~static_cast<size_t>(15)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ds = wbase + static_cast<size_t>(warp) * (Tp + 2 * HD);
    float *qrow = ds + Tp;
    float *dorow = qrow + HD;

    const int bh = blockIdx.y, b = bh / n_heads, h = bh % n_heads;
    const __nv_bfloat16 *base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
    stage_rows<HD>(base + Dm, ld, T, Ks);
    stage_rows<HD>(base + 2 * Dm, ld, T, Vs);
    __syncthreads();

    const int i_end = min(T, (static_cast<int>(blockIdx.x) + 1) * kRowsPerCta);
    for (int i = blockIdx.x * kRowsPerCta + warp; i < i_end; i += kWarps) {
        const int64_t row = static_cast<int64_t>(b) * T + i;
        const __nv_bfloat16 *qg = base + static_cast<int64_t>(i) * ld;
        const __nv_bfloat16 *og = o + row * ldo + h * HD, *dog = d_o + row * ldo + h * HD;
        float dl = 0.f;
        for (int d = lane; d < HD; d += 32) {
            const float g = __bfloat162float(dog[d]);
            qrow[d] = __bfloat162float(qg[d]) * scale_log2;
            dorow[d] = g;
            dl = fmaf(g, __bfloat162float(og[d]), dl);
        }
        dl = warp_sum(dl);
        __syncwarp();
        const float L = lse[static_cast<int64_t>(bh) * T + i];
        const uint64_t e0 = (static_cast<uint64_t>(bh) * T + i) * T;
        for (int j = lane; j < T; j += 32) {
            const float p = exp2f(dot_row<HD>(qrow, Ks + j * HDP) - L);
            const float dpd = dot_row<HD>(dorow, Vs + j * HDP);
            ds[j] = p * (dpd * drop_scale(dp, e0 + j) - dl);
        }
        if (lane == 0) delta[static_cast<int64_t>(bh) * T + i] = dl;
        __syncwarp();
        float acc[HD / 32];
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) acc[c] = 0.f;
        for (int j = 0; j < T; ++j) {
            const float sj = ds[j];
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) acc[c] = fmaf(sj, __bfloat162float(Ks[j * HDP + lane + 32 * c]), acc[c]);
        }
        __nv_bfloat16 *dq = dqkv + row * ld + h * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) dq[lane + 32 * c] = __float2bfloat16(acc[c] * scale);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// backward, dK / dV pass: grid (ceil(T / 32), B * n_heads), warp per key row j, all query rows staged
// shared: Q[T][HDP] dO[T][HDP] bf16 | lse[Tp] delta[Tp] fp32 | per warp: ds[Tp], pd[Tp], k[HD], v[HD] fp32
// ---------------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kWarps * 32) attn_train_bwd_dkv_kernel(const __nv_bfloat16 *__restrict__ qkv, int64_t ld, int Dm,
                                                                         const __nv_bfloat16 *__restrict__ d_o, int64_t ldo, const float *__restrict__ lse,
                                                                         const float *__restrict__ delta, __nv_bfloat16 *__restrict__ dqkv, int T,
                                                                         int n_heads, float scale, float scale_log2, DropParams dp) {
    constexpr int HDP = Smem<HD>::HDP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tp = (T + 3) & ~3;
    __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *dOs = Qs + static_cast<size_t>(T) * HDP;
    float *fbase = reinterpret_cast<float *>(smem_raw + ((static_cast<size_t>(2) * T * HDP * 2 + 15) & ~static_cast<size_t>(15)));
    float *lse_s = fbase, *delta_s = fbase + Tp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ds = fbase + 2 * Tp + static_cast<size_t>(warp) * (2 * Tp + 2 * HD);
    float *pd = ds + Tp;
    float *krow = pd + Tp;
    float *vrow = krow + HD;

    const int bh = blockIdx.y, b = bh / n_heads, h = bh % n_heads;
    const __nv_bfloat16 *base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
    stage_rows<HD>(base, ld, T, Qs);
    stage_rows<HD>(d_o + static_cast<int64_t>(b) * T * ldo + h * HD, ldo, T, dOs);
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        lse_s[i] = lse[static_cast<int64_t>(bh) * T + i];
        delta_s[i] = delta[static_cast<int64_t>(bh) * T + i];
    }
    __syncthreads();

    const int j_end = min(T, (static_cast<int>(blockIdx.x) + 1) * kRowsPerCta);
    for (int j = blockIdx.x * kRowsPerCta + warp; j < j_end; j += kWarps) {
        const __nv_bfloat16 *kg = base + static_cast<int64_t>(j) * ld + Dm, *vg = kg + Dm;
        for (int d = lane; d < HD; d += 32) {
            krow[d] = __bfloat162float(kg[d]) * scale_log2;
            vrow[d] = __bfloat162float(vg[d]);
        }
        __syncwarp();
        for (int i = lane; i < T; i += 32) {
            const float p = exp2f(dot_row<HD>(krow, Qs + i * HDP) - lse_s[i]);
            const float dpd = dot_row<HD>(vrow, dOs + i * HDP);
            const float m = drop_scale(dp, (static_cast<uint64_t>(bh) * T + i) * T + j);
            pd[i] = p * m;
            ds[i] = p * (dpd * m - delta_s[i]);
        }
        __syncwarp();
        float ak[HD / 32], av[HD / 32];
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) ak[c] = 0.f, av[c] = 0.f;
        for (int i = 0; i < T; ++i) {
            const float si = ds[i], pi = pd[i];
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
                ak[c] = fmaf(si, __bfloat162float(Qs[i * HDP + lane + 32 * c]), ak[c]);
                av[c] = fmaf(pi, __bfloat162float(dOs[i * HDP + lane + 32 * c]), av[c]);
            }
        }
        __nv_bfloat16 *dk = dqkv + (static_cast<int64_t>(b) * T + j) * ld + Dm + h * HD, *dv = dk + Dm;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
            dk[lane + 32 * c] = __float2bfloat16(ak[c] * scale);
            dv[lane + 32 * c] = __float2bfloat16(av[c]);
        }
        __syncwarp();
    }
}

static size_t smem_fwd(int T, int HD) { return ((static_cast<size_t>(2) * T * (HD + 2) * 2 + 15) & ~static_cast<size_t>(15)) + sizeof(float) * kWarps * (((T + 3) & ~3) + HD); }
static size_t smem_dq(int T, int HD) { return ((static_cast<size_t>(2) * T * (HD + 2) * 2 + 15) & ~static_cast<size_t>(15)) + sizeof(float) * kWarps * (((T + 3) & ~3) + 2 * HD); }
static size_t smem_dkv(int T, int HD) {
    const size_t Tp = (T + 3) & ~3;
    return ((static_cast<size_t>(2) * T * (HD + 2) * 2 + 15) & ~static_cast<size_t>(15)) + sizeof(float) * (2 * Tp + kWarps * (2 * Tp + 2 * HD));
}
constexpr size_t kMaxSmem = 227 * 1024;

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    SFB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    return SFB_OK;
}

static int check_common(const char *what, const void *qkv, int B, int T, int n_heads, int head_dim, float p) {
    SFB_CHECK_ARG(qkv != nullptr && B > 0 && T > 0 && n_heads > 0, "%s: bad arguments (B=%d T=%d heads=%d)", what, B, T, n_heads);
    SFB_CHECK_ARG(head_dim == 96 || head_dim == 64, "%s: head_dim %d not in {64, 96}", what, head_dim);
    SFB_CHECK_ARG(p >= 0.f && p < 1.f, "%s: p=%f outside [0, 1)", what, p);
    SFB_CHECK_ARG(static_cast<int64_t>(B) * n_heads <= 65535, "%s: B * n_heads = %lld exceeds 65535 problems per launch", what,
                  static_cast<long long>(B) * n_heads);
    if (smem_dkv(T, head_dim) > kMaxSmem) {
        set_error("%s: T=%d needs %zu bytes of shared memory (> 227 KB)", what, T, smem_dkv(T, head_dim));
        return SFB_E_UNSUPPORTED;
    }
    return SFB_OK;
}

}  // namespace attn_train
}  // namespace sfb

extern "C" int sfb_attention_train_fwd(const void *qkv, void *out, float *lse, int B, int T, int n_heads, int head_dim, float scale, float p_drop,
                                       uint64_t seed, uint32_t site, void *stream) {
    using namespace sfb;
    using namespace sfb::attn_train;
    int rc = check_common("sfb_attention_train_fwd", qkv, B, T, n_heads, head_dim, p_drop);
    if (rc != SFB_OK) return rc;
    SFB_CHECK_ARG(out && lse && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "sfb_attention_train_fwd: null / unaligned pointer");
    const int Dm = n_heads * head_dim;
    const DropParams dp = make_drop_params(p_drop, seed, site);
    const dim3 grid((T + kRowsPerCta - 1) / kRowsPerCta, B * n_heads);
    const float sl2 = scale * 1.4426950408889634f;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = smem_fwd(T, head_dim);
    const __nv_bfloat16 *q = reinterpret_cast<const __nv_bfloat16 *>(qkv);
    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(out);
    if (head_dim == 96) {
        if ((rc = set_smem(attn_train_fwd_kernel<96>, smem)) != SFB_OK) return rc;
        attn_train_fwd_kernel<96><<<grid, kWarps * 32, smem, st>>>(q, 3 * Dm, Dm, o, Dm, lse, T, n_heads, sl2, dp);
    } else {
        if ((rc = set_smem(attn_train_fwd_kernel<64>, smem)) != SFB_OK) return rc;
        attn_train_fwd_kernel<64><<<grid, kWarps * 32, smem, st>>>(q, 3 * Dm, Dm, o, Dm, lse, T, n_heads, sl2, dp);
    }
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_attention_train_bwd(const void *qkv, const void *out, const void *d_out, const float *lse, float *delta, void *dqkv, int B, int T,
                                       int n_heads, int head_dim, float scale, float p_drop, uint64_t seed, uint32_t site, void *stream) {
    using namespace sfb;
    using namespace sfb::attn_train;
    int rc = check_common("sfb_attention_train_bwd", qkv, B, T, n_heads, head_dim, p_drop);
    if (rc != SFB_OK) return rc;
    SFB_CHECK_ARG(out && d_out && lse && delta && dqkv && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0,
                  "sfb_attention_train_bwd: null / unaligned pointer");
    const int Dm = n_heads * head_dim;
    const DropParams dp = make_drop_params(p_drop, seed, site);
    const dim3 grid((T + kRowsPerCta - 1) / kRowsPerCta, B * n_heads);
    const float sl2 = scale * 1.4426950408889634f;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t s1 = smem_dq(T, head_dim), s2 = smem_dkv(T, head_dim);
    const __nv_bfloat16 *q = reinterpret_cast<const __nv_bfloat16 *>(qkv), *o = reinterpret_cast<const __nv_bfloat16 *>(out),
                        *g = reinterpret_cast<const __nv_bfloat16 *>(d_out);
    __nv_bfloat16 *dq = reinterpret_cast<__nv_bfloat16 *>(dqkv);
    if (head_dim == 96) {
        if ((rc = set_smem(attn_train_bwd_dq_kernel<96>, s1)) != SFB_OK) return rc;
        if ((rc = set_smem(attn_train_bwd_dkv_kernel<96>, s2)) != SFB_OK) return rc;
        attn_train_bwd_dq_kernel<96><<<grid, kWarps * 32, s1, st>>>(q, 3 * Dm, Dm, o, g, Dm, lse, delta, dq, T, n_heads, scale, sl2, dp);
        SFB_CHECK_LAUNCH();
        attn_train_bwd_dkv_kernel<96><<<grid, kWarps * 32, s2, st>>>(q, 3 * Dm, Dm, g, Dm, lse, delta, dq, T, n_heads, scale, sl2, dp);
    } else {
        if ((rc = set_smem(attn_train_bwd_dq_kernel<64>, s1)) != SFB_OK) return rc;
        if ((rc = set_smem(attn_train_bwd_dkv_kernel<64>, s2)) != SFB_OK) return rc;
        attn_train_bwd_dq_kernel<64><<<grid, kWarps * 32, s1, st>>>(q, 3 * Dm, Dm, o, g, Dm, lse, delta, dq, T, n_heads, scale, sl2, dp);
        SFB_CHECK_LAUNCH();
        attn_train_bwd_dkv_kernel<64><<<grid, kWarps * 32, s2, st>>>(q, 3 * Dm, Dm, g, Dm, lse, delta, dq, T, n_heads, scale, sl2, dp);
    }
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
