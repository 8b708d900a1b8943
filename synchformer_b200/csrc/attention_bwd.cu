// N1: backward of the encoders' attention on the strided views of sfb_attention (no rearrange copies in the backward
// either).  Two entry points:
//
//   sfb_attention_bwd               problems (outer, inner, head) of Lq x (Lk [+ 1 prefix key]) with Lk + 1 <= ~480:
//                                   Motionformer time attention (8 x 9) and space attention (196 x 197) with the CLS key/value as the
//                                   prefix row (vit_helper.py:126-146), the CLS aggregators (1 x 197, 1 x 13; motionformer.py:301-334)
//   sfb_attention_bwd_global_query  one query per (outer, head) over ALL Lk rows of its outer index (the Motionformer CLS query,
//                                   1 x 1569, vit_helper.py:124); its dK / dV contributions are ADDED to what sfb_attention_bwd wrote
//
// Math per problem (P recomputed, nothing of size Lq x Lk stored):
//   S = scale Q K^T, P = softmax(S), O = P V;   D_i = dO_i . O_i,  dS = P o (dO V^T - D),  dQ = scale dS K,  dK = scale dS^T Q,  dV = P^T dO
// Pass 1 (warp per query row, K / V staged in shared memory) writes dQ and the row statistics (log-sum-exp, D); pass 2 (warp per key
// row, Q / dO staged) writes dK / dV.  The prefix key is shared by all inner problems of an outer index, so its dK / dV go to a
// per-problem fp32 buffer that the caller reduces (sfb_colsum) - no atomics, deterministic.  fp32 math on the CUDA cores: these
// problems are 3 % of the encoder FLOPs in the forward; a tensor-core version is a later optimisation, correctness comes first.
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace sfb {
namespace attn_bwd {

constexpr int kWarps = 8;
constexpr int kRowsPerCta = 32;

struct Desc {
    const __nv_bfloat16 *q, *k, *v, *k_prefix, *v_prefix, *o, *d_o;
    __nv_bfloat16 *dq, *dk, *dv;
    float *dprefix;   // [(inner * n_outer + outer) * n_heads + head][2][HD] fp32, or nullptr when there is no prefix
    float *stats;     // [problem][Lq][2] = { lse (log2 units), D }
    int64_t q_outer, q_inner, q_row, kv_outer, kv_inner, kv_row, o_outer, o_inner, o_row, prefix_outer;
    int n_outer, n_inner, n_heads, Lq, Lk;
    float scale, scale_log2;
};

template <int HD>
__device__ __forceinline__ void stage_rows(const __nv_bfloat16 *__restrict__ g, int64_t row_stride, int rows, __nv_bfloat16 *s) {
    constexpr int HDP = HD + 2, CH = HD / 8;
    for (int c = threadIdx.x; c < rows * CH; c += blockDim.x) {
        const int r = c / CH, kk = c % CH;
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(g + static_cast<int64_t>(r) * row_stride) + kk);
        uint32_t *d = reinterpret_cast<uint32_t *>(s + r * HDP + kk * 8);
        d[0] = u.x, d[1] = u.y, d[2] = u.z, d[3] = u.w;
    }
}

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

__device__ __forceinline__ void decode_problem(const Desc &d, int p, int &o, int &i, int &h) {
    h = p % d.n_heads;
    const int oi = p / d.n_heads;
    i = oi % d.n_inner;
    o = oi / d.n_inner;
}

// pass 1: grid (problems, ceil(Lq / 32)); shared: K[Lt][HDP] V[Lt][HDP] | per warp: p[Ltp] q[HD] do[HD]
template <int HD>
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_dq_kernel(const Desc d) {
    constexpr int HDP = HD + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lt = d.Lk + (d.k_prefix != nullptr ? 1 : 0);
    const int Ltp = (Lt + 3) & ~3;
    __nv_bfloat16 *Ks = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *Vs = Ks + static_cast<size_t>(Lt) * HDP;
    float *wbase = reinterpret_cast<float *>(smem_raw + ((static_cast<size_t>(2) * Lt * HDP * 2 + 15) & ~static_cast<size_t>(15)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ds = wbase + static_cast<size_t>(warp) * (Ltp + 2 * HD);
    float *qrow = ds + Ltp;
    float *dorow = qrow + HD;

    int o, i, h;
    decode_problem(d, blockIdx.x, o, i, h);
    const int64_t kv_off = o * d.kv_outer + i * d.kv_inner + h * HD;
    stage_rows<HD>(d.k + kv_off, d.kv_row, d.Lk, Ks);
    stage_rows<HD>(d.v + kv_off, d.kv_row, d.Lk, Vs);
    if (d.k_prefix != nullptr) {     // the prefix row is stored LAST (key order is irrelevant to softmax)
        stage_rows<HD>(d.k_prefix + o * d.prefix_outer + h * HD, 0, 1, Ks + d.Lk * HDP);
        stage_rows<HD>(d.v_prefix + o * d.prefix_outer + h * HD, 0, 1, Vs + d.Lk * HDP);
    }
    __syncthreads();

    const int r_end = min(d.Lq, (static_cast<int>(blockIdx.y) + 1) * kRowsPerCta);
    for (int r = blockIdx.y * kRowsPerCta + warp; r < r_end; r += kWarps) {
        const __nv_bfloat16 *qg = d.q + o * d.q_outer + i * d.q_inner + h * HD + r * d.q_row;
        const int64_t o_off = o * d.o_outer + i * d.o_inner + h * HD + r * d.o_row;
        float dl = 0.f;
        for (int c = lane; c < HD; c += 32) {
            const float g = __bfloat162float(d.d_o[o_off + c]);
            qrow[c] = __bfloat162float(qg[c]) * d.scale_log2;
            dorow[c] = g;
            dl = fmaf(g, __bfloat162float(d.o[o_off + c]), dl);
        }
        dl = warp_sum(dl);
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j < Lt; j += 32) {
            const float s = dot_row<HD>(qrow, Ks + j * HDP);
            ds[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < Lt; j += 32) sum += exp2f(ds[j] - mx);
        sum = warp_sum(sum);
        const float L = mx + log2f(sum);
        for (int j = lane; j < Lt; j += 32) {
            const float p = exp2f(ds[j] - L);
            ds[j] = p * (dot_row<HD>(dorow, Vs + j * HDP) - dl);
        }
        if (lane == 0) {
            float *st = d.stats + (static_cast<int64_t>(blockIdx.x) * d.Lq + r) * 2;
            st[0] = L, st[1] = dl;
        }
        __syncwarp();
        float acc[HD / 32];
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) acc[c] = 0.f;
        for (int j = 0; j < Lt; ++j) {
            const float sj = ds[j];
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) acc[c] = fmaf(sj, __bfloat162float(Ks[j * HDP + lane + 32 * c]), acc[c]);
        }
        __nv_bfloat16 *dq = d.dq + o * d.q_outer + i * d.q_inner + h * HD + r * d.q_row;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) dq[lane + 32 * c] = __float2bfloat16(acc[c] * d.scale);
        __syncwarp();
    }
}

// pass 2: grid (problems, ceil((Lk + prefix) / 32)); shared: Q[Lq][HDP] dO[Lq][HDP] | lse[Lqp] D[Lqp] | per warp: ds[Lqp] p[Lqp] k[HD] v[HD]
template <int HD>
__global__ void __launch_bounds__(kWarps * 32) attn_bwd_dkv_kernel(const Desc d) {
    constexpr int HDP = HD + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lt = d.Lk + (d.k_prefix != nullptr ? 1 : 0);
    const int Lqp = (d.Lq + 3) & ~3;
    __nv_bfloat16 *Qs = reinterpret_cast<__nv_bfloat16 *>(smem_raw);
    __nv_bfloat16 *dOs = Qs + static_cast<size_t>(d.Lq) * HDP;
    float *fbase = reinterpret_cast<float *>(smem_raw + ((static_cast<size_t>(2) * d.Lq * HDP * 2 + 15) & ~static_cast<size_t>(15)));
    float *lse_s = fbase, *delta_s = fbase + Lqp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ds = fbase + 2 * Lqp + static_cast<size_t>(warp) * (2 * Lqp + 2 * HD);
    float *pp = ds + Lqp;
    float *krow = pp + Lqp;
    float *vrow = krow + HD;

    int o, i, h;
    decode_problem(d, blockIdx.x, o, i, h);
    stage_rows<HD>(d.q + o * d.q_outer + i * d.q_inner + h * HD, d.q_row, d.Lq, Qs);
    stage_rows<HD>(d.d_o + o * d.o_outer + i * d.o_inner + h * HD, d.o_row, d.Lq, dOs);
    for (int r = threadIdx.x; r < d.Lq; r += blockDim.x) {
        const float *st = d.stats + (static_cast<int64_t>(blockIdx.x) * d.Lq + r) * 2;
        lse_s[r] = st[0], delta_s[r] = st[1];
    }
    __syncthreads();

    const int j_end = min(Lt, (static_cast<int>(blockIdx.y) + 1) * kRowsPerCta);
    for (int j = blockIdx.y * kRowsPerCta + warp; j < j_end; j += kWarps) {
        const bool is_prefix = j >= d.Lk;
        const int64_t kv_off = o * d.kv_outer + i * d.kv_inner + h * HD + static_cast<int64_t>(j) * d.kv_row;
        const __nv_bfloat16 *kg = is_prefix ? d.k_prefix + o * d.prefix_outer + h * HD : d.k + kv_off;
        const __nv_bfloat16 *vg = is_prefix ? d.v_prefix + o * d.prefix_outer + h * HD : d.v + kv_off;
        for (int c = lane; c < HD; c += 32) {
            krow[c] = __bfloat162float(kg[c]) * d.scale_log2;
            vrow[c] = __bfloat162float(vg[c]);
        }
        __syncwarp();
        for (int r = lane; r < d.Lq; r += 32) {
            const float p = exp2f(dot_row<HD>(krow, Qs + r * HDP) - lse_s[r]);
            pp[r] = p;
            ds[r] = p * (dot_row<HD>(vrow, dOs + r * HDP) - delta_s[r]);
        }
        __syncwarp();
        float ak[HD / 32], av[HD / 32];
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) ak[c] = 0.f, av[c] = 0.f;
        for (int r = 0; r < d.Lq; ++r) {
            const float sr = ds[r], pr = pp[r];
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
                ak[c] = fmaf(sr, __bfloat162float(Qs[r * HDP + lane + 32 * c]), ak[c]);
                av[c] = fmaf(pr, __bfloat162float(dOs[r * HDP + lane + 32 * c]), av[c]);
            }
        }
        if (is_prefix) {
            float *dst = d.dprefix + ((static_cast<int64_t>(i) * d.n_outer + o) * d.n_heads + h) * 2 * HD;
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
                dst[lane + 32 * c] = ak[c] * d.scale;
                dst[HD + lane + 32 * c] = av[c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
                d.dk[kv_off + lane + 32 * c] = __float2bfloat16(ak[c] * d.scale);
                d.dv[kv_off + lane + 32 * c] = __float2bfloat16(av[c]);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Tensor-core variant for the large problems (head_dim 64, 64 <= Lq, Lq and Lk + prefix <= 256: the Motionformer space attention,
// 196 x 197): mma.sync m16n8k16 with the fragment idioms of attn_mma_kernel (attention.cu).  One CTA (8 warps) per problem; Q, K, V, dO
// are staged once in shared memory (144-byte row pitch).
//   phase 0  D_i = dO_i . O_i                                                     (warp per row)
//   phase 1  warp = 16 query rows: pass A row log-sum-exp; pass B per 64-key chunk  S -> P,  dP = dO V^T,  dS = P (dP - D),
//            dQ += dS K  (dS re-used as the A operand straight from its accumulator layout, K^T fragments via ldmatrix.trans)
//   phase 2  warp = 16 key rows: per 64-query chunk  S^T = K Q^T -> P^T (column statistics from shared memory),  dP^T = V dO^T,
//            dV += P^T dO,  dK += dS^T Q
// P and dS enter the second-stage MMAs as bf16 (as in the forward); accumulation is fp32.  Same outputs as the CUDA-core pair above
// (which stays for the small problems and as the cross-check: desc->impl == 1 forces it).
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kMmaWarps = 8;
constexpr int kPitch = 64 * 2 + 16;      // bytes per staged row: ldmatrix rows land on distinct bank groups

// A-operand fragments of 16 consecutive rows starting at row r0 of a staged row-major [row][64] matrix
__device__ __forceinline__ void load_a_frags(uint32_t base, int r0, int lane, uint32_t (&a)[4][4]) {
    const uint32_t addr = base + (r0 + (lane & 15)) * kPitch + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(addr + ks * 32, a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}
// acc[2 kb], acc[2 kb + 1] (16 x 16 columns = staged rows c0 + 16 kb ..) += A (16 x 64) * M[rows]^T  for kb < nkb
__device__ __forceinline__ void mma_a_rowsT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t base, int c0, int nkb, int lane) {
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
        if (kb < nkb) {
            const uint32_t addr = base + (c0 + kb * 16 + (lane & 7) + ((lane >> 4) << 3)) * kPitch + ((lane >> 3) & 1) * 16;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(addr + ks * 32, b0, b1, b2, b3);
                mma_bf16_16816(acc[2 * kb], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b0, b1);
                mma_bf16_16816(acc[2 * kb + 1], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b2, b3);
            }
        }
    }
}
// out (16 x 64) += P (16 x 64 columns = staged rows c0 ..; taken from its accumulator layout, rounded to bf16) * M[rows c0 ..] (64 x 64)
__device__ __forceinline__ void mma_p_rows(float (&out)[8][4], const float (&p)[8][4], uint32_t base, int c0, int nkb, int lane) {
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
        if (kb < nkb) {
            const uint32_t a0 = pack_bf16x2(p[2 * kb][0], p[2 * kb][1]), a1 = pack_bf16x2(p[2 * kb][2], p[2 * kb][3]);
            const uint32_t a2 = pack_bf16x2(p[2 * kb + 1][0], p[2 * kb + 1][1]), a3 = pack_bf16x2(p[2 * kb + 1][2], p[2 * kb + 1][3]);
            const uint32_t addr = base + (c0 + kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 16;
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(addr + n2 * 32, b0, b1, b2, b3);
                mma_bf16_16816(out[2 * n2], a0, a1, a2, a3, b0, b1);
                mma_bf16_16816(out[2 * n2 + 1], a0, a1, a2, a3, b2, b3);
            }
        }
    }
}

__global__ void __launch_bounds__(kMmaWarps * 32) attn_bwd_mma_kernel(const Desc d, int Lq_pad, int Lk_pad) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int has_prefix = d.k_prefix != nullptr ? 1 : 0;
    const int Lkp = d.Lk + has_prefix;
    const uint32_t sQ = s0, sdO = sQ + Lq_pad * kPitch, sK = sdO + Lq_pad * kPitch, sV = sK + Lk_pad * kPitch;
    float *lse_s = reinterpret_cast<float *>(smem + (2 * Lq_pad + 2 * Lk_pad) * kPitch);
    float *delta_s = lse_s + Lq_pad;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    int o, i, h;
    decode_problem(d, blockIdx.x, o, i, h);

    // ---- stage Q, dO, K, V (prefix key / value first, as in the forward); padding rows are zero so they add nothing to any sum ----
    {
        const __nv_bfloat16 *qg = d.q + o * d.q_outer + i * d.q_inner + h * 64, *dog = d.d_o + o * d.o_outer + i * d.o_inner + h * 64;
        for (int c = tid; c < Lq_pad * 8; c += nthr) {
            const int r = c >> 3, cc = c & 7;
            if (r < d.Lq) {
                cp_async16(sQ + r * kPitch + cc * 16, qg + static_cast<int64_t>(r) * d.q_row + cc * 8);
                cp_async16(sdO + r * kPitch + cc * 16, dog + static_cast<int64_t>(r) * d.o_row + cc * 8);
            } else {
                *reinterpret_cast<uint4 *>(smem + r * kPitch + cc * 16) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(smem + (Lq_pad + r) * kPitch + cc * 16) = make_uint4(0, 0, 0, 0);
            }
        }
        const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * 64, pre_base = o * d.prefix_outer + h * 64;
        for (int c = tid; c < Lk_pad * 8; c += nthr) {
            const int r = c >> 3, cc = c & 7;
            if (r < Lkp) {
                const bool pre = has_prefix && r == 0;
                const int64_t off = (pre ? pre_base : kv_base + static_cast<int64_t>(r - has_prefix) * d.kv_row) + cc * 8;
                cp_async16(sK + r * kPitch + cc * 16, (pre ? d.k_prefix : d.k) + off);
                cp_async16(sV + r * kPitch + cc * 16, (pre ? d.v_prefix : d.v) + off);
            } else {
                *reinterpret_cast<uint4 *>(smem + (2 * Lq_pad + r) * kPitch + cc * 16) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(smem + (2 * Lq_pad + Lk_pad + r) * kPitch + cc * 16) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    // ---- phase 0: D_r = dO_r . O_r (O read from global, one warp per row, two elements per lane) ----
    for (int r = warp; r < Lq_pad; r += kMmaWarps) {
        float dl = 0.f;
        if (r < d.Lq) {
            const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(smem + (Lq_pad + r) * kPitch + lane * 4));
            const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(d.o + o * d.o_outer + i * d.o_inner + h * 64 + static_cast<int64_t>(r) * d.o_row + lane * 2));
            dl = a.x * b.x + a.y * b.y;
        }
        dl = warp_sum(dl);
        if (lane == 0) delta_s[r] = dl;
    }
    __syncthreads();

    const float sl2 = d.scale_log2;
    // ---- phase 1: dQ and the row statistics ----
    for (int q0 = warp * 16; q0 < Lq_pad; q0 += kMmaWarps * 16) {
        uint32_t qa[4][4], doa[4][4];
        load_a_frags(sQ, q0, lane, qa);
        load_a_frags(sdO, q0, lane, doa);
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        for (int kc = 0; kc < Lk_pad; kc += 64) {                 // pass A: log-sum-exp of rows q0 + g, q0 + g + 8
            const int nkb = min(4, (Lk_pad - kc) >> 4);
            float s[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            mma_a_rowsT(s, qa, sK, kc, nkb, lane);
            float mx0 = m0, mx1 = m1;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int key = kc + n * 8 + t4 * 2;
                const bool live = n < 2 * nkb;
                s[n][0] = (live && key < Lkp) ? s[n][0] * sl2 : -INFINITY;
                s[n][1] = (live && key + 1 < Lkp) ? s[n][1] * sl2 : -INFINITY;
                s[n][2] = (live && key < Lkp) ? s[n][2] * sl2 : -INFINITY;
                s[n][3] = (live && key + 1 < Lkp) ? s[n][3] * sl2 : -INFINITY;
                mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
                mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                ps0 += exp2f(s[n][0] - mx0) + exp2f(s[n][1] - mx0);
                ps1 += exp2f(s[n][2] - mx1) + exp2f(s[n][3] - mx1);
            }
            l0 = l0 * exp2f(m0 - mx0) + ps0;
            l1 = l1 * exp2f(m1 - mx1) + ps1;
            m0 = mx0, m1 = mx1;
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        // padded query rows get +inf so that phase 2 sees P = 0 for them
        const float L0 = q0 + g < d.Lq ? m0 + log2f(l0) : INFINITY, L1 = q0 + g + 8 < d.Lq ? m1 + log2f(l1) : INFINITY;
        if (t4 == 0) lse_s[q0 + g] = L0, lse_s[q0 + g + 8] = L1;
        const float D0 = delta_s[q0 + g], D1 = delta_s[q0 + g + 8];
        float dq[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
        for (int kc = 0; kc < Lk_pad; kc += 64) {                 // pass B
            const int nkb = min(4, (Lk_pad - kc) >> 4);
            float s[8][4], dp[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
            mma_a_rowsT(s, qa, sK, kc, nkb, lane);
            mma_a_rowsT(dp, doa, sV, kc, nkb, lane);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int key = kc + n * 8 + t4 * 2;
                const bool live = n < 2 * nkb;
                const bool v0 = live && key < Lkp, v1 = live && key + 1 < Lkp;
                s[n][0] = v0 ? exp2f(s[n][0] * sl2 - L0) * (dp[n][0] - D0) : 0.f;
                s[n][1] = v1 ? exp2f(s[n][1] * sl2 - L0) * (dp[n][1] - D0) : 0.f;
                s[n][2] = v0 ? exp2f(s[n][2] * sl2 - L1) * (dp[n][2] - D1) : 0.f;
                s[n][3] = v1 ? exp2f(s[n][3] * sl2 - L1) * (dp[n][3] - D1) : 0.f;
            }
            mma_p_rows(dq, s, sK, kc, nkb, lane);                 // dQ += dS K
        }
        __nv_bfloat16 *dqg = d.dq + o * d.q_outer + i * d.q_inner + h * 64;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            if (q0 + g < d.Lq)
                *reinterpret_cast<uint32_t *>(dqg + static_cast<int64_t>(q0 + g) * d.q_row + n * 8 + t4 * 2) = pack_bf16x2(dq[n][0] * d.scale, dq[n][1] * d.scale);
            if (q0 + g + 8 < d.Lq)
                *reinterpret_cast<uint32_t *>(dqg + static_cast<int64_t>(q0 + g + 8) * d.q_row + n * 8 + t4 * 2) = pack_bf16x2(dq[n][2] * d.scale, dq[n][3] * d.scale);
        }
    }
    __syncthreads();                                              // lse_s is complete

    // ---- phase 2: dK, dV for 16 staged key rows per warp (staged row 0 is the prefix key when there is one) ----
    for (int k0 = warp * 16; k0 < Lk_pad; k0 += kMmaWarps * 16) {
        uint32_t ka[4][4], va[4][4];
        load_a_frags(sK, k0, lane, ka);
        load_a_frags(sV, k0, lane, va);
        const bool kv0 = k0 + g < Lkp, kv1 = k0 + g + 8 < Lkp;
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
        for (int qc = 0; qc < Lq_pad; qc += 64) {
            const int nqb = min(4, (Lq_pad - qc) >> 4);
            float st[8][4], dpt[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) st[n][0] = st[n][1] = st[n][2] = st[n][3] = dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
            mma_a_rowsT(st, ka, sQ, qc, nqb, lane);               // S^T = K Q^T
            mma_a_rowsT(dpt, va, sdO, qc, nqb, lane);             // dP^T = V dO^T
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const bool live = n < 2 * nqb;
                const int qi = qc + n * 8 + t4 * 2;
                const float La = live ? lse_s[qi] : INFINITY, Lb = live ? lse_s[qi + 1] : INFINITY;
                const float Da = live ? delta_s[qi] : 0.f, Db = live ? delta_s[qi + 1] : 0.f;
                const float p0 = kv0 ? exp2f(st[n][0] * sl2 - La) : 0.f, p1 = kv0 ? exp2f(st[n][1] * sl2 - Lb) : 0.f;
                const float p2 = kv1 ? exp2f(st[n][2] * sl2 - La) : 0.f, p3 = kv1 ? exp2f(st[n][3] * sl2 - Lb) : 0.f;
                st[n][0] = p0, st[n][1] = p1, st[n][2] = p2, st[n][3] = p3;
                dpt[n][0] = p0 * (dpt[n][0] - Da), dpt[n][1] = p1 * (dpt[n][1] - Db);
                dpt[n][2] = p2 * (dpt[n][2] - Da), dpt[n][3] = p3 * (dpt[n][3] - Db);
            }
            mma_p_rows(dv, st, sdO, qc, nqb, lane);               // dV += P^T dO
            mma_p_rows(dk, dpt, sQ, qc, nqb, lane);               // dK += dS^T Q
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = k0 + g + 8 * half;                      // staged key row
            if (r >= Lkp) continue;
            if (has_prefix && r == 0) {
                float *dst = d.dprefix + ((static_cast<int64_t>(i) * d.n_outer + o) * d.n_heads + h) * 2 * 64;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    dst[n * 8 + t4 * 2] = dk[n][2 * half] * d.scale, dst[n * 8 + t4 * 2 + 1] = dk[n][2 * half + 1] * d.scale;
                    dst[64 + n * 8 + t4 * 2] = dv[n][2 * half], dst[64 + n * 8 + t4 * 2 + 1] = dv[n][2 * half + 1];
                }
            } else {
                const int64_t off = o * d.kv_outer + i * d.kv_inner + h * 64 + static_cast<int64_t>(r - has_prefix) * d.kv_row;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    *reinterpret_cast<uint32_t *>(d.dk + off + n * 8 + t4 * 2) = pack_bf16x2(dk[n][2 * half] * d.scale, dk[n][2 * half + 1] * d.scale);
                    *reinterpret_cast<uint32_t *>(d.dv + off + n * 8 + t4 * 2) = pack_bf16x2(dv[n][2 * half], dv[n][2 * half + 1]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// global query: one CTA (256 threads) per (outer, head).  Writes dq, and per key j the pair { scale * dS_j, P_j } to `coef` for the
// element-wise accumulation kernel below.
// ---------------------------------------------------------------------------------------------------------------------------
struct GDesc {
    const __nv_bfloat16 *q, *k, *v, *o, *d_o;
    __nv_bfloat16 *dq, *dk, *dv;
    const float *prefix_grad;   // (n_outer, n_heads, 2, HD) fp32 added to key row 0, or nullptr
    float *coef;                // (n_outer, n_heads, Lk, 2) fp32 scratch
    int64_t q_outer, kv_outer, kv_row, o_outer;
    int n_outer, n_heads, Lk;
    float scale, scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(256) global_query_bwd_stats_kernel(const GDesc d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s = reinterpret_cast<float *>(smem_raw);              // [Lk] scores, then dS
    __shared__ float qs[HD], dos[HD], red[8], part[4][HD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x / d.n_heads, h = blockIdx.x % d.n_heads;
    const __nv_bfloat16 *kb = d.k + o * d.kv_outer + h * HD, *vb = d.v + o * d.kv_outer + h * HD;
    float dl = 0.f;
    if (threadIdx.x < HD) {
        const float g = __bfloat162float(d.d_o[o * d.o_outer + h * HD + threadIdx.x]);
        qs[threadIdx.x] = __bfloat162float(d.q[o * d.q_outer + h * HD + threadIdx.x]) * d.scale_log2;
        dos[threadIdx.x] = g;
        dl = g * __bfloat162float(d.o[o * d.o_outer + h * HD + threadIdx.x]);
    }
    dl = warp_sum(dl);
    if (lane == 0) red[warp] = dl;
    __syncthreads();
    float delta = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) delta += red[w];
    __syncthreads();
    // scores (thread per key), block max and sum
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < d.Lk; j += 256) {
        const uint32_t *kr = reinterpret_cast<const uint32_t *>(kb + static_cast<int64_t>(j) * d.kv_row);
        float acc = 0.f;
#pragma unroll 8
        for (int w = 0; w < HD / 2; ++w) {
            const float2 kk = unpack_bf16x2(__ldg(kr + w));
            acc = fmaf(qs[2 * w], kk.x, fmaf(qs[2 * w + 1], kk.y, acc));
        }
        s[j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x; j < d.Lk; j += 256) sum += exp2f(s[j] - mx);
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    const float L = mx + log2f(sum);
    // dS_j = P_j (dO . V_j - D); coefficients for the accumulation pass
    float *cf = d.coef + (static_cast<int64_t>(blockIdx.x) * d.Lk) * 2;
    for (int j = threadIdx.x; j < d.Lk; j += 256) {
        const uint32_t *vr = reinterpret_cast<const uint32_t *>(vb + static_cast<int64_t>(j) * d.kv_row);
        float acc = 0.f;
#pragma unroll 8
        for (int w = 0; w < HD / 2; ++w) {
            const float2 vv = unpack_bf16x2(__ldg(vr + w));
            acc = fmaf(dos[2 * w], vv.x, fmaf(dos[2 * w + 1], vv.y, acc));
        }
        const float p = exp2f(s[j] - L);
        const float dsj = p * (acc - delta);
        s[j] = dsj;
        cf[2 * j] = dsj * d.scale;
        cf[2 * j + 1] = p;
    }
    __syncthreads();
    // dq[c] = scale sum_j dS_j k_j[c]: thread = (key slice, c)
    {
        const int c = threadIdx.x % HD, slice = threadIdx.x / HD, n_slices = 256 / HD;
        float acc = 0.f;
        for (int j = slice; j < d.Lk; j += n_slices) acc = fmaf(s[j], __bfloat162float(kb[static_cast<int64_t>(j) * d.kv_row + c]), acc);
        part[slice][c] = acc;
    }
    __syncthreads();
    if (threadIdx.x < HD) {
        float acc = 0.f;
        for (int sl = 0; sl < 256 / HD; ++sl) acc += part[sl][threadIdx.x];
        d.dq[o * d.q_outer + h * HD + threadIdx.x] = __float2bfloat16(acc * d.scale);
    }
}

// dK[o, j, h, :] += coef_ds * q[o, h, :],  dV[o, j, h, :] += coef_p * dO[o, h, :]  (+ prefix_grad on row 0): thread per (o, j, h, pair of c)
template <int HD>
__global__ void __launch_bounds__(256) global_query_bwd_accum_kernel(const GDesc d, int64_t n_items) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_items) return;
    const int c2 = static_cast<int>(idx % (HD / 2));
    int64_t t = idx / (HD / 2);
    const int h = static_cast<int>(t % d.n_heads);
    t /= d.n_heads;
    const int j = static_cast<int>(t % d.Lk);
    const int o = static_cast<int>(t / d.Lk);
    const float *cf = d.coef + ((static_cast<int64_t>(o) * d.n_heads + h) * d.Lk + j) * 2;
    const float cds = cf[0], cp = cf[1];
    const float2 q = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(d.q + o * d.q_outer + h * HD + 2 * c2));
    const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(d.d_o + o * d.o_outer + h * HD + 2 * c2));
    const int64_t off = o * d.kv_outer + static_cast<int64_t>(j) * d.kv_row + h * HD + 2 * c2;
    float2 dk = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(d.dk + off));
    float2 dv = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(d.dv + off));
    dk.x = fmaf(cds, q.x, dk.x), dk.y = fmaf(cds, q.y, dk.y);
    dv.x = fmaf(cp, g.x, dv.x), dv.y = fmaf(cp, g.y, dv.y);
    if (j == 0 && d.prefix_grad != nullptr) {
        const float *pg = d.prefix_grad + (static_cast<int64_t>(o) * d.n_heads + h) * 2 * HD + 2 * c2;
        dk.x += pg[0], dk.y += pg[1];
        dv.x += pg[HD], dv.y += pg[HD + 1];
    }
    *reinterpret_cast<uint32_t *>(d.dk + off) = pack_bf16x2(dk.x, dk.y);
    *reinterpret_cast<uint32_t *>(d.dv + off) = pack_bf16x2(dv.x, dv.y);
}

// ---------------------------------------------------------------------------------------------------------------------------
// DropPath (timm; vit_helper.py:371,375): out = residual + in * (keep(sample) ? 1 / (1 - p) : 0), one decision per sample
// (= rows_per_sample consecutive rows of 768).  Same counter-based generator as sfb_dropout, element index = sample index.
// ---------------------------------------------------------------------------------------------------------------------------
template <bool kOutBf16>
__global__ void __launch_bounds__(256) droppath_kernel(const float4 *__restrict__ in, const float4 *__restrict__ res, void *__restrict__ out,
                                                       int64_t n_vec, int64_t vec_per_sample, DropParams dp) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_vec) return;
    const float m = drop_scale(dp, static_cast<uint64_t>(idx / vec_per_sample));
    float4 v = __ldg(in + idx);
    v.x *= m, v.y *= m, v.z *= m, v.w *= m;
    if (res != nullptr) {
        const float4 rr = __ldg(res + idx);
        v.x += rr.x, v.y += rr.y, v.z += rr.z, v.w += rr.w;
    }
    if (kOutBf16)
        reinterpret_cast<uint2 *>(out)[idx] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    else
        reinterpret_cast<float4 *>(out)[idx] = v;
}

// out[r] (bf16, 768) = in[(r / group) * group_stride + offset + r % group] (fp32): warp per row
__global__ void __launch_bounds__(256) gather_rows_bf16_kernel(const float *__restrict__ in, int64_t ld, __nv_bfloat16 *__restrict__ out, int rows,
                                                               int group, int group_stride, int offset) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= rows) return;
    const int64_t src = static_cast<int64_t>(r / group) * group_stride + offset + (r % group);
    const float4 *ip = reinterpret_cast<const float4 *>(in + src * ld);
    uint2 *op = reinterpret_cast<uint2 *>(out + static_cast<int64_t>(r) * kD);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float4 v = __ldg(ip + lane + 32 * j);
        op[lane + 32 * j] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

static size_t smem_dq(int Lt, int HD) { return ((static_cast<size_t>(2) * Lt * (HD + 2) * 2 + 15) & ~static_cast<size_t>(15)) + sizeof(float) * kWarps * (((Lt + 3) & ~3) + 2 * HD); }
static size_t smem_dkv(int Lq, int HD) {
    const size_t Lqp = (Lq + 3) & ~3;
    return ((static_cast<size_t>(2) * Lq * (HD + 2) * 2 + 15) & ~static_cast<size_t>(15)) + sizeof(float) * (2 * Lqp + kWarps * (2 * Lqp + 2 * HD));
}
constexpr size_t kMaxSmem = 227 * 1024;

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    SFB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    return SFB_OK;
}

static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace attn_bwd
}  // namespace sfb

extern "C" int64_t sfb_attention_bwd_stats_floats(const sfb_attn_desc *f) {
    return f == nullptr ? 0 : static_cast<int64_t>(f->n_outer) * f->n_inner * f->n_heads * f->Lq * 2;
}

extern "C" int sfb_attention_bwd(const sfb_attn_desc *f, const void *d_out, void *dq, void *dk, void *dv, float *dprefix, float *stats,
                                 void *stream) {
    using namespace sfb;
    using namespace sfb::attn_bwd;
    SFB_CHECK_ARG(f && f->q && f->k && f->v && f->out && d_out && dq && dk && dv && stats, "sfb_attention_bwd: null pointer");
    SFB_CHECK_ARG(f->head_dim == 64 || f->head_dim == 96, "sfb_attention_bwd: head_dim %d not in {64, 96}", f->head_dim);
    SFB_CHECK_ARG(f->n_outer > 0 && f->n_inner > 0 && f->n_heads > 0 && f->Lq > 0 && f->Lk > 0, "sfb_attention_bwd: bad sizes");
    SFB_CHECK_ARG((f->k_prefix == nullptr) == (f->v_prefix == nullptr) && (f->k_prefix == nullptr) == (dprefix == nullptr),
                  "sfb_attention_bwd: k_prefix, v_prefix and dprefix must be given together");
    SFB_CHECK_ARG(f->q_extra == nullptr, "sfb_attention_bwd: the fused extra query has its own backward (sfb_attention_bwd_global_query)");
    // 16-byte row chunks are staged with vector loads: every row start must be 16-byte aligned
    SFB_CHECK_ARG(al16(f->q) && al16(f->k) && al16(f->v) && al16(d_out) && (f->k_prefix == nullptr || (al16(f->k_prefix) && al16(f->v_prefix))) &&
                      f->q_outer % 8 == 0 && f->q_inner % 8 == 0 && f->q_row % 8 == 0 && f->kv_outer % 8 == 0 && f->kv_inner % 8 == 0 &&
                      f->kv_row % 8 == 0 && f->o_outer % 8 == 0 && f->o_inner % 8 == 0 && f->o_row % 8 == 0 && f->prefix_outer % 8 == 0,
                  "sfb_attention_bwd: pointers must be 16-byte aligned and strides multiples of 8 elements");
    const int Lt = f->Lk + (f->k_prefix ? 1 : 0);
    const size_t s1 = smem_dq(Lt, f->head_dim), s2 = smem_dkv(f->Lq, f->head_dim);
    if (s1 > kMaxSmem || s2 > kMaxSmem) {
        set_error("sfb_attention_bwd: Lq=%d Lk=%d need %zu / %zu bytes of shared memory (> 227 KB)", f->Lq, f->Lk, s1, s2);
        return SFB_E_UNSUPPORTED;
    }
    Desc d;
    d.q = reinterpret_cast<const __nv_bfloat16 *>(f->q), d.k = reinterpret_cast<const __nv_bfloat16 *>(f->k);
    d.v = reinterpret_cast<const __nv_bfloat16 *>(f->v), d.k_prefix = reinterpret_cast<const __nv_bfloat16 *>(f->k_prefix);
    d.v_prefix = reinterpret_cast<const __nv_bfloat16 *>(f->v_prefix), d.o = reinterpret_cast<const __nv_bfloat16 *>(f->out);
    d.d_o = reinterpret_cast<const __nv_bfloat16 *>(d_out);
    d.dq = reinterpret_cast<__nv_bfloat16 *>(dq), d.dk = reinterpret_cast<__nv_bfloat16 *>(dk), d.dv = reinterpret_cast<__nv_bfloat16 *>(dv);
    d.dprefix = dprefix, d.stats = stats;
    d.q_outer = f->q_outer, d.q_inner = f->q_inner, d.q_row = f->q_row, d.kv_outer = f->kv_outer, d.kv_inner = f->kv_inner, d.kv_row = f->kv_row;
    d.o_outer = f->o_outer, d.o_inner = f->o_inner, d.o_row = f->o_row, d.prefix_outer = f->prefix_outer;
    d.n_outer = f->n_outer, d.n_inner = f->n_inner, d.n_heads = f->n_heads, d.Lq = f->Lq, d.Lk = f->Lk;
    d.scale = f->scale, d.scale_log2 = f->scale * 1.4426950408889634f;
    const int64_t n_prob = static_cast<int64_t>(f->n_outer) * f->n_inner * f->n_heads;
    SFB_CHECK_ARG(n_prob < (1ll << 31), "sfb_attention_bwd: too many problems for one launch");
    cudaStream_t st0 = reinterpret_cast<cudaStream_t>(stream);
    // large head-dim-64 problems (space attention): the mma.sync kernel; impl == 1 or SFB_ATTN_BWD_MMA=0 keeps the CUDA-core pair
    static const bool mma_enabled = !(getenv("SFB_ATTN_BWD_MMA") && atoi(getenv("SFB_ATTN_BWD_MMA")) == 0);
    const int Lq_pad = (f->Lq + 15) & ~15, Lk_pad = (Lt + 15) & ~15;
    const bool al4 = ((reinterpret_cast<uintptr_t>(f->out) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 3) == 0;
    if (mma_enabled && f->impl != 1 && f->head_dim == 64 && f->Lq >= 64 && Lq_pad <= 256 && Lk_pad <= 256 && al4) {      // 4-byte O loads / gradient stores
        const size_t smem = static_cast<size_t>(2 * Lq_pad + 2 * Lk_pad) * kPitch + 2 * Lq_pad * sizeof(float);
        int rc0;
        if ((rc0 = set_smem(attn_bwd_mma_kernel, smem)) != SFB_OK) return rc0;
        attn_bwd_mma_kernel<<<static_cast<unsigned>(n_prob), kMmaWarps * 32, smem, st0>>>(d, Lq_pad, Lk_pad);
        SFB_CHECK_LAUNCH();
        return SFB_OK;
    }
    const dim3 g1(static_cast<unsigned>(n_prob), (f->Lq + kRowsPerCta - 1) / kRowsPerCta), g2(static_cast<unsigned>(n_prob), (Lt + kRowsPerCta - 1) / kRowsPerCta);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    if (f->head_dim == 64) {
        if ((rc = set_smem(attn_bwd_dq_kernel<64>, s1)) != SFB_OK) return rc;
        if ((rc = set_smem(attn_bwd_dkv_kernel<64>, s2)) != SFB_OK) return rc;
        attn_bwd_dq_kernel<64><<<g1, kWarps * 32, s1, st>>>(d);
        SFB_CHECK_LAUNCH();
        attn_bwd_dkv_kernel<64><<<g2, kWarps * 32, s2, st>>>(d);
    } else {
        if ((rc = set_smem(attn_bwd_dq_kernel<96>, s1)) != SFB_OK) return rc;
        if ((rc = set_smem(attn_bwd_dkv_kernel<96>, s2)) != SFB_OK) return rc;
        attn_bwd_dq_kernel<96><<<g1, kWarps * 32, s1, st>>>(d);
        SFB_CHECK_LAUNCH();
        attn_bwd_dkv_kernel<96><<<g2, kWarps * 32, s2, st>>>(d);
    }
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_attention_bwd_global_query(const void *q, int64_t q_outer, const void *k, const void *v, int64_t kv_outer, int64_t kv_row,
                                              const void *out, const void *d_out, int64_t o_outer, void *dq, void *dk, void *dv,
                                              const float *prefix_grad, float *coef, int n_outer, int n_heads, int head_dim, int Lk, float scale,
                                              void *stream) {
    using namespace sfb;
    using namespace sfb::attn_bwd;
    SFB_CHECK_ARG(q && k && v && out && d_out && dq && dk && dv && coef, "sfb_attention_bwd_global_query: null pointer");
    SFB_CHECK_ARG(head_dim == 64, "sfb_attention_bwd_global_query: head_dim %d (only 64: the Motionformer CLS query)", head_dim);
    SFB_CHECK_ARG(n_outer > 0 && n_heads > 0 && Lk > 0 && Lk <= 12288, "sfb_attention_bwd_global_query: bad sizes (Lk <= 12288)");
    SFB_CHECK_ARG(q_outer % 2 == 0 && kv_outer % 2 == 0 && kv_row % 2 == 0 && o_outer % 2 == 0 &&
                      (reinterpret_cast<uintptr_t>(q) & 3) == 0 && (reinterpret_cast<uintptr_t>(k) & 3) == 0 && (reinterpret_cast<uintptr_t>(v) & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(d_out) & 3) == 0 && (reinterpret_cast<uintptr_t>(dk) & 3) == 0 && (reinterpret_cast<uintptr_t>(dv) & 3) == 0,
                  "sfb_attention_bwd_global_query: pointers must be 4-byte aligned and strides even");
    GDesc d;
    d.q = reinterpret_cast<const __nv_bfloat16 *>(q), d.k = reinterpret_cast<const __nv_bfloat16 *>(k), d.v = reinterpret_cast<const __nv_bfloat16 *>(v);
    d.o = reinterpret_cast<const __nv_bfloat16 *>(out), d.d_o = reinterpret_cast<const __nv_bfloat16 *>(d_out);
    d.dq = reinterpret_cast<__nv_bfloat16 *>(dq), d.dk = reinterpret_cast<__nv_bfloat16 *>(dk), d.dv = reinterpret_cast<__nv_bfloat16 *>(dv);
    d.prefix_grad = prefix_grad, d.coef = coef;
    d.q_outer = q_outer, d.kv_outer = kv_outer, d.kv_row = kv_row, d.o_outer = o_outer;
    d.n_outer = n_outer, d.n_heads = n_heads, d.Lk = Lk, d.scale = scale, d.scale_log2 = scale * 1.4426950408889634f;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(Lk) * sizeof(float);
    int rc;
    if ((rc = set_smem(global_query_bwd_stats_kernel<64>, smem)) != SFB_OK) return rc;
    global_query_bwd_stats_kernel<64><<<n_outer * n_heads, 256, smem, st>>>(d);
    SFB_CHECK_LAUNCH();
    const int64_t n_items = static_cast<int64_t>(n_outer) * Lk * n_heads * (64 / 2);
    global_query_bwd_accum_kernel<64><<<static_cast<unsigned>((n_items + 255) / 256), 256, 0, st>>>(d, n_items);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_droppath(const float *in, const float *residual, void *out, int out_bf16, int64_t n_rows, int rows_per_sample, float p,
                            uint64_t seed, uint32_t site, void *stream) {
    using namespace sfb;
    using namespace sfb::attn_bwd;
    SFB_CHECK_ARG(in && out && n_rows > 0 && rows_per_sample > 0 && n_rows % rows_per_sample == 0, "sfb_droppath: bad arguments");
    SFB_CHECK_ARG(al16(in) && al16(out) && al16(residual), "sfb_droppath: pointers must be 16-byte aligned");
    SFB_CHECK_ARG(p >= 0.f && p < 1.f, "sfb_droppath: p=%f outside [0, 1)", p);
    const DropParams dp = make_drop_params(p, seed, site);
    const int64_t n_vec = n_rows * (kD / 4), vps = static_cast<int64_t>(rows_per_sample) * (kD / 4);
    const unsigned grid = static_cast<unsigned>((n_vec + 255) / 256);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out_bf16)
        droppath_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<const float4 *>(residual), out, n_vec, vps, dp);
    else
        droppath_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<const float4 *>(residual), out, n_vec, vps, dp);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_gather_rows_bf16(const float *in, int64_t ld, void *out, int rows, int group, int group_stride, int offset, void *stream) {
    using namespace sfb;
    using namespace sfb::attn_bwd;
    SFB_CHECK_ARG(in && out && rows > 0 && group > 0 && ld >= kD && ld % 4 == 0 && al16(in) && al16(out), "sfb_gather_rows_bf16: bad arguments");
    gather_rows_bf16_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, ld, reinterpret_cast<__nv_bfloat16 *>(out), rows,
                                                                                                group, group_stride, offset);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
