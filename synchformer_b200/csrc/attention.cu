// K6-K11: softmax attention on strided bf16 Q/K/V views of the fused qkv activations (see sfb_attn_desc).
//
// Kernels behind one entry point (plus attn_space_tc_kernel in attention_tc.cu: tcgen05 / TMEM, used for the 196 x 197 space attention):
//   attn_mma_kernel<HD>    Lq >= 16: persistent CTAs loop over (outer, inner, head) problems.  Q/K/V rows are staged in padded
//                          shared memory with cp.async, double-buffered so the next problem streams in while this one is
//                          computed; each warp owns 16 query rows and walks the keys in chunks of 64 with
//                          mma.sync.m16n8k16 (bf16 in, fp32 accumulate) for QK^T and PV and a register-resident online
//                          softmax (quad shuffles only).  Space attention 196x197, AST 74x74, sync 198x198 (hd 96).
//   attn_time_mma_kernel   Lq = Lk = 8 + CLS prefix, hd 64 (Motionformer time attention): two locations per block-diagonal
//                          m16n8k16 problem, one warp per head, one CTA per location pair (two resident per SM).
//   attn_time_kernel       the same on CUDA cores (one thread per (query frame, head)); kept for odd location counts.
//   attn_row1_kernel       Lq = 1, hd 64 (Motionformer CLS query 1x1569, aggregator CLS rows 1x197 / 1x13): one CTA per problem,
//                          8 lanes per key row, 32 private online-softmax states merged through shared memory.
//   attn_generic_kernel<HD> one warp per query row, lanes split the head dim: the bring-up cross-check for the other three.
// All keep scores and statistics in fp32; nothing is materialised in HBM (the reference materialises the attention
// matrix and ~20 rearrange/cat copies per block: vit_helper.py:34-42,106-153; modeling_ast.py:156-176;
// modules/transformer.py:67-70).
#include <stdlib.h>

#include "common.cuh"
#include "attention.cuh"

namespace sfb {
namespace attn {


__device__ __forceinline__ void bf16x8_to_f32(const uint4 u, float *f) {
    f[0] = __uint_as_float(u.x << 16), f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16), f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16), f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16), f[7] = __uint_as_float(u.w & 0xffff0000u);
}

// ------------------------------------------------------------------------------------------- generic kernel
template <int HD>
__global__ void __launch_bounds__(256) attn_generic_kernel(const Desc d) {
    constexpr int E = HD / 32;  // elements per lane
    const int lane = threadIdx.x & 31;
    const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int64_t total = static_cast<int64_t>(d.n_outer) * d.n_inner * d.n_heads * d.Lq;
    if (wid >= total) return;
    const int r = static_cast<int>(wid % d.Lq);
    int64_t t = wid / d.Lq;
    const int h = static_cast<int>(t % d.n_heads);
    t /= d.n_heads;
    const int i = static_cast<int>(t % d.n_inner);
    const int o = static_cast<int>(t / d.n_inner);

    const __nv_bfloat16 *qp = d.q + o * d.q_outer + i * d.q_inner + static_cast<int64_t>(r) * d.q_row + h * HD + lane * E;
    const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * HD + lane * E;
    const int64_t pre_base = o * d.prefix_outer + h * HD + lane * E;
    float q[E], acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) q[e] = __bfloat162float(qp[e]) * d.scale, acc[e] = 0.f;
    float m = -INFINITY, l = 0.f;
    const int Lkp = d.Lk + d.has_prefix;
    for (int j0 = 0; j0 < Lkp; j0 += 4) {
        float kk[4][E], vv[4][E], s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j < Lkp) {
                const bool pre = d.has_prefix && j == 0;
                const int64_t off = pre ? pre_base : kv_base + static_cast<int64_t>(j - d.has_prefix) * d.kv_row;
                const __nv_bfloat16 *kr = (pre ? d.kp : d.k) + off, *vr = (pre ? d.vp : d.v) + off;
#pragma unroll
                for (int e = 0; e < E; ++e) kk[u][e] = __bfloat162float(kr[e]), vv[u][e] = __bfloat162float(vr[e]);
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) kk[u][e] = 0.f, vv[u][e] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float p = 0.f;
#pragma unroll
            for (int e = 0; e < E; ++e) p = fmaf(q[e], kk[u][e], p);
            s[u] = (j0 + u < Lkp) ? warp_sum(p) : -INFINITY;
        }
        const float m_new = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), m);
        const float corr = __expf(m - m_new);
        l *= corr;
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] *= corr;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float p = __expf(s[u] - m_new);
            l += p;
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = fmaf(p, vv[u][e], acc[e]);
        }
        m = m_new;
    }
    __nv_bfloat16 *op = d.out + o * d.o_outer + i * d.o_inner + static_cast<int64_t>(r) * d.o_row + h * HD + lane * E;
    const float inv = 1.0f / l;
#pragma unroll
    for (int e = 0; e < E; ++e) op[e] = __float2bfloat16_rn(acc[e] * inv);
}

// ------------------------------------------------------------------------------ Motionformer time attention
// lane = frame (0..7) + 8 * (head within a group of 4); a warp covers one (segment, position, head-group).

__global__ void __launch_bounds__(256) attn_time_kernel(const Desc d) {
    constexpr int HD = 64, L = 8;
    const int lane = threadIdx.x & 31;
    const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int hg_per = d.n_heads / 4;
    const int64_t total = static_cast<int64_t>(d.n_outer) * d.n_inner * hg_per;
    if (wid >= total) return;
    const int hg = static_cast<int>(wid % hg_per);
    const int64_t t = wid / hg_per;
    const int i = static_cast<int>(t % d.n_inner);
    const int o = static_cast<int>(t / d.n_inner);
    const int f = lane & 7;
    const int h = hg * 4 + (lane >> 3);

    const uint4 *qp = reinterpret_cast<const uint4 *>(d.q + o * d.q_outer + i * d.q_inner + static_cast<int64_t>(f) * d.q_row + h * HD);
    float q[HD];
#pragma unroll
    for (int c = 0; c < 8; ++c) bf16x8_to_f32(__ldg(qp + c), q + 8 * c);
    const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * HD;
    const int64_t pre_base = o * d.prefix_outer + h * HD;

    float s[L + 1];
#pragma unroll
    for (int j = 0; j < L + 1; ++j) {
        const uint4 *kp = reinterpret_cast<const uint4 *>(j == 0 ? d.kp + pre_base : d.k + kv_base + static_cast<int64_t>(j - 1) * d.kv_row);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float kf[8];
            bf16x8_to_f32(__ldg(kp + c), kf);
#pragma unroll
            for (int e = 0; e < 8; e += 2) a0 = fmaf(q[8 * c + e], kf[e], a0), a1 = fmaf(q[8 * c + e + 1], kf[e + 1], a1);
        }
        s[j] = (a0 + a1) * d.scale;
    }
    float m = s[0];
#pragma unroll
    for (int j = 1; j < L + 1; ++j) m = fmaxf(m, s[j]);
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < L + 1; ++j) s[j] = __expf(s[j] - m), l += s[j];
    const float inv = 1.0f / l;
    // reuse q[] as the output accumulator
#pragma unroll
    for (int e = 0; e < HD; ++e) q[e] = 0.f;
#pragma unroll
    for (int j = 0; j < L + 1; ++j) {
        const uint4 *vp = reinterpret_cast<const uint4 *>(j == 0 ? d.vp + pre_base : d.v + kv_base + static_cast<int64_t>(j - 1) * d.kv_row);
        const float p = s[j] * inv;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float vf[8];
            bf16x8_to_f32(__ldg(vp + c), vf);
#pragma unroll
            for (int e = 0; e < 8; ++e) q[8 * c + e] = fmaf(p, vf[e], q[8 * c + e]);
        }
    }
    uint4 *op = reinterpret_cast<uint4 *>(d.out + o * d.o_outer + i * d.o_inner + static_cast<int64_t>(f) * d.o_row + h * HD);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(q[8 * c], q[8 * c + 1]);
        u.y = pack_bf16x2(q[8 * c + 2], q[8 * c + 3]);
        u.z = pack_bf16x2(q[8 * c + 4], q[8 * c + 5]);
        u.w = pack_bf16x2(q[8 * c + 6], q[8 * c + 7]);
        op[c] = u;
    }
}

// ------------------------------------------------------------------------------- single-query (CLS) kernel
// Lq == 1, hd 64: one CTA of 256 threads per (outer, inner, head).  Eight lanes share a key row (16 bytes = 8 dims each, so one
// load instruction covers 4 rows x 128 contiguous bytes), the 32 lane-groups of the CTA stride over the keys with a private
// online softmax (3 shuffles per key), K and V of 4 keys in flight per lane, and the 32 partial (max, sum, acc) states are
// merged through shared memory.  One pass over K and V: HBM-bound at large batch, latency-tolerant at small batch.
__global__ void __launch_bounds__(256, 4) attn_row1_kernel(const Desc d) {
    constexpr int HD = 64, G = 32;           // lane groups per CTA
    __shared__ float s_m[G], s_l[G];
    __shared__ float s_acc[G][HD + 4];
    const int tid = threadIdx.x, lane = tid & 31;
    const int sub = lane & 7;                // which 8 dims of the row
    const int grp = tid >> 3;                // 0..31
    int pidx = blockIdx.x;
    const int h = pidx % d.n_heads;
    pidx /= d.n_heads;
    const int i = pidx % d.n_inner;
    const int o = pidx / d.n_inner;
    const int Lkp = d.Lk + d.has_prefix;

    float q[8];
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(d.q + o * d.q_outer + i * d.q_inner + h * HD) + sub), q);
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] *= d.scale;
    const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * HD + sub * 8;
    const int64_t pre_base = o * d.prefix_outer + h * HD + sub * 8;

    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const int n_iter = (Lkp + 4 * G - 1) / (4 * G);      // CTA-uniform trip count: the shuffles below need every lane of the warp
    for (int it = 0; it < n_iter; ++it) {
        const int j0 = grp + it * 4 * G;
        uint4 kk[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * G;
            if (j < Lkp) {
                const bool pre = d.has_prefix && j == 0;
                const int64_t off = pre ? pre_base : kv_base + static_cast<int64_t>(j - d.has_prefix) * d.kv_row;
                kk[u] = __ldg(reinterpret_cast<const uint4 *>((pre ? d.kp : d.k) + off));
                vv[u] = __ldg(reinterpret_cast<const uint4 *>((pre ? d.vp : d.v) + off));
            } else {
                kk[u] = make_uint4(0, 0, 0, 0), vv[u] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float kf[8], vf[8];
            bf16x8_to_f32(kk[u], kf);
            float sdot = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) sdot = fmaf(q[e], kf[e], sdot);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
            if (j0 + u * G < Lkp) {          // uniform within the 8-lane group
                const float m_new = fmaxf(m, sdot);
                const float corr = __expf(m - m_new), pj = __expf(sdot - m_new);
                bf16x8_to_f32(vv[u], vf);
                l = l * corr + pj;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf[e], acc[e] * corr);
                m = m_new;
            }
        }
    }
    if (sub == 0) s_m[grp] = m, s_l[grp] = l;
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[grp][sub * 8 + e] = acc[e];
    __syncthreads();
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll 8
        for (int g = 0; g < G; ++g) M = fmaxf(M, s_m[g]);
        float L = 0.f, out = 0.f;
#pragma unroll 8
        for (int g = 0; g < G; ++g) {
            const float w = s_m[g] == -INFINITY ? 0.f : __expf(s_m[g] - M);
            L = fmaf(s_l[g], w, L);
            out = fmaf(s_acc[g][tid], w, out);
        }
        d.out[o * d.o_outer + i * d.o_inner + h * HD + tid] = __float2bfloat16_rn(out / L);
    }
}

// ------------------------------------------------------------------------------------- tensor-core kernel
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(HD == 64 ? 512 : 416) attn_mma_kernel(const Desc d, int Lq_pad, int Lk_pad, int n_buf, int n_prob) {
    constexpr int PITCH = HD * 2 + 16;      // bytes per smem row; +16 keeps ldmatrix rows on distinct bank groups
    constexpr int CHUNKS = HD / 8;          // 16-byte chunks per row
    constexpr int KSTEPS = HD / 16;         // k16 steps over the head dim
    constexpr int NT_O = HD / 8;            // n8 tiles of the output
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t buf_bytes = static_cast<uint32_t>(Lq_pad + 2 * Lk_pad) * PITCH;
    const int Lkp = d.Lk + d.has_prefix;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int q0 = warp * 16;

    // K / V padding rows are zeroed once per buffer (0 * garbage must never produce NaN); cp.async only ever writes valid rows.
    // Q padding rows feed query rows that are never stored, so they need no initialisation.
    for (int bsel = 0; bsel < n_buf; ++bsel) {
        uint8_t *bk = smem + bsel * buf_bytes + Lq_pad * PITCH;
        for (int c = tid; c < (Lk_pad - Lkp) * CHUNKS; c += nthr) {
            const int r = Lkp + c / CHUNKS, cc = c % CHUNKS;
            *reinterpret_cast<uint4 *>(bk + r * PITCH + cc * 16) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(bk + (Lk_pad + r) * PITCH + cc * 16) = make_uint4(0, 0, 0, 0);
        }
    }

    // issue the cp.async loads of one problem's Q, K, V rows into buffer `bsel` (one commit group)
    auto issue_loads = [&](int prob, int bsel) {
        const int h = prob % d.n_heads;
        const int i = (prob / d.n_heads) % d.n_inner;
        const int o = prob / (d.n_heads * d.n_inner);
        const uint32_t sQ = s0 + bsel * buf_bytes, sK = sQ + Lq_pad * PITCH, sV = sK + Lk_pad * PITCH;
        const __nv_bfloat16 *qg = d.q + o * d.q_outer + i * d.q_inner + h * HD;
        for (int c = tid; c < d.Lq * CHUNKS; c += nthr) {
            const int r = c / CHUNKS, cc = c % CHUNKS;
            cp_async16(sQ + r * PITCH + cc * 16, qg + static_cast<int64_t>(r) * d.q_row + cc * 8);
        }
        const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * HD;
        const int64_t pre_base = o * d.prefix_outer + h * HD;
        for (int c = tid; c < Lkp * CHUNKS; c += nthr) {
            const int r = c / CHUNKS, cc = c % CHUNKS;
            const bool pre = d.has_prefix && r == 0;
            const int64_t off = (pre ? pre_base : kv_base + static_cast<int64_t>(r - d.has_prefix) * d.kv_row) + cc * 8;
            cp_async16(sK + r * PITCH + cc * 16, (pre ? d.kp : d.k) + off);
            cp_async16(sV + r * PITCH + cc * 16, (pre ? d.vp : d.v) + off);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const float sl2 = d.scale * 1.4426950408889634f;            // scores are kept in log2 units
    int bsel = 0;
    if (static_cast<int>(blockIdx.x) < n_prob) issue_loads(blockIdx.x, 0);
    // persistent loop: while the tensor cores work on problem p, cp.async streams problem p + gridDim.x into the other buffer
    for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                        // buffer `bsel` is complete and visible; the other buffer is free
        const int next = prob + gridDim.x;
        if (n_buf == 2 && next < n_prob) issue_loads(next, bsel ^ 1);

        const uint32_t sQ = s0 + bsel * buf_bytes, sK = sQ + Lq_pad * PITCH, sV = sK + Lk_pad * PITCH;
        // Q fragments for this warp's 16 rows (A operand, row-major): lanes 0-15 -> rows, lanes 16-31 -> +8 columns
        uint32_t qa[KSTEPS][4];
        {
            const uint32_t base = sQ + (q0 + (lane & 15)) * PITCH + (lane >> 4) * 16;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) ldmatrix_x4(base + ks * 32, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        }
        float oacc[NT_O][4];
#pragma unroll
        for (int n = 0; n < NT_O; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g+8

        for (int kc = 0; kc < Lk_pad; kc += 64) {
            const int nkb = min(4, (Lk_pad - kc) >> 4);   // 16-key blocks in this chunk (warp-uniform)
            float s[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            // ---- S = Q K^T : per 16-key block two n8 tiles; B fragments straight from the row-major K rows ----
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                if (kb < nkb) {
                    const uint32_t kaddr = sK + (kc + kb * 16 + (lane & 7) + ((lane >> 4) << 3)) * PITCH + ((lane >> 3) & 1) * 16;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        uint32_t b0, b1, b2, b3;
                        ldmatrix_x4(kaddr + ks * 32, b0, b1, b2, b3);
                        mma_bf16_16816(s[2 * kb], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
                        mma_bf16_16816(s[2 * kb + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
                    }
                }
            }
            // ---- scale, mask the padded keys, online softmax ----
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
            const float c0 = fast_exp2(m0 - mx0), c1 = fast_exp2(m1 - mx1);   // first chunk: exp2(-inf) = 0
            m0 = mx0, m1 = mx1;
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                s[n][0] = fast_exp2(s[n][0] - mx0), s[n][1] = fast_exp2(s[n][1] - mx0);
                s[n][2] = fast_exp2(s[n][2] - mx1), s[n][3] = fast_exp2(s[n][3] - mx1);
                ps0 += s[n][0] + s[n][1];
                ps1 += s[n][2] + s[n][3];
            }
            l0 = l0 * c0 + ps0;   // per-thread partial sums; reduced over the quad at the end
            l1 = l1 * c1 + ps1;
#pragma unroll
            for (int n = 0; n < NT_O; ++n) oacc[n][0] *= c0, oacc[n][1] *= c0, oacc[n][2] *= c1, oacc[n][3] *= c1;
            // ---- O += P V : P from the S accumulators (C layout == A layout), V^T fragments via ldmatrix.trans ----
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                if (kb < nkb) {
                    const uint32_t a0 = pack_bf16x2(s[2 * kb][0], s[2 * kb][1]);
                    const uint32_t a1 = pack_bf16x2(s[2 * kb][2], s[2 * kb][3]);
                    const uint32_t a2 = pack_bf16x2(s[2 * kb + 1][0], s[2 * kb + 1][1]);
                    const uint32_t a3 = pack_bf16x2(s[2 * kb + 1][2], s[2 * kb + 1][3]);
                    const uint32_t vaddr = sV + (kc + kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 16;
#pragma unroll
                    for (int n2 = 0; n2 < NT_O / 2; ++n2) {
                        uint32_t b0, b1, b2, b3;
                        ldmatrix_x4_trans(vaddr + n2 * 32, b0, b1, b2, b3);
                        mma_bf16_16816(oacc[2 * n2], a0, a1, a2, a3, b0, b1);
                        mma_bf16_16816(oacc[2 * n2 + 1], a0, a1, a2, a3, b2, b3);
                    }
                }
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;

        // ---- stage the 16 x HD output tile in this warp's (now dead) Q rows, then store 16 B per lane ----
        __syncwarp();
        uint8_t *so = smem + bsel * buf_bytes + q0 * PITCH;
#pragma unroll
        for (int n = 0; n < NT_O; ++n) {
            *reinterpret_cast<uint32_t *>(so + g * PITCH + n * 16 + t4 * 4) = pack_bf16x2(oacc[n][0] * i0, oacc[n][1] * i0);
            *reinterpret_cast<uint32_t *>(so + (g + 8) * PITCH + n * 16 + t4 * 4) = pack_bf16x2(oacc[n][2] * i1, oacc[n][3] * i1);
        }
        __syncwarp();
        {
            const int h = prob % d.n_heads;
            const int i = (prob / d.n_heads) % d.n_inner;
            const int o = prob / (d.n_heads * d.n_inner);
            __nv_bfloat16 *og = d.out + o * d.o_outer + i * d.o_inner + h * HD;
            for (int c = lane; c < 16 * CHUNKS; c += 32) {
                const int r = c / CHUNKS, cc = c % CHUNKS;
                if (q0 + r < d.Lq)
                    *reinterpret_cast<uint4 *>(og + static_cast<int64_t>(q0 + r) * d.o_row + cc * 8) = *reinterpret_cast<const uint4 *>(so + r * PITCH + cc * 16);
            }
        }
        if (n_buf == 2) {
            bsel ^= 1;
        } else {
            __syncthreads();                                    // single buffer: everybody is done before it is refilled
            if (next < n_prob) issue_loads(next, 0);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// --------------------------------------------------------------- Motionformer time attention on tensor cores
// Lq = Lk = 8 frames + CLS prefix key, hd 64.  One CTA handles one PAIR of spatial locations for all heads (one warp per head,
// THREE CTAs per SM so that loads, math and stores of different pairs overlap: 56 registers per thread and exactly 75 KB of shared memory -
// the rows are 128B-swizzled (16-byte chunk ^ (row & 7)) instead of padded, which is what makes the third CTA fit): the
// 16 x (3 * n_heads * 64) tile [2 locations x 8 frames] x [q | k | v] is
// staged with cp.async (fully coalesced 1.5 KB row pieces), and the two 8 x 9 attentions of a pair are evaluated as ONE block-
// diagonal m16n8k16 problem: S = [Q_a; Q_b] [K_a | K_b | k_cls]^T keeps the two diagonal 8 x 8 blocks and the CLS column,
// O = P V with P zero off the diagonal blocks.  28 mma.sync per (pair, head) instead of 2 x 9 x 64 x 2 x 8 CUDA-core FMAs.
template <int H, bool kExtra = false>   // heads == warps per CTA; kExtra: also serve the fused extra (CLS) query, see the end of the kernel
__global__ void __launch_bounds__(H * 32, 3) attn_time_mma_kernel(const Desc d) {
    constexpr int HD = 64;
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int seg_bytes = H * HD * 2;             // one of q / k / v for one token, all heads
    constexpr int PITCH = 3 * seg_bytes;              // no padding: chunk c of row r is stored at chunk c ^ (r & 7), so ldmatrix rows land on distinct bank groups
    constexpr int cls_off = 16 * PITCH;               // [k_cls | v_cls] of the segment
    constexpr int chunks_seg = seg_bytes / 16;
    const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;       // warp == head
    const int g = lane >> 2, t4 = lane & 3;
    const int pairs_per_outer = d.n_inner / 2;
    const int pair = blockIdx.x;                      // one CTA per pair of locations; two CTAs per SM overlap load and math
    const int o = pair / pairs_per_outer;
    const int i0p = (pair % pairs_per_outer) * 2;
    // 48 (row, part) segments of seg_bytes each + the CLS k / v segments: warp-uniform source pointers, lanes walk the 16-byte chunks
    for (int sgm = warp; sgm < 50; sgm += H) {
        const __nv_bfloat16 *src;
        uint32_t dst;
        if (sgm < 48) {
            const int r = sgm / 3, part = sgm % 3;    // r = location (r >> 3), frame (r & 7); part 0 q, 1 k, 2 v
            const int64_t tok = static_cast<int64_t>(i0p + (r >> 3));
            src = part == 0 ? d.q + o * d.q_outer + tok * d.q_inner + static_cast<int64_t>(r & 7) * d.q_row
                            : (part == 1 ? d.k : d.v) + o * d.kv_outer + tok * d.kv_inner + static_cast<int64_t>(r & 7) * d.kv_row;
            dst = s0 + r * PITCH + part * seg_bytes;
        } else {
            src = (sgm == 48 ? d.kp : d.vp) + o * d.prefix_outer;
            dst = s0 + cls_off + (sgm - 48) * seg_bytes;
        }
        const int swz = sgm < 48 ? ((sgm / 3) & 7) : 0;             // the CLS segments are read as one broadcast row: not swizzled
        for (int cc = lane; cc < chunks_seg; cc += 32) cp_async16(dst + ((cc ^ swz) << 4), src + cc * 8);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float sl2 = d.scale * 1.4426950408889634f;
    {
        {
            const uint32_t sb = s0;
            const uint32_t hq = sb + warp * (HD * 2), hk = hq + seg_bytes, hv = hk + seg_bytes;
            const uint32_t ck = sb + cls_off + warp * (HD * 2), cv = ck + seg_bytes;
            float s[3][4];
#pragma unroll
            for (int n = 0; n < 3; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t a0, a1, a2, a3, b0, b1, b2, b3, c0, c1;
                ldmatrix_x4(hq + (lane & 15) * PITCH + ((((lane >> 4) + ks * 2) ^ (lane & 7)) << 4), a0, a1, a2, a3);
                ldmatrix_x4(hk + ((lane & 7) + ((lane >> 4) << 3)) * PITCH + (((((lane >> 3) & 1) + ks * 2) ^ (lane & 7)) << 4), b0, b1, b2, b3);
                // CLS key: every row address of the 8 x 8 matrices points at the single k_cls row (columns 1..7 are duplicates, ignored)
                asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(c0), "=r"(c1) : "r"(ck + ((lane >> 3) & 1) * 16 + ks * 32));
                mma_bf16_16816(s[0], a0, a1, a2, a3, b0, b1);     // x keys of location a: rows 0-7 valid
                mma_bf16_16816(s[1], a0, a1, a2, a3, b2, b3);     // x keys of location b: rows 8-15 valid
                mma_bf16_16816(s[2], a0, a1, a2, a3, c0, c1);     // x CLS key: column 0 valid
            }
            // row g belongs to location a (scores s[0][0..1] over the quad), row g+8 to location b (s[1][2..3]); CLS score from lane t4 == 0
            const float cls0 = __shfl_sync(0xffffffffu, s[2][0], lane & ~3) * sl2;
            const float cls1 = __shfl_sync(0xffffffffu, s[2][2], lane & ~3) * sl2;
            float x0 = s[0][0] * sl2, x1 = s[0][1] * sl2, y0 = s[1][2] * sl2, y1 = s[1][3] * sl2;
            float mx0 = fmaxf(fmaxf(x0, x1), cls0), mx1 = fmaxf(fmaxf(y0, y1), cls1);
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            x0 = fast_exp2(x0 - mx0), x1 = fast_exp2(x1 - mx0), y0 = fast_exp2(y0 - mx1), y1 = fast_exp2(y1 - mx1);
            const float pc0 = fast_exp2(cls0 - mx0), pc1 = fast_exp2(cls1 - mx1);
            float l0 = x0 + x1, l1 = y0 + y1;
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            const float i0 = 1.0f / (l0 + pc0), i1 = 1.0f / (l1 + pc1);
            // P (block diagonal) as the A operand of two k16 steps: [keys a | keys b] and [CLS, 15 zeros]
            const uint32_t pa0 = pack_bf16x2(x0 * i0, x1 * i0), pa3 = pack_bf16x2(y0 * i1, y1 * i1);
            const uint32_t pb0 = t4 == 0 ? pack_bf16x2(pc0 * i0, 0.f) : 0u, pb1 = t4 == 0 ? pack_bf16x2(pc1 * i1, 0.f) : 0u;
            float oacc[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(hv + ((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + ((((lane >> 4) + n2 * 2) ^ (lane & 7)) << 4), b0, b1, b2, b3);
                mma_bf16_16816(oacc[2 * n2], pa0, 0u, 0u, pa3, b0, b1);
                mma_bf16_16816(oacc[2 * n2 + 1], pa0, 0u, 0u, pa3, b2, b3);
                // CLS value: all 16 "key" rows of this k16 step point at v_cls (finite), only key 0 has a non-zero probability
                ldmatrix_x4_trans(cv + (lane >> 4) * 16 + n2 * 32, b0, b1, b2, b3);
                mma_bf16_16816(oacc[2 * n2], pb0, pb1, 0u, 0u, b0, b1);
                mma_bf16_16816(oacc[2 * n2 + 1], pb0, pb1, 0u, 0u, b2, b3);
            }
            // stage the 16 x 64 output in this head's (dead) Q columns, then 16-byte coalesced stores
            __syncwarp();
            uint8_t *so = smem + warp * (HD * 2);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                *reinterpret_cast<uint32_t *>(so + g * PITCH + ((n ^ g) << 4) + t4 * 4) = pack_bf16x2(oacc[n][0], oacc[n][1]);
                *reinterpret_cast<uint32_t *>(so + (g + 8) * PITCH + ((n ^ g) << 4) + t4 * 4) = pack_bf16x2(oacc[n][2], oacc[n][3]);
            }
            __syncwarp();
            __nv_bfloat16 *og = d.out + o * d.o_outer + warp * HD;
#pragma unroll
            for (int c = lane; c < 16 * 8; c += 32) {
                const int r = c >> 3, cc = c & 7;
                *reinterpret_cast<uint4 *>(og + static_cast<int64_t>(i0p + (r >> 3)) * d.o_inner + static_cast<int64_t>(r & 7) * d.o_row + cc * 8) =
                    *reinterpret_cast<const uint4 *>(so + r * PITCH + ((cc ^ (r & 7)) << 4));
            }
        }
    }
    if (kExtra) {
        // Fused extra query (the Motionformer CLS query of the time attention, vit_helper.py:124): one more query row per (outer, head)
        // that attends to ALL keys of its outer index.  This CTA holds the 8 keys / values of its two locations (and the CLS key / value):
        // per head it emits the query's softmax state over each location's keys - { max (log2 units), sum, out / sum } - to
        // xpartial[((o * H + head) * n_inner + location) * 66]; the CLS key is counted for location 0 only; sfb_attention_merge_partials
        // combines the n_inner states.  17 dot products of 64 per warp on the CUDA cores: nothing next to the HBM traffic of the kernel.
        const float2 qx = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(d.xq + o * d.xq_outer + warp * HD + lane * 2));
        const uint8_t *kbase = smem + warp * (HD * 2) + seg_bytes, *vbase = kbase + seg_bytes;
        const uint8_t *kcls = smem + cls_off + warp * (HD * 2), *vcls = kcls + seg_bytes;
        float sc = 0.f;
        for (int j = 0; j < 17; ++j) {
            const float2 kk = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(j < 16 ? kbase + j * PITCH + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4 : kcls + lane * 4));
            const float p = warp_sum(qx.x * kk.x + qx.y * kk.y);
            if (lane == j) sc = p * sl2;
        }
        const bool in_a = lane < 8 || (lane == 16 && i0p == 0), in_b = lane >= 8 && lane < 16;
        const float ma = warp_max(in_a ? sc : -INFINITY), mb = warp_max(in_b ? sc : -INFINITY);
        const float pa = in_a ? exp2f(sc - ma) : 0.f, pb = in_b ? exp2f(sc - mb) : 0.f;
        const float la = warp_sum(pa), lb = warp_sum(pb);
        float2 oa = make_float2(0.f, 0.f), ob = make_float2(0.f, 0.f);
        for (int j = 0; j < 17; ++j) {
            const float wa = __shfl_sync(0xffffffffu, pa, j), wb = __shfl_sync(0xffffffffu, pb, j);
            const float2 vv = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(j < 16 ? vbase + j * PITCH + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4 : vcls + lane * 4));
            oa.x = fmaf(wa, vv.x, oa.x), oa.y = fmaf(wa, vv.y, oa.y);
            ob.x = fmaf(wb, vv.x, ob.x), ob.y = fmaf(wb, vv.y, ob.y);
        }
        float *dst_a = d.xpartial + ((static_cast<int64_t>(o) * H + warp) * d.n_inner + i0p) * (HD + 2), *dst_b = dst_a + (HD + 2);
        if (lane == 0) dst_a[0] = ma, dst_a[1] = la, dst_b[0] = mb, dst_b[1] = lb;
        dst_a[2 + 2 * lane] = oa.x / la, dst_a[3 + 2 * lane] = oa.y / la;
        dst_b[2 + 2 * lane] = ob.x / lb, dst_b[3 + 2 * lane] = ob.y / lb;
    }
}

}  // namespace attn
}  // namespace sfb

namespace sfb {
namespace attn {
__global__ void __launch_bounds__(64) merge_partials_kernel(const float *__restrict__ partial, __nv_bfloat16 *__restrict__ out, int64_t out_outer,
                                                            int n_inner, int n_heads, int hd) {
    const int h = blockIdx.x % n_heads, o = blockIdx.x / n_heads, d = threadIdx.x;
    if (d >= hd) return;
    const float *p = partial + static_cast<int64_t>(blockIdx.x) * n_inner * (hd + 2);
    float M = -INFINITY;
    for (int i = 0; i < n_inner; ++i) M = fmaxf(M, p[i * (hd + 2)]);
    float L = 0.f, acc = 0.f;
    for (int i = 0; i < n_inner; ++i) {
        const float w = p[i * (hd + 2) + 1] * exp2f(p[i * (hd + 2)] - M);
        L += w;
        acc = fmaf(w, p[i * (hd + 2) + 2 + d], acc);
    }
    out[o * out_outer + h * hd + d] = __float2bfloat16_rn(acc / L);
}
}  // namespace attn
}  // namespace sfb

static bool sfb_attn_aligned16(const sfb_attn_desc *d) {
    return ((reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->k) | reinterpret_cast<uintptr_t>(d->v) | reinterpret_cast<uintptr_t>(d->out) |
             reinterpret_cast<uintptr_t>(d->k_prefix) | reinterpret_cast<uintptr_t>(d->v_prefix)) & 15) == 0 &&
           ((d->q_outer | d->q_inner | d->q_row | d->kv_outer | d->kv_inner | d->kv_row | d->o_outer | d->o_inner | d->o_row | d->prefix_outer) % 8) == 0;
}

static bool sfb_time_extra_supported(const sfb_attn_desc *desc);

extern "C" int sfb_attention_extra_supported(const sfb_attn_desc *desc) {
    using namespace sfb::attn;
    if (desc == nullptr || desc->impl != 0 || desc->head_dim != 64 || !sfb_attn_aligned16(desc)) return 0;
    if (sfb_time_extra_supported(desc)) return 1;
    if (getenv("SFB_ATTN_TC") && atoi(getenv("SFB_ATTN_TC")) == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(desc->q_extra) & 15) != 0 || desc->q_extra_outer % 8 != 0) return 0;
    Desc d = {};
    d.Lq = desc->Lq, d.Lk = desc->Lk, d.has_prefix = desc->k_prefix != nullptr, d.scale = desc->scale;
    return tc_supported(d) && desc->Lq < 256 ? 1 : 0;     // the extra query occupies query row Lq of the second 128-row tile
}

// the block-diagonal time-attention kernel can serve the extra query too; opt-in (SFB_TIME_CLS_FUSED=1) until measured on hardware
static bool sfb_time_extra_supported(const sfb_attn_desc *desc) {
    static const bool enabled = getenv("SFB_TIME_CLS_FUSED") && atoi(getenv("SFB_TIME_CLS_FUSED")) == 1;
    return enabled && desc->impl == 0 && desc->head_dim == 64 && desc->Lq == 8 && desc->Lk == 8 && desc->k_prefix != nullptr && desc->n_inner % 2 == 0 &&
           desc->n_heads == 12 && sfb_attn_aligned16(desc) && (reinterpret_cast<uintptr_t>(desc->q_extra) & 15) == 0 && desc->q_extra_outer % 8 == 0;
}

extern "C" int sfb_attention_merge_partials(const float *partial, void *out, int64_t out_outer, int n_outer, int n_inner, int n_heads, int head_dim,
                                            void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(partial && out && n_outer > 0 && n_inner > 0 && n_heads > 0 && head_dim > 0 && head_dim <= 64, "sfb_attention_merge_partials: bad arguments");
    attn::merge_partials_kernel<<<static_cast<unsigned>(n_outer) * n_heads, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        partial, reinterpret_cast<__nv_bfloat16 *>(out), out_outer, n_inner, n_heads, head_dim);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_attention(const sfb_attn_desc *desc, void *stream) {
    using namespace sfb;
    using namespace sfb::attn;
    SFB_CHECK_ARG(desc != nullptr, "sfb_attention: null descriptor");
    SFB_CHECK_ARG(desc->q && desc->k && desc->v && desc->out, "sfb_attention: null pointer");
    SFB_CHECK_ARG(desc->head_dim == 64 || desc->head_dim == 96, "sfb_attention: head_dim %d not in {64, 96}", desc->head_dim);
    SFB_CHECK_ARG(desc->n_outer > 0 && desc->n_inner > 0 && desc->n_heads > 0 && desc->Lq > 0 && desc->Lk > 0, "sfb_attention: bad sizes");
    Desc d;
    d.q = reinterpret_cast<const __nv_bfloat16 *>(desc->q);
    d.k = reinterpret_cast<const __nv_bfloat16 *>(desc->k);
    d.v = reinterpret_cast<const __nv_bfloat16 *>(desc->v);
    d.out = reinterpret_cast<__nv_bfloat16 *>(desc->out);
    d.q_outer = desc->q_outer, d.q_inner = desc->q_inner, d.q_row = desc->q_row;
    d.kv_outer = desc->kv_outer, d.kv_inner = desc->kv_inner, d.kv_row = desc->kv_row;
    d.o_outer = desc->o_outer, d.o_inner = desc->o_inner, d.o_row = desc->o_row;
    SFB_CHECK_ARG((desc->k_prefix == nullptr) == (desc->v_prefix == nullptr), "sfb_attention: k_prefix / v_prefix must be given together");
    d.kp = reinterpret_cast<const __nv_bfloat16 *>(desc->k_prefix);
    d.vp = reinterpret_cast<const __nv_bfloat16 *>(desc->v_prefix);
    d.prefix_outer = desc->prefix_outer, d.has_prefix = desc->k_prefix != nullptr ? 1 : 0;
    d.n_outer = desc->n_outer, d.n_inner = desc->n_inner, d.n_heads = desc->n_heads, d.Lq = desc->Lq, d.Lk = desc->Lk;
    d.scale = desc->scale;
    d.xq = reinterpret_cast<const __nv_bfloat16 *>(desc->q_extra), d.xq_outer = desc->q_extra_outer, d.xpartial = desc->extra_partial;
    SFB_CHECK_ARG((d.xq == nullptr) == (d.xpartial == nullptr), "sfb_attention: q_extra and extra_partial must be given together");
    SFB_CHECK_ARG(d.xq == nullptr || sfb_attention_extra_supported(desc) == 1,
                  "sfb_attention: the fused extra query is not supported for this descriptor (ask sfb_attention_extra_supported first)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int HD = desc->head_dim;

    // vectorised kernels need 16-byte aligned rows
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(d.q) | reinterpret_cast<uintptr_t>(d.k) | reinterpret_cast<uintptr_t>(d.v) |
                             reinterpret_cast<uintptr_t>(d.out) | reinterpret_cast<uintptr_t>(d.kp) | reinterpret_cast<uintptr_t>(d.vp)) & 15) == 0 &&
                           ((d.q_outer | d.q_inner | d.q_row | d.kv_outer | d.kv_inner | d.kv_row | d.o_outer | d.o_inner | d.o_row |
                             d.prefix_outer) % 8) == 0;
    const int Lkp = d.Lk + d.has_prefix;
    const int Lq_pad = (d.Lq + 15) & ~15, Lk_pad = (Lkp + 15) & ~15;
    const int64_t mma_smem = static_cast<int64_t>(Lq_pad + 2 * Lk_pad) * (HD * 2 + 16);
    const int64_t n_prob = static_cast<int64_t>(d.n_outer) * d.n_inner * d.n_heads;

    static const bool tc_enabled = !(getenv("SFB_ATTN_TC") && atoi(getenv("SFB_ATTN_TC")) == 0);   // A/B switch for measurements
    if (desc->impl == 0 && aligned16 && HD == 64 && tc_enabled && tc_supported(d)) return launch_tc(d, st);
    if (desc->impl == 0 && aligned16 && d.Lq >= 16 && Lq_pad <= (HD == 64 ? 256 : 208) && mma_smem <= 200 * 1024) {
        const int threads = (Lq_pad / 16) * 32;
        SFB_CHECK_ARG(n_prob < (1ll << 31), "sfb_attention: too many problems");
        // two staging buffers (load of the next problem overlaps the math of this one) whenever they fit
        const int n_buf = 2 * mma_smem <= 220 * 1024 ? 2 : 1;
        const int smem_bytes = static_cast<int>(mma_smem) * n_buf;
        int occ = 1;
        if (HD == 64) {
            static int cur = 0;
            if (smem_bytes > cur) {
                SFB_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
                cur = smem_bytes;
            }
            SFB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, attn_mma_kernel<64>, threads, smem_bytes));
        } else {
            static int cur = 0;
            if (smem_bytes > cur) {
                SFB_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
                cur = smem_bytes;
            }
            SFB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, attn_mma_kernel<96>, threads, smem_bytes));
        }
        if (occ < 1) occ = 1;
        const int64_t max_ctas = static_cast<int64_t>(num_sms()) * occ;
        const unsigned grid = static_cast<unsigned>(n_prob < max_ctas ? n_prob : max_ctas);
        if (HD == 64)
            attn_mma_kernel<64><<<grid, threads, smem_bytes, st>>>(d, Lq_pad, Lk_pad, n_buf, static_cast<int>(n_prob));
        else
            attn_mma_kernel<96><<<grid, threads, smem_bytes, st>>>(d, Lq_pad, Lk_pad, n_buf, static_cast<int>(n_prob));
        SFB_CHECK_LAUNCH();
        return SFB_OK;
    }
    if (desc->impl == 0 && aligned16 && HD == 64 && d.Lq == 1) {
        SFB_CHECK_ARG(n_prob < (1ll << 31), "sfb_attention: too many problems");
        attn_row1_kernel<<<static_cast<unsigned>(n_prob), 256, 0, st>>>(d);
        SFB_CHECK_LAUNCH();
        return SFB_OK;
    }
    if (desc->impl == 0 && aligned16 && HD == 64 && d.Lq == 8 && d.Lk == 8 && d.has_prefix && d.n_inner % 2 == 0 && d.n_heads == 12) {
        constexpr int H = 12;
        constexpr int seg_bytes = H * 64 * 2;
        constexpr int smem_bytes = 16 * 3 * seg_bytes + 2 * seg_bytes;                   // exactly 75 KB: three CTAs per SM (3 x (75 + 1) KB = 228 KB)
        static PerDeviceOnce attr_once;
        if (attr_once.first()) {
            SFB_CHECK_CUDA(cudaFuncSetAttribute(attn_time_mma_kernel<H, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            SFB_CHECK_CUDA(cudaFuncSetAttribute(attn_time_mma_kernel<H, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        }
        const int64_t n_pairs = static_cast<int64_t>(d.n_outer) * (d.n_inner / 2);
        SFB_CHECK_ARG(n_pairs < (1ll << 31), "sfb_attention: too many problems");
        if (d.xq != nullptr)
            attn_time_mma_kernel<H, true><<<static_cast<unsigned>(n_pairs), 32 * H, smem_bytes, st>>>(d);
        else
            attn_time_mma_kernel<H, false><<<static_cast<unsigned>(n_pairs), 32 * H, smem_bytes, st>>>(d);
        SFB_CHECK_LAUNCH();
        return SFB_OK;
    }
    if (desc->impl == 0 && aligned16 && HD == 64 && d.Lq == 8 && d.Lk == 8 && d.has_prefix && d.n_heads % 4 == 0) {
        const int64_t warps = static_cast<int64_t>(d.n_outer) * d.n_inner * (d.n_heads / 4);
        attn_time_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, st>>>(d);
        SFB_CHECK_LAUNCH();
        return SFB_OK;
    }
    const int64_t warps = n_prob * d.Lq;
    SFB_CHECK_ARG((warps + 7) / 8 < (1ll << 31), "sfb_attention: too many rows");
    if (HD == 64)
        attn_generic_kernel<64><<<static_cast<unsigned>((warps + 7) / 8), 256, 0, st>>>(d);
    else
        attn_generic_kernel<96><<<static_cast<unsigned>((warps + 7) / 8), 256, 0, st>>>(d);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
