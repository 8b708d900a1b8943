// K7 on the 5th-generation tensor cores: Motionformer space attention (196 queries x 197 keys, hd 64; vit_helper.py:100-158 with
// einops_to '(b f) n d') as tcgen05.mma with S, P and O resident in TMEM.
//
// One persistent CTA per SM (384 threads) loops over (segment, frame, head) problems:
//   warps 9-11  producers   a 2-stage shared-memory ring of [Q | K | V] tiles in the 128B-swizzled K-major layout the UMMA descriptors
//                           expect.  When the row strides are regular (they are for the fused qkv activations) ONE thread issues
//                           three TMA box loads (cp.async.bulk.tensor, 196 x 64 each) per problem and a second warp gathers the
//                           single CLS key / value row with cp.async; otherwise three warps gather everything with cp.async.
//                           The CLS prefix row is stored LAST (key order is irrelevant to softmax), so the TMA boxes start on
//                           swizzle-atom boundaries
//   warp 8      MMA issuer  S_t = Q_t K^T   (2 row tiles t of 128 queries; 4 x tcgen05.mma M128 N208 K16, operands from smem)
//                           O_t = P_t V     (13 x tcgen05.mma M128 N64 K16, A = P_t read from TMEM, B = V as an MN-major operand:
//                           the [key][64 dims] rows are used as they are, no transpose)
//   warps 0-7   softmax     thread = query row (TMEM lane).  One sweep over the fp32 scores with software-pipelined tcgen05.ld
//                           (exp2 against a reference value, row sum, running maximum), P written back as packed bf16 with
//                           tcgen05.st, then the O epilogue: tcgen05.ld, 1/sum, bf16, rows into a swizzled staging tile that leaves
//                           through one TMA store per tile (issued by warp 11).
// TMEM (512 columns): tile t owns columns [256t, 256t+248): S fp32 in the first 208; P (bf16 pairs) of keys 0..127 re-uses columns
// [0,64) as the softmax consumes S, P of keys 128..207 goes to the spare columns [208,248), O accumulates in columns [64,128).
// Nothing touches HBM between the qkv activations and the attention output.
//
// What bounds it (round 2, clock64 phase stamps of one CTA, tools/attn_pipe.py; 1.76 -> 1.06 ms per launch at 512 segments):
//   * the MMA phases are short bursts whose cost is issue + ~400 clocks of commit latency, not tensor work: fully unrolled issue sequences
//     with constant-offset descriptors (56 instead of 150 clocks per tcgen05.mma, tools/ubench/mma_cost.cu);
//   * the two tiles of a problem must not run in lock-step: issue order S(0,i) PV(1,i-1) S(1,i) PV(0,i) puts one tile's exp2 pass (XU
//     pipe) next to the other tile's MMA / barrier latencies (tools/ubench/mma_commit_order.cu: a commit's barrier is not delayed by MMAs
//     issued after it);
//   * a row-per-thread st.global of the output (32 different lines per instruction) held the softmax warps ~1 800 clocks per problem:
//     rows go to a swizzled staging tile and leave through one TMA store per tile, issued by the otherwise idle producer warp;
//   * three integer divisions by run-time divisors at the top of every iteration (~430 clocks) moved behind the wait for P V;
//   * the loop body of the softmax warps has to fit the instruction cache: the masked variants of the chunk code are compiled out for
//     Lk >= 192 and the chunk loops are rolled (89 KB -> 58 KB of SASS: 1.37 -> 1.19 ms on its own).
//   * the loads of problem i+1 must not wait for the last P V of problem i-1: Q / K and V of a stage are handed back on separate barriers
//     (Q and K are dead once both S tiles exist), 1.19 -> 1.13 ms;
//   * softmax in ONE sweep over TMEM (reference value = maximum of the first 32 scores; a row whose sum of 2^(score - reference) reaches
//     2^110 is flagged and recomputed by attn_space_fixup_kernel) and the first 8 of the 13 P V instructions issued while the scores of
//     keys 128.. are still being exponentiated (P / O laid out so that O never overlaps live scores): 1.13 -> 1.06 ms.  What is left is the XU pipe: 2 x 32 rows x 208 exp2 per sub-partition and problem at ~10.5 clocks per
//     MUFU.EX2 warp instruction (tools/ubench/xu_pipe.cu) = 4 400 of the ~5 800 clocks of a problem.
// Tried and measured, not adopted: a single-pass softmax whose reference value is the Cauchy-Schwarz bound |q| max|k| scale (the norms from
// shared memory plus a named barrier cost 30 %); a lazily raised reference with an in-line rescale branch per chunk (drains the XU pipe at
// every chunk: 1.24 ms) or a retry loop around the chunk (spills: 2.1 ms); 3 of 8 exp2 as a degree-3 polynomial on the FMA / ALU pipes
// (FA4 style; 7.5e-5 relative error): the polynomial costs ~14 clocks per warp and element against 10.5 for MUFU.EX2 and the mix runs at
// 1.25 ms - this pass is bound by issue slots and dependent-instruction latency of two co-resident warps as much as by the XU pipe.
#include <stdlib.h>
#include <string.h>

#include "attention.cuh"
#include "common.cuh"
#include "tcgen05.cuh"

namespace sfb {
namespace attn {

using namespace sfb::tc;

namespace {

// phase timestamps of CTA 0 (experiment aid, dbg == nullptr on the product path): dbg[it * 24 + slot] = clock64()
#define SFB_TS(slot) do { if (dbg != nullptr && blockIdx.x == 0 && it < 12 && lane == 0) dbg[it * 24 + (slot)] = clock64(); } while (0)

constexpr int HD = 64;
constexpr int kSoftmaxWarps = 8, kProducerWarps = 3;   // 12 warps: 170 registers per thread are available
constexpr int kThreadsTc = (kSoftmaxWarps + 1 + kProducerWarps) * 32;   // 384
constexpr uint32_t Q_BYTES = 256 * 128;                                 // two 128-row tiles, 128 B per row
constexpr uint32_t KV_ROWS = 208;                                       // keys padded to a multiple of 16
constexpr uint32_t KV_BYTES = KV_ROWS * 128;                            // 26 KB, a multiple of 1024
constexpr uint32_t STAGE_BYTES_TC = Q_BYTES + 2 * KV_BYTES;             // 84 KB
constexpr uint32_t O_STAGE_BYTES = 128 * 128;                           // one tile's bf16 output rows (128 rows x 64 dims), 128B-swizzled
constexpr uint32_t SMEM_TC = 2 * STAGE_BYTES_TC + 2 * O_STAGE_BYTES + 1024;
// TMEM columns of a tile: S fp32 in [0,208).  P (bf16 pairs, 16 columns per 32 keys) of keys 0..127 over the dead S columns [0,64), P of keys
// 128..207 in the spare columns [208,248), O in [64,128): the first 8 of the 13 P V instructions can be issued - and accumulate into O - while
// the softmax is still reading the scores of keys 128.. from columns [128,208)
constexpr uint32_t TILE_COLS = 256, P_COL = 0, PB_COL = 208, O_COL = 64;
constexpr int kKeysA = 128;              // keys covered by the first P V burst
constexpr float kSumLimit = 1.2980742e33f;   // 2^110: single-pass softmax, a row whose sum of 2^(score - reference) reaches this is recomputed

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// (head, inner, outer) of the problems a CTA visits: prob = first, first + stride, ...; advanced with two carries instead of three integer
// divisions by run-time divisors per problem.  The divisions (3 x I2F / MUFU.RCP / F2I, on the XU pipe that one tile's exp2 pass saturates)
// cost the softmax warps 430 - 790 clocks of critical path wherever they were written: the compiler sinks them next to their first use.
struct ProblemIndex {
    int h, i, o, step_h, step_i, step_o, n_heads, n_inner;
    __device__ __forceinline__ ProblemIndex(int first, int stride, int n_heads_, int n_inner_) : n_heads(n_heads_), n_inner(n_inner_) {
        h = first % n_heads, i = (first / n_heads) % n_inner, o = first / (n_heads * n_inner);
        step_h = stride % n_heads, step_i = (stride / n_heads) % n_inner, step_o = stride / (n_heads * n_inner);
    }
    __device__ __forceinline__ void advance() {
        h += step_h;
        if (h >= n_heads) h -= n_heads, ++i;
        i += step_i;
        if (i >= n_inner) i -= n_inner, ++o;
        o += step_o;
    }
};

__device__ __forceinline__ float max3f(float a, float b, float c) {
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
}
__device__ __forceinline__ void softmax_max32(const uint32_t (&r)[32], float &mx) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) mx = max3f(mx, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
}
// 32 scores -> 16 packed bf16 probability pairs, running sum
__device__ __forceinline__ void softmax_exp32(const uint32_t (&r)[32], uint32_t (&pk)[16], float sl2, float mxs, float &sum) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float p0 = ex2f(fmaf(__uint_as_float(r[j]), sl2, -mxs));
        const float p1 = ex2f(fmaf(__uint_as_float(r[j + 1]), sl2, -mxs));
        s0 += p0, s1 += p1;
        pk[j >> 1] = pack_bf16x2(p0, p1);
    }
    sum += s0 + s1;
}

// kFull192: Lk >= 192, i.e. the six 32-column chunks of a score row are never masked (the product shape, Lk = 196): the masked variants of
// the chunk code are not generated, which halves the softmax warps' loop body (89 KB of SASS for the generic kernel)
template <bool kFull192>
__global__ void __launch_bounds__(kThreadsTc, 1)
attn_space_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                     const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1, int tma_out, int o_rows_outer, int o_rows_inner,
                     const Desc d, int n_prob, int use_tma, int q_rows_outer, int q_rows_inner, int kv_rows_outer, int kv_rows_inner, int pipe,
                     uint32_t *__restrict__ flags, long long *dbg) {
    extern __shared__ uint8_t smem_raw[];
    // barriers: full_qk[2] empty_qk[2] | s_ready[2] p_ready[2] o_ready[2] tmem_free[2] | o_stage_full[2] o_stage_free[2] (output staging
    // tiles <-> store-issuing warp) | full_v[2] empty_v[2].  Q / K and V of a stage are handed over separately: Q and K are dead as soon as
    // S of tile 1 has been formed, V only after the last P V - with one barrier per stage the loads of problem i+1 could not start before the
    // middle of problem i and arrived ~700 clocks late every period (clock64 stamps)
    __shared__ __align__(8) uint64_t bars[22];         // ... | p_late_ready[2]
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_g = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };            // Q, K (+ CLS key row, extra query row)
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
    auto fullv_bar = [&](int s) { return bar0 + 8u * (16 + s); };    // V (+ CLS value row)
    auto emptyv_bar = [&](int s) { return bar0 + 8u * (18 + s); };
    auto s_bar = [&](int t) { return bar0 + 8u * (4 + t); };
    auto p_bar = [&](int t) { return bar0 + 8u * (6 + t); };           // P of keys 0..127 stored
    auto pb_bar = [&](int t) { return bar0 + 8u * (20 + t); };         // P of the remaining keys stored
    auto o_bar = [&](int t) { return bar0 + 8u * (8 + t); };
    auto free_bar = [&](int t) { return bar0 + 8u * (10 + t); };
    auto ofull_bar = [&](int t) { return bar0 + 8u * (12 + t); };
    auto ofree_bar = [&](int t) { return bar0 + 8u * (14 + t); };
    const int Lkp = d.Lk + d.has_prefix;            // <= 208; key row d.Lk is the prefix (CLS) row
    const int n_tiles = (d.Lq + 127) / 128;         // 2 for the space attention

    // zero the staging buffers once: rows that are never written (key padding, query padding) must hold finite values
    for (uint32_t i = threadIdx.x; i < 2 * STAGE_BYTES_TC / 16; i += kThreadsTc) reinterpret_cast<uint4 *>(base_g)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 8 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(full_bar(s), use_tma ? 2 : kProducerWarps);   // TMA: expect_tx arrive + the CLS-row warp; else 3 gather warps
            mbar_init(empty_bar(s), 1);
            mbar_init(fullv_bar(s), use_tma ? 2 : kProducerWarps);
            mbar_init(emptyv_bar(s), 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(s_bar(t), 1);
            mbar_init(p_bar(t), 4);
            mbar_init(pb_bar(t), 4);
            mbar_init(o_bar(t), 1);
            mbar_init(free_bar(t), 4);
            mbar_init(ofull_bar(t), 4);
            mbar_init(ofree_bar(t), 1);
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 9 && lane == 0 && use_tma) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
    }
    fence_proxy_async_smem();      // the zero fill must be visible to the tensor cores / TMA (async proxy) too
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp >= 9) {
        // ================================ producers ===================================
        const int pw = warp - 9;                        // 0..2
        if (use_tma && pw >= 2) {
            // warp 11: issues the output TMA stores, so that neither the issue (~500 clocks) nor the wait for the bulk read of the staging
            // tile sits on a softmax warp.  Tiles complete in the order 0, 1 within a problem.
            if (tma_out && lane == 0) {
                int it = 0;
                ProblemIndex pidx(blockIdx.x, gridDim.x, d.n_heads, d.n_inner);
                for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it, pidx.advance()) {
                    const int h = pidx.h, i = pidx.i, o = pidx.o;
                    for (int t = 0; t < n_tiles; ++t) {
                        mbar_wait(ofull_bar(t), it & 1);                 // the tile's 4 warps have written (and fenced) their rows
                        tma_store_2d(t == 0 ? &tm_o0 : &tm_o1, base + 2 * STAGE_BYTES_TC + t * O_STAGE_BYTES, h * HD, o * o_rows_outer + i * o_rows_inner + t * 128);
                        tma_store_commit();
                    }
                    // both stores of this problem have finished READING shared memory -> the tiles may be rewritten (a full period later)
                    tma_store_wait_read();
                    mbar_arrive(ofree_bar(0));
                    mbar_arrive(ofree_bar(1));
                }
                tma_store_wait_all();                                    // shared memory must outlive the bulk stores reading it
            }
        } else {
            int it = 0;
            ProblemIndex pidx(blockIdx.x, gridDim.x, d.n_heads, d.n_inner);
            for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it, pidx.advance()) {
                const int s = it & 1;
                const uint32_t empty_parity = ((it >> 1) & 1) ^ 1u;
                mbar_wait(empty_bar(s), empty_parity);
                const int h = pidx.h, i = pidx.i, o = pidx.o;
                const uint32_t sQ = base + s * STAGE_BYTES_TC, sK = sQ + Q_BYTES, sV = sK + KV_BYTES;
                const int64_t pre_base = o * d.prefix_outer + h * HD;
                if (use_tma) {
                    if (pw == 0) {
                        if (lane == 0) {
                            mbar_arrive_expect_tx(full_bar(s), static_cast<uint32_t>(d.Lq + d.Lk) * 128u);
                            tma_load_2d(sQ, &tm_q, full_bar(s), h * HD, o * q_rows_outer + i * q_rows_inner);
                            tma_load_2d(sK, &tm_k, full_bar(s), h * HD, o * kv_rows_outer + i * kv_rows_inner);
                            mbar_wait(emptyv_bar(s), empty_parity);
                            mbar_arrive_expect_tx(fullv_bar(s), static_cast<uint32_t>(d.Lk) * 128u);
                            tma_load_2d(sV, &tm_v, fullv_bar(s), h * HD, o * kv_rows_outer + i * kv_rows_inner);
                        }
                    } else {   // pw == 1: the CLS key / value row -> row d.Lk of the K / V tiles
                        const int cc = lane & 7;
                        if (d.has_prefix && lane < 8) {
                            const int r = d.Lk;
                            cp_async16(sK + r * 128 + ((cc ^ (r & 7)) << 4), d.kp + pre_base + cc * 8);
                        } else if (d.xq != nullptr && lane >= 16 && lane < 24) {      // fused extra query -> query row Lq
                            const int r = d.Lq;
                            cp_async16(sQ + r * 128 + ((cc ^ (r & 7)) << 4), d.xq + o * d.xq_outer + h * HD + cc * 8);
                        }
                        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_bar(s));
                        mbar_wait(emptyv_bar(s), empty_parity);
                        if (d.has_prefix && lane >= 8 && lane < 16) {
                            const int r = d.Lk;
                            cp_async16(sV + r * 128 + ((cc ^ (r & 7)) << 4), d.vp + pre_base + cc * 8);
                        }
                        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(fullv_bar(s));
                    }
                } else {
                    mbar_wait(emptyv_bar(s), empty_parity);
                    const int ptid = threadIdx.x - 9 * 32;          // 0..95
                    const __nv_bfloat16 *qg = d.q + o * d.q_outer + i * d.q_inner + h * HD;
                    for (int c = ptid; c < d.Lq * 8; c += kProducerWarps * 32) {
                        const int r = c >> 3, cc = c & 7;
                        cp_async16(sQ + r * 128 + ((cc ^ (r & 7)) << 4), qg + static_cast<int64_t>(r) * d.q_row + cc * 8);
                    }
                    if (d.xq != nullptr && ptid < 8) {
                        const int r = d.Lq;
                        cp_async16(sQ + r * 128 + ((ptid ^ (r & 7)) << 4), d.xq + o * d.xq_outer + h * HD + ptid * 8);
                    }
                    const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * HD;
                    for (int c = ptid; c < Lkp * 8; c += kProducerWarps * 32) {
                        const int r = c >> 3, cc = c & 7;
                        const bool pre = r == d.Lk;                 // only reached when has_prefix
                        const int64_t off = (pre ? pre_base : kv_base + static_cast<int64_t>(r) * d.kv_row) + cc * 8;
                        const uint32_t sw = r * 128 + ((cc ^ (r & 7)) << 4);
                        cp_async16(sK + sw, (pre ? d.kp : d.k) + off);
                        cp_async16(sV + sw, (pre ? d.vp : d.v) + off);
                    }
                    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                    fence_proxy_async_smem();               // generic-proxy writes -> visible to tcgen05.mma (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full_bar(s)), mbar_arrive(fullv_bar(s));
                }
            }
        }
    } else if (warp == 8) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_major(128, KV_ROWS, 0, 0);     // S = Q K^T, both K-major
            const uint32_t idesc_o = make_idesc_major(128, HD, 0, 1);          // O = P V, V is MN-major
            // descriptors differ by constants only: a fully unrolled sequence issues in ~56 clocks per tcgen05.mma, a rolled loop that
            // rebuilds them in ~150 (tools/ubench/mma_cost.cu)
            auto issue_s = [&](int t, int it_) {
                const uint32_t sQ = base + (it_ & 1) * STAGE_BYTES_TC, sK = sQ + Q_BYTES;
                mbar_wait(free_bar(t), (it_ & 1) ^ 1u);             // O_t of the previous problem has been read out
                tc_fence_after();
                const uint64_t dq = make_sw128_desc(sQ + t * 16384), dk = make_sw128_desc(sK);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + t * TILE_COLS, dq + 2 * k, dk + 2 * k, idesc_s, static_cast<uint32_t>(k != 0));
                umma_commit(s_bar(t));
            };
            auto issue_pv = [&](int t, int it_) {
                const uint32_t sV = base + (it_ & 1) * STAGE_BYTES_TC + Q_BYTES + KV_BYTES;
                const uint64_t dv = make_sw128_mn_desc(sV, KV_BYTES);
                const uint32_t to = tmem_base + t * TILE_COLS + O_COL, tp = tmem_base + t * TILE_COLS + P_COL, tpb = tmem_base + t * TILE_COLS + PB_COL;
                mbar_wait(p_bar(t), it_ & 1);                       // softmax has written P_t of keys 0..127 into TMEM
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < kKeysA / 16; ++k)                             // 16 keys = 2 KB of V = 128 units of the descriptor's address field
                    umma_bf16_ts(to, tp + k * 8, dv + 128 * k, idesc_o, static_cast<uint32_t>(k != 0));
                mbar_wait(pb_bar(t), it_ & 1);                      // ... and of the remaining keys
                tc_fence_after();
#pragma unroll
                for (int k = kKeysA / 16; k < static_cast<int>(KV_ROWS) / 16; ++k)
                    umma_bf16_ts(to, tpb + (k - kKeysA / 16) * 8, dv + 128 * k, idesc_o, 1u);
                umma_commit(o_bar(t));
            };
            int it = 0;
            if (!pipe) {
                for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it) {
                    mbar_wait(full_bar(it & 1), (it >> 1) & 1);
                    mbar_wait(fullv_bar(it & 1), (it >> 1) & 1);
                    tc_fence_after();
                    SFB_TS(0);
                    issue_s(0, it);
                    SFB_TS(1);
                    issue_s(1, it);
                    SFB_TS(2);
                    issue_pv(0, it);
                    SFB_TS(3);
                    issue_pv(1, it);
                    SFB_TS(4);
                    umma_commit(empty_bar(it & 1));                     // every MMA that reads this stage has been issued
                    umma_commit(emptyv_bar(it & 1));
                }
            } else {
                // The two TMEM slots (tiles) run half a period apart: tile 1 of problem i-1 gets its P V and tile 1 of problem i its S while
                // tile 0 of problem i is in its softmax, and vice versa - each tile's exp2 pass has the XU pipe to itself and the other tile's
                // MMA / barrier latencies hide behind it.  Order per problem:  S(0,i)  PV(1,i-1)  S(1,i)  PV(0,i).  The phase offset is set
                // once, by holding back the first S of tile 1 until tile 0 has finished its first softmax; nothing in the steady state changes it.
                for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it) {
                    mbar_wait(full_bar(it & 1), (it >> 1) & 1);
                    tc_fence_after();
                    SFB_TS(0);
                    issue_s(0, it);
                    SFB_TS(1);
                    if (it > 0) {
                        issue_pv(1, it - 1);
                        umma_commit(emptyv_bar((it - 1) & 1));           // V of problem i-1 is dead
                    } else {
                        const long long t0 = clock64();
                        while (clock64() - t0 < pipe) {}                 // `pipe` = initial stagger in clocks (~ half a period)
                    }
                    SFB_TS(2);
                    issue_s(1, it);
                    umma_commit(empty_bar(it & 1));                      // Q, K of problem i are dead once both S tiles exist
                    SFB_TS(3);
                    mbar_wait(fullv_bar(it & 1), (it >> 1) & 1);
                    tc_fence_after();
                    issue_pv(0, it);
                    SFB_TS(4);
                }
                if (it > 0) {
                    issue_pv(1, it - 1);
                    umma_commit(emptyv_bar((it - 1) & 1));
                }
            }
        }
    } else {
        // ================================ softmax + epilogue ===========================
        const int t = warp >> 2;                                    // row tile
        const int row = t * 128 + (warp & 3) * 32 + lane;           // query row == TMEM lane
        const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + t * TILE_COLS;
        const float sl2 = d.scale * 1.4426950408889634f;
        const bool tile_live = t < n_tiles;
        const int Lq_eff = d.Lq + (d.xq != nullptr ? 1 : 0);        // the fused extra query sits in query row Lq
        const bool rows_live = t * 128 + (warp & 3) * 32 < Lq_eff;  // warp-uniform: a warp whose 32 rows are all padding only keeps the barriers in step
        const bool is_x = d.xq != nullptr && row == d.Lq;
        int it = 0;
        ProblemIndex pidx(blockIdx.x, gridDim.x, d.n_heads, d.n_inner);
        for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it) {
            if (!tile_live) continue;
            const int ph = pidx.h, pi = pidx.i, po = pidx.o;
            // keys this row may see: all Lk (+ prefix); the extra query counts the prefix key in inner problem 0 only
            const int lk = (is_x && pi != 0) ? d.Lk : Lkp;
            if ((warp & 3) == 0) SFB_TS(5 + 8 * t);
            mbar_wait(s_bar(t), it & 1);
            tc_fence_after();
            if ((warp & 3) == 0) SFB_TS(6 + 8 * t);
            float inv = 0.f, mxs_keep = 0.f, sum_keep = 1.f;
            bool redo = false;
            if (rows_live) {
                uint32_t ra[32], rb[32];
                // ---- ONE pass over the scores: p = 2^(s*scale*log2e - ref), row sum, P (bf16 pairs) written over the already consumed S
                // columns.  `ref` is not the row maximum - finding that first is a second sweep over TMEM, 980 of the 6 000 clocks of a
                // problem - but the maximum of the first 32 scores.  Softmax is invariant to the reference as long as nothing overflows,
                // and whether anything came close is visible in the row sum afterwards: sum < 2^110 means every p < 2^110 (bf16 P, the
                // fp32 sum and the fp32 O accumulators are all far from overflow; the sum is >= 1 because the reference is one of the row's
                // own scores), so the row is exact.  A row beyond that (a score more than ~100 log2 units above the first 32; never seen
                // outside the adversarial unit test) is reported in `flags` and recomputed by attn_space_fixup_kernel afterwards.
                float sum = 0.f;
                uint32_t pk[16];
                tmem_ld32(trow, ra);
                tmem_ld_wait_dep(ra);
                float mxs;
                {
                    float m0 = -INFINITY;
                    if (kFull192 || 32 <= d.Lk) softmax_max32(ra, m0);
                    else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (j < lk) m0 = fmaxf(m0, __uint_as_float(ra[j]));
                    }
                    mxs = m0 * sl2;
                }
                if ((warp & 3) == 0) SFB_TS(7 + 8 * t);
#pragma unroll 1          // rolled: the softmax loop body has to stay inside the instruction cache (see the header)
                for (int c = 0; c < 6; c += 2) {
                    tmem_ld32(trow + (c + 1) * 32, rb);
                    if (kFull192 || (c + 1) * 32 <= d.Lk) softmax_exp32(ra, pk, sl2, mxs, sum);
                    else {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float p0 = c * 32 + j < lk ? ex2f(fmaf(__uint_as_float(ra[j]), sl2, -mxs)) : 0.f;
                            const float p1 = c * 32 + j + 1 < lk ? ex2f(fmaf(__uint_as_float(ra[j + 1]), sl2, -mxs)) : 0.f;
                            sum += p0 + p1;
                            pk[j >> 1] = pack_bf16x2(p0, p1);
                        }
                    }
                    tmem_ld_wait_dep(rb);
                    tmem_st16(trow + (c < 4 ? P_COL + c * 16 : PB_COL + (c - 4) * 16), pk);
                    if (c + 2 < 6) tmem_ld32(trow + (c + 2) * 32, ra); else tmem_ld16_into32(trow + 192, ra);
                    if (kFull192 || (c + 2) * 32 <= d.Lk) softmax_exp32(rb, pk, sl2, mxs, sum);
                    else {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float p0 = (c + 1) * 32 + j < lk ? ex2f(fmaf(__uint_as_float(rb[j]), sl2, -mxs)) : 0.f;
                            const float p1 = (c + 1) * 32 + j + 1 < lk ? ex2f(fmaf(__uint_as_float(rb[j + 1]), sl2, -mxs)) : 0.f;
                            sum += p0 + p1;
                            pk[j >> 1] = pack_bf16x2(p0, p1);
                        }
                    }
                    tmem_ld_wait_dep(ra);
                    tmem_st16(trow + (c < 4 ? P_COL + (c + 1) * 16 : PB_COL + (c - 3) * 16), pk);
                    if (c == 2) {            // P of keys 0..127 is complete: the tensor core may start on O while the rest is exponentiated
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(p_bar(t));
                    }
                }
                {
                    uint32_t pt[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float p0 = 192 + j < lk ? ex2f(fmaf(__uint_as_float(ra[j]), sl2, -mxs)) : 0.f;
                        const float p1 = 192 + j + 1 < lk ? ex2f(fmaf(__uint_as_float(ra[j + 1]), sl2, -mxs)) : 0.f;
                        sum += p0 + p1;
                        pt[j >> 1] = pack_bf16x2(p0, p1);
                    }
                    tmem_st8(trow + PB_COL + 32, pt);
                }
                redo = sum >= kSumLimit;         // +inf included; NaN rows (non-finite inputs) are left alone: they come out non-finite either way, as in the reference
                tmem_st_wait();
                inv = 1.0f / sum;
                mxs_keep = mxs, sum_keep = sum;
            }
            tc_fence_before();
            {
                const uint32_t redo_rows = __ballot_sync(0xffffffffu, redo);       // always written: the fix-up kernel needs no memset
                if (lane == 0) flags[static_cast<int64_t>(prob) * 8 + warp] = redo_rows;
            }
            if ((warp & 3) == 0) SFB_TS(8 + 8 * t);
            if (lane == 0) {
                if (!rows_live) mbar_arrive(p_bar(t));              // an all-padding warp skipped the pass (and its mid-pass arrival)
                mbar_arrive(pb_bar(t));
            }
            // ---- while the tensor core forms O_t: next problem's indices, and the staging tile must be free again
            pidx.advance();
            if (tma_out && it > 0) mbar_wait(ofree_bar(t), (it - 1) & 1);    // the previous problem's store has finished reading the staging tile
            // ---- epilogue: O_t / sum -> bf16 -> one 128-byte row per thread
            mbar_wait(o_bar(t), it & 1);
            tc_fence_after();
            if ((warp & 3) == 0) SFB_TS(9 + 8 * t);
            uint32_t o0[32], o1[32];
            if (rows_live) {
                tmem_ld32(trow + O_COL, o0);
                tmem_ld32(trow + O_COL + 32, o1);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if ((warp & 3) == 0) SFB_TS(10 + 8 * t);
            if (lane == 0) mbar_arrive(free_bar(t));                // TMEM columns of tile t may be overwritten by the next problem
            if (is_x) {       // softmax state of the extra query over this problem's keys (sfb_attention_merge_partials combines them)
                float *xp = d.xpartial + ((static_cast<int64_t>(po) * d.n_heads + ph) * d.n_inner + pi) * (HD + 2);
                float2 *xp2 = reinterpret_cast<float2 *>(xp);                 // entries are 66 floats apart: 8-byte aligned
                xp2[0] = make_float2(mxs_keep, sum_keep);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    xp2[1 + k] = make_float2(__uint_as_float(o0[2 * k]) * inv, __uint_as_float(o0[2 * k + 1]) * inv);
                    xp2[17 + k] = make_float2(__uint_as_float(o1[2 * k]) * inv, __uint_as_float(o1[2 * k + 1]) * inv);
                }
            } else if (!tma_out && rows_live && row < d.Lq) {
                const int h = ph, i = pi, o = po;
                uint4 *og = reinterpret_cast<uint4 *>(d.out + o * d.o_outer + i * d.o_inner + static_cast<int64_t>(row) * d.o_row + h * HD);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    og[k] = make_uint4(pack_bf16x2(__uint_as_float(o0[8 * k]) * inv, __uint_as_float(o0[8 * k + 1]) * inv),
                                       pack_bf16x2(__uint_as_float(o0[8 * k + 2]) * inv, __uint_as_float(o0[8 * k + 3]) * inv),
                                       pack_bf16x2(__uint_as_float(o0[8 * k + 4]) * inv, __uint_as_float(o0[8 * k + 5]) * inv),
                                       pack_bf16x2(__uint_as_float(o0[8 * k + 6]) * inv, __uint_as_float(o0[8 * k + 7]) * inv));
                    og[4 + k] = make_uint4(pack_bf16x2(__uint_as_float(o1[8 * k]) * inv, __uint_as_float(o1[8 * k + 1]) * inv),
                                           pack_bf16x2(__uint_as_float(o1[8 * k + 2]) * inv, __uint_as_float(o1[8 * k + 3]) * inv),
                                           pack_bf16x2(__uint_as_float(o1[8 * k + 4]) * inv, __uint_as_float(o1[8 * k + 5]) * inv),
                                           pack_bf16x2(__uint_as_float(o1[8 * k + 6]) * inv, __uint_as_float(o1[8 * k + 7]) * inv));
                }
            }
            if (tma_out) {
                // Output rows leave through ONE TMA store per (problem, tile): a row-per-thread st.global of 128 bytes touches 32 different
                // lines per instruction and kept the softmax warps ~1 800 clocks (clock64 stamps) before they could take the next problem.
                // The tile's 4 warps write their rows into a 128B-swizzled staging tile (conflict-free per quarter-warp) and arrive on an
                // mbarrier; the otherwise idle producer warp 11 hands the tile to the TMA unit.  The box of tile 1 has Lq - 128 rows, so
                // padding rows and the extra-query row never reach memory.
                const int rt = (warp & 3) * 32 + lane;                         // row inside the tile
                if (rows_live) {
                    uint8_t *orow = base_g + 2 * STAGE_BYTES_TC + t * O_STAGE_BYTES + rt * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        *reinterpret_cast<uint4 *>(orow + ((k ^ (rt & 7)) << 4)) =
                            make_uint4(pack_bf16x2(__uint_as_float(o0[8 * k]) * inv, __uint_as_float(o0[8 * k + 1]) * inv),
                                       pack_bf16x2(__uint_as_float(o0[8 * k + 2]) * inv, __uint_as_float(o0[8 * k + 3]) * inv),
                                       pack_bf16x2(__uint_as_float(o0[8 * k + 4]) * inv, __uint_as_float(o0[8 * k + 5]) * inv),
                                       pack_bf16x2(__uint_as_float(o0[8 * k + 6]) * inv, __uint_as_float(o0[8 * k + 7]) * inv));
                        *reinterpret_cast<uint4 *>(orow + (((4 + k) ^ (rt & 7)) << 4)) =
                            make_uint4(pack_bf16x2(__uint_as_float(o1[8 * k]) * inv, __uint_as_float(o1[8 * k + 1]) * inv),
                                       pack_bf16x2(__uint_as_float(o1[8 * k + 2]) * inv, __uint_as_float(o1[8 * k + 3]) * inv),
                                       pack_bf16x2(__uint_as_float(o1[8 * k + 4]) * inv, __uint_as_float(o1[8 * k + 5]) * inv),
                                       pack_bf16x2(__uint_as_float(o1[8 * k + 6]) * inv, __uint_as_float(o1[8 * k + 7]) * inv));
                    }
                }
                if ((warp & 3) == 0) SFB_TS(20 + 2 * t);
                fence_proxy_async_smem();                                      // generic-proxy writes -> visible to the TMA unit
                __syncwarp();
                if (lane == 0) mbar_arrive(ofull_bar(t));
                if ((warp & 3) == 0) SFB_TS(21 + 2 * t);
            }
            if ((warp & 3) == 0) SFB_TS(11 + 8 * t);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// Rows the single-pass softmax reported in `flags` (a score more than ~100 log2 units above the first 32 scores: row sum >= kSumLimit): recomputed from
// the operands in global memory with an online softmax, one warp per row, and written over what the tensor-core kernel stored - the
// output row, or the partial state of the fused extra query (which the merge kernel reads afterwards).  Word w of a problem's 8 flag words
// covers query rows 32 w .. 32 w + 31.  With no row reported (always, outside the unit test) this kernel reads 32 bytes per problem.
__global__ void __launch_bounds__(256) attn_space_fixup_kernel(const Desc d, const uint32_t *__restrict__ flags, int n_prob, int n_words) {
    const int lane = threadIdx.x & 31;
    const float sl2 = d.scale * 1.4426950408889634f;
    for (int prob = blockIdx.x * 8 + (threadIdx.x >> 5); prob < n_prob; prob += gridDim.x * 8) {
        const uint32_t mine = lane < n_words ? flags[static_cast<int64_t>(prob) * 8 + lane] : 0u;
        if (!__any_sync(0xffffffffu, mine != 0u)) continue;
        const int h = prob % d.n_heads, i = (prob / d.n_heads) % d.n_inner, o = prob / (d.n_heads * d.n_inner);
        for (int w = 0; w < n_words; ++w) {
            uint32_t rows = __shfl_sync(0xffffffffu, mine, w);
            while (rows != 0u) {
                const int r = w * 32 + __ffs(rows) - 1;
                rows &= rows - 1;
                const bool is_x = d.xq != nullptr && r == d.Lq;
                if (r > d.Lq || (r == d.Lq && !is_x)) continue;                    // padding rows are never reported; belt and braces
                const __nv_bfloat16 *qp = is_x ? d.xq + o * d.xq_outer + h * HD : d.q + o * d.q_outer + i * d.q_inner + static_cast<int64_t>(r) * d.q_row + h * HD;
                const float2 q = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(qp + 2 * lane));
                const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * HD + 2 * lane;
                const bool with_prefix = d.has_prefix && (!is_x || i == 0);       // the extra query counts the prefix key once, in inner problem 0
                float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f;
                for (int j = 0; j < d.Lk + (with_prefix ? 1 : 0); ++j) {
                    const bool pre = j == d.Lk;
                    const int64_t off = pre ? o * d.prefix_outer + h * HD + 2 * lane : kv_base + static_cast<int64_t>(j) * d.kv_row;
                    const float2 k = unpack_bf16x2(*reinterpret_cast<const uint32_t *>((pre ? d.kp : d.k) + off));
                    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t *>((pre ? d.vp : d.v) + off));
                    const float sc = warp_sum(fmaf(q.x, k.x, q.y * k.y)) * sl2;
                    const float m_new = fmaxf(m, sc);
                    const float corr = exp2f(m - m_new), p = exp2f(sc - m_new);
                    l = fmaf(l, corr, p), a0 = fmaf(a0, corr, p * v.x), a1 = fmaf(a1, corr, p * v.y);
                    m = m_new;
                }
                const float inv = 1.0f / l;
                if (is_x) {
                    float *xp = d.xpartial + ((static_cast<int64_t>(o) * d.n_heads + h) * d.n_inner + i) * (HD + 2);
                    if (lane == 0) xp[0] = m, xp[1] = l;
                    xp[2 + 2 * lane] = a0 * inv, xp[3 + 2 * lane] = a1 * inv;
                } else {
                    __nv_bfloat16 *op = d.out + o * d.o_outer + i * d.o_inner + static_cast<int64_t>(r) * d.o_row + h * HD;
                    *reinterpret_cast<uint32_t *>(op + 2 * lane) = pack_bf16x2(a0 * inv, a1 * inv);
                }
            }
        }
    }
}

}  // namespace

bool tc_supported(const Desc &d) {
    const int Lkp = d.Lk + d.has_prefix;
    // scale > 0: the single-pass softmax compares chunk maxima of the raw scores with its reference value
    return d.Lq > 128 && d.Lq <= 256 && Lkp >= 16 && Lkp <= static_cast<int>(KV_ROWS) && d.scale > 0.f;
}

int launch_tc(const Desc &d, cudaStream_t st) {
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        SFB_CHECK_CUDA(cudaFuncSetAttribute(attn_space_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TC));
        SFB_CHECK_CUDA(cudaFuncSetAttribute(attn_space_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TC));
    }
    const int64_t n_prob = static_cast<int64_t>(d.n_outer) * d.n_inner * d.n_heads;
    SFB_CHECK_ARG(n_prob < (1ll << 31), "sfb_attention: too many problems");
    // TMA staging needs every problem's rows on one regular 2D grid: outer / inner strides must be whole rows
    const bool regular = d.q_row > 0 && d.kv_row > 0 && d.q_outer % d.q_row == 0 && d.q_inner % d.q_row == 0 && d.kv_outer % d.kv_row == 0 &&
                         d.kv_inner % d.kv_row == 0 && d.q_row >= d.n_heads * HD && d.kv_row >= d.n_heads * HD;
    static const bool tma_enabled = !(getenv("SFB_ATTN_TMA") && atoi(getenv("SFB_ATTN_TMA")) == 0);
    CUtensorMap tq, tk, tv;
    memset(&tq, 0, sizeof(tq)), memset(&tk, 0, sizeof(tk)), memset(&tv, 0, sizeof(tv));
    int use_tma = 0, qo = 0, qi = 0, ko = 0, ki = 0;
    if (regular && tma_enabled) {
        qo = static_cast<int>(d.q_outer / d.q_row), qi = static_cast<int>(d.q_inner / d.q_row);
        ko = static_cast<int>(d.kv_outer / d.kv_row), ki = static_cast<int>(d.kv_inner / d.kv_row);
        const int64_t q_rows = static_cast<int64_t>(d.n_outer - 1) * qo + static_cast<int64_t>(d.n_inner - 1) * qi + d.Lq;
        const int64_t kv_rows = static_cast<int64_t>(d.n_outer - 1) * ko + static_cast<int64_t>(d.n_inner - 1) * ki + d.Lk;
        int rc = encode_tmap_bf16_2d(&tq, d.q, q_rows, d.n_heads * HD, d.q_row, d.Lq, HD);
        if (rc == SFB_OK) rc = encode_tmap_bf16_2d(&tk, d.k, kv_rows, d.n_heads * HD, d.kv_row, d.Lk, HD);
        if (rc == SFB_OK) rc = encode_tmap_bf16_2d(&tv, d.v, kv_rows, d.n_heads * HD, d.kv_row, d.Lk, HD);
        if (rc != SFB_OK) return rc;
        use_tma = 1;
    }
    // output through TMA stores when the output rows also lie on one regular 2D grid
    CUtensorMap to0, to1;
    memset(&to0, 0, sizeof(to0)), memset(&to1, 0, sizeof(to1));
    int tma_out = 0, oo = 0, oi = 0;
    static const bool tma_out_enabled = !(getenv("SFB_ATTN_TMA_OUT") && atoi(getenv("SFB_ATTN_TMA_OUT")) == 0);
    if (use_tma && tma_out_enabled && d.o_row > 0 && d.o_outer % d.o_row == 0 && d.o_inner % d.o_row == 0 && d.o_row >= d.n_heads * HD && d.o_row % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.Lq > 128) {
        oo = static_cast<int>(d.o_outer / d.o_row), oi = static_cast<int>(d.o_inner / d.o_row);
        const int64_t o_rows = static_cast<int64_t>(d.n_outer - 1) * oo + static_cast<int64_t>(d.n_inner - 1) * oi + d.Lq;
        int rc = encode_tmap_bf16_2d(&to0, d.out, o_rows, d.n_heads * HD, d.o_row, 128, HD);
        if (rc == SFB_OK) rc = encode_tmap_bf16_2d(&to1, d.out, o_rows, d.n_heads * HD, d.o_row, d.Lq - 128, HD);
        if (rc != SFB_OK) return rc;
        tma_out = 1;
    }
    const unsigned grid = static_cast<unsigned>(n_prob < num_sms() ? n_prob : num_sms());
    const int pipe = getenv("SFB_ATTN_PIPE") ? atoi(getenv("SFB_ATTN_PIPE")) : 3000;               // initial stagger in clocks; 0 = lock-step issue order (A/B aid)
    long long *dbg = reinterpret_cast<long long *>(getenv("SFB_ATTN_DBG_PTR") ? strtoull(getenv("SFB_ATTN_DBG_PTR"), nullptr, 0) : 0ull);
    // per-launch flag words of the single-pass softmax (8 per problem), stream-ordered so that concurrent launches on other streams cannot
    // share them; the pool keeps the block, so this is a free-list pop after the first call
    static PerDeviceOnce pool_once;
    if (pool_once.first()) {
        int dev = 0;
        cudaMemPool_t pool;
        uint64_t keep = UINT64_MAX;
        SFB_CHECK_CUDA(cudaGetDevice(&dev));
        SFB_CHECK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        SFB_CHECK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    uint32_t *flags = nullptr;
    SFB_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&flags), static_cast<size_t>(n_prob) * 8 * sizeof(uint32_t), st));
    if (d.Lk >= 192)
        attn_space_tc_kernel<true><<<grid, kThreadsTc, SMEM_TC, st>>>(tq, tk, tv, to0, to1, tma_out, oo, oi, d, static_cast<int>(n_prob), use_tma, qo, qi, ko, ki, pipe, flags, dbg);
    else
        attn_space_tc_kernel<false><<<grid, kThreadsTc, SMEM_TC, st>>>(tq, tk, tv, to0, to1, tma_out, oo, oi, d, static_cast<int>(n_prob), use_tma, qo, qi, ko, ki, pipe, flags, dbg);
    cudaError_t launch_err = cudaGetLastError();
    if (launch_err == cudaSuccess) {
        const int64_t want = (n_prob + 7) / 8;
        const unsigned fix_grid = static_cast<unsigned>(want < 4 * num_sms() ? want : 4 * num_sms());
        attn_space_fixup_kernel<<<fix_grid, 256, 0, st>>>(d, flags, static_cast<int>(n_prob), 4 * ((d.Lq + 127) / 128));
        launch_err = cudaGetLastError();
    }
    cudaFreeAsync(flags, st);
    SFB_CHECK_CUDA(launch_err);
    return SFB_OK;
}

}  // namespace attn
}  // namespace sfb
