// K7 on the 5th-generation tensor cores: Motionformer space attention (196 queries x 197 keys, hd 64; vit_helper.py:100-158 with
// einops_to '(b f) n d') as tcgen05.mma with S, P and O resident in TMEM.
//
// One persistent CTA per SM (416 threads) loops over (segment, frame, head) problems:
//   warps 9-12  producers   cp.async 16-byte gathers of the Q / K / V rows (CLS prefix key first) into a 2-stage shared-memory
//                           ring, written directly in the 128B-swizzled K-major layout the UMMA descriptors expect;
//                           cp.async.wait_group + fence.proxy.async + mbarrier arrive publishes a stage
//   warp 8      MMA issuer  S_t = Q_t K^T   (2 row tiles t of 128 queries; 4 x tcgen05.mma M128 N208 K16, operands from smem)
//                           O_t = P_t V     (13 x tcgen05.mma M128 N64 K16, A = P_t read from TMEM, B = V as an MN-major operand:
//                           the [key][64 dims] rows are used as they are, no transpose)
//   warps 0-7   softmax     thread = query row (TMEM lane).  Two passes over the fp32 scores with tcgen05.ld (row max, then
//                           exp2 / row sum), P written back over S as packed bf16 with tcgen05.st, then the O epilogue:
//                           tcgen05.ld, 1/sum, bf16, one 128-byte row store per thread.
// TMEM (512 columns): tile t owns columns [256t, 256t+208): S fp32 there; P (bf16 pairs) re-uses columns [0,104) of the same
// range as the softmax consumes S; O accumulates in columns [128,192) once S is dead.
// Nothing touches HBM between the qkv activations and the attention output.
#include "attention.cuh"
#include "common.cuh"
#include "tcgen05.cuh"

namespace sfb {
namespace attn {

using namespace sfb::tc;

namespace {

constexpr int HD = 64;
constexpr int kSoftmaxWarps = 8, kProducerWarps = 4;
constexpr int kThreadsTc = (kSoftmaxWarps + 1 + kProducerWarps) * 32;   // 416
constexpr uint32_t Q_BYTES = 256 * 128;                                 // two 128-row tiles, 128 B per row
constexpr uint32_t KV_ROWS = 208;                                       // keys padded to a multiple of 16
constexpr uint32_t KV_BYTES = KV_ROWS * 128;                            // 26 KB, a multiple of 1024
constexpr uint32_t STAGE_BYTES_TC = Q_BYTES + 2 * KV_BYTES;             // 84 KB
constexpr uint32_t SMEM_TC = 2 * STAGE_BYTES_TC + 1024;
constexpr uint32_t TILE_COLS = 256, P_COL = 0, O_COL = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kThreadsTc, 1) attn_space_tc_kernel(const Desc d, int n_prob) {
    extern __shared__ uint8_t smem_raw[];
    // barriers: full[2] empty[2] | s_ready[2] p_ready[2] o_ready[2] tmem_free[2]
    __shared__ __align__(8) uint64_t bars[12];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_g = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
    auto s_bar = [&](int t) { return bar0 + 8u * (4 + t); };
    auto p_bar = [&](int t) { return bar0 + 8u * (6 + t); };
    auto o_bar = [&](int t) { return bar0 + 8u * (8 + t); };
    auto free_bar = [&](int t) { return bar0 + 8u * (10 + t); };
    const int Lkp = d.Lk + d.has_prefix;            // <= 208
    const int n_tiles = (d.Lq + 127) / 128;         // 2 for the space attention

    // zero the staging buffers once: rows that cp.async never writes (key padding, query padding) must hold finite values
    for (uint32_t i = threadIdx.x; i < 2 * STAGE_BYTES_TC / 16; i += kThreadsTc) reinterpret_cast<uint4 *>(base_g)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 8 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(full_bar(s), kProducerWarps);
            mbar_init(empty_bar(s), 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(s_bar(t), 1);
            mbar_init(p_bar(t), 4);
            mbar_init(o_bar(t), 1);
            mbar_init(free_bar(t), 4);
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();      // the zero fill must be visible to the tensor cores (async proxy) too
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp >= 9) {
        // ================================ producers ===================================
        const int ptid = threadIdx.x - 9 * 32;          // 0..127
        int it = 0;
        for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it) {
            const int s = it & 1;
            mbar_wait(empty_bar(s), ((it >> 1) & 1) ^ 1u);
            const int h = prob % d.n_heads;
            const int i = (prob / d.n_heads) % d.n_inner;
            const int o = prob / (d.n_heads * d.n_inner);
            const uint32_t sQ = base + s * STAGE_BYTES_TC, sK = sQ + Q_BYTES, sV = sK + KV_BYTES;
            const __nv_bfloat16 *qg = d.q + o * d.q_outer + i * d.q_inner + h * HD;
            for (int c = ptid; c < d.Lq * 8; c += 128) {
                const int r = c >> 3, cc = c & 7;
                cp_async16(sQ + r * 128 + ((cc ^ (r & 7)) << 4), qg + static_cast<int64_t>(r) * d.q_row + cc * 8);
            }
            const int64_t kv_base = o * d.kv_outer + i * d.kv_inner + h * HD;
            const int64_t pre_base = o * d.prefix_outer + h * HD;
            for (int c = ptid; c < Lkp * 8; c += 128) {
                const int r = c >> 3, cc = c & 7;
                const bool pre = d.has_prefix && r == 0;
                const int64_t off = (pre ? pre_base : kv_base + static_cast<int64_t>(r - d.has_prefix) * d.kv_row) + cc * 8;
                const uint32_t sw = r * 128 + ((cc ^ (r & 7)) << 4);
                cp_async16(sK + sw, (pre ? d.kp : d.k) + off);
                cp_async16(sV + sw, (pre ? d.vp : d.v) + off);
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            fence_proxy_async_smem();               // generic-proxy writes -> visible to tcgen05.mma (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
        }
    } else if (warp == 8) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_major(128, KV_ROWS, 0, 0);     // S = Q K^T, both K-major
            const uint32_t idesc_o = make_idesc_major(128, HD, 0, 1);          // O = P V, V is MN-major
            int it = 0;
            for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it) {
                const int s = it & 1;
                const uint32_t sQ = base + s * STAGE_BYTES_TC, sK = sQ + Q_BYTES, sV = sK + KV_BYTES;
                mbar_wait(full_bar(s), (it >> 1) & 1);
                tc_fence_after();
                for (int t = 0; t < n_tiles; ++t) {
                    mbar_wait(free_bar(t), (it & 1) ^ 1u);          // O_t of the previous problem has been read out
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(tmem_base + t * TILE_COLS, make_sw128_desc(sQ + t * 16384 + k * 32), make_sw128_desc(sK + k * 32), idesc_s,
                                  static_cast<uint32_t>(k != 0));
                    umma_commit(s_bar(t));
                }
                for (int t = 0; t < n_tiles; ++t) {
                    mbar_wait(p_bar(t), it & 1);                    // softmax has written P_t into TMEM
                    tc_fence_after();
#pragma unroll 1
                    for (int k = 0; k < static_cast<int>(KV_ROWS) / 16; ++k)
                        umma_bf16_ts(tmem_base + t * TILE_COLS + O_COL, tmem_base + t * TILE_COLS + P_COL + k * 8,
                                     make_sw128_mn_desc(sV + k * 2048, KV_BYTES), idesc_o, static_cast<uint32_t>(k != 0));
                    umma_commit(o_bar(t));
                }
                umma_commit(empty_bar(s));                          // every MMA that reads this stage has been issued
            }
        }
    } else {
        // ================================ softmax + epilogue ===========================
        const int t = warp >> 2;                                    // row tile
        const int row = t * 128 + (warp & 3) * 32 + lane;           // query row == TMEM lane
        const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + t * TILE_COLS;
        const float sl2 = d.scale * 1.4426950408889634f;
        const bool tile_live = t < n_tiles;
        int it = 0;
        for (int prob = blockIdx.x; prob < n_prob; prob += gridDim.x, ++it) {
            if (!tile_live) continue;
            mbar_wait(s_bar(t), it & 1);
            tc_fence_after();
            // pass 1: row maximum over the live keys
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < static_cast<int>(KV_ROWS) / 16; ++c) {
                uint32_t r[16];
                tmem_ld16(trow + c * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c * 16 + j < Lkp) mx = fmaxf(mx, __uint_as_float(r[j]));
            }
            const float mxs = mx * sl2;
            // pass 2: p = 2^(s*scale*log2e - max), row sum, P (bf16 pairs) written over the consumed S columns
            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < static_cast<int>(KV_ROWS) / 16; ++c) {
                uint32_t r[16], pk[8];
                tmem_ld16(trow + c * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float p0 = c * 16 + j < Lkp ? ex2f(fmaf(__uint_as_float(r[j]), sl2, -mxs)) : 0.f;
                    const float p1 = c * 16 + j + 1 < Lkp ? ex2f(fmaf(__uint_as_float(r[j + 1]), sl2, -mxs)) : 0.f;
                    sum += p0 + p1;
                    pk[j >> 1] = pack_bf16x2(p0, p1);
                }
                tmem_st8(trow + P_COL + c * 8, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_bar(t));
            const float inv = 1.0f / sum;
            // epilogue: O_t / sum -> bf16 -> one 128-byte row per thread
            mbar_wait(o_bar(t), it & 1);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld32(trow + O_COL, o0);
            tmem_ld32(trow + O_COL + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(free_bar(t));                // TMEM columns of tile t may be overwritten by the next problem
            if (row < d.Lq) {
                const int h = prob % d.n_heads;
                const int i = (prob / d.n_heads) % d.n_inner;
                const int o = prob / (d.n_heads * d.n_inner);
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
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

bool tc_supported(const Desc &d) {
    const int Lkp = d.Lk + d.has_prefix;
    return d.Lq > 128 && d.Lq <= 256 && Lkp >= 16 && Lkp <= static_cast<int>(KV_ROWS);
}

int launch_tc(const Desc &d, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        SFB_CHECK_CUDA(cudaFuncSetAttribute(attn_space_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TC));
        attr_set = true;
    }
    const int64_t n_prob = static_cast<int64_t>(d.n_outer) * d.n_inner * d.n_heads;
    SFB_CHECK_ARG(n_prob < (1ll << 31), "sfb_attention: too many problems");
    const unsigned grid = static_cast<unsigned>(n_prob < num_sms() ? n_prob : num_sms());
    attn_space_tc_kernel<<<grid, kThreadsTc, SMEM_TC, st>>>(d, static_cast<int>(n_prob));
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

}  // namespace attn
}  // namespace sfb
