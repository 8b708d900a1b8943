// K5: bf16 GEMM with fused bias / exact-GELU / fp32-residual epilogue for every nn.Linear on the path.
//
//   out[M,N] = epi( A[M,K] * W[N,K]^T + bias )          A, W bf16 K-major; fp32 accumulation
//
// sm_100a design (persistent CTAs, one per SM, 576 threads, warp-specialised).  Two instantiations of one template:
//   CG = 2 (product path)  CTA PAIRS (cluster of 2) compute 256 x 256 tiles with tcgen05.mma.cta_group::2: each CTA stages its own
//                          128 rows of A and HALF of the W tile (128 rows), so a pair moves 2/3 of the bytes per FLOP that two
//                          independent 128 x 256 CTAs would.  Measured on B200: the single-CTA kernel saturates L2->SM bandwidth
//                          (~6.4 KB/clk chip-wide) at ~50 % tensor-pipe activity, which is what makes the pairing pay.
//                          5-stage ring of 32 KB stages; the pair leader issues the MMAs; commits are multicast to both CTAs.
//   CG = 1                 single-CTA 128 x 256 tiles, 3-stage ring of 48 KB stages (small M, and kept for A/B comparison).
// Roles:
//   warp 0      TMA producer   cp.async.bulk.tensor 2D loads of a 128x64 A box and a (256 / CG)x64 W box (128B swizzle)
//                              into the shared-memory ring, completion on the (leader's) "full" mbarriers
//   warp 1      MMA issuer     one thread issues 4 x tcgen05.mma.kind::f16 (M 128*CG, N256, K16) per stage,
//                              accumulating in TMEM; tcgen05.commit releases the stage ("empty") and, after the
//                              last k-block, publishes the accumulator ("tmem_full")
//   warps 2-17  epilogue       tcgen05.ld 32x32b.x32 (TMEM -> registers), bias / GELU, a swizzled per-warp shared-memory tile in the
//                              TMA box layout, TMA store (bf16 outputs).  fp32 outputs (EPI_F32_TMA, 8 epilogue warps): the fp32
//                              residual box comes in by TMA load, the accumulator is added in place, the tile leaves by TMA store -
//                              no per-thread global access (the older per-thread fp32 epilogues remain for broadcast residual rows
//                              and as the A/B baseline);
//                              the 512 TMEM columns hold TWO 128x256 fp32 accumulators so the epilogue of tile i
//                              overlaps the main loop of tile i+1
// LayerNorm fusion (the 36 LayerNorm passes of the Motionformer blocks disappear; vit_helper.py:366-375):
//   SFB_GEMM_EMIT_LN  a residual GEMM (proj / fc2) also writes a bf16 copy xb of its fp32 output rows and, per row and 64-column group,
//                     (sum, sum of squares) of the fp32 values: the LayerNorm statistics of the NEXT norm, for free in the epilogue
//   SFB_GEMM_LN_FOLD  the consumer (qkv / fc1) multiplies the UN-normalised xb by weights with gamma folded in and applies
//                         out = rstd_row (acc - mean_row colsum_j) + bias'_j        colsum_j = sum_k bf16(gamma_k W_jk),  bias' = b + W beta
//                     in its epilogue - algebraically LayerNorm(x) W^T + b with the same bf16 operand rounding budget as the unfused path
// Epilogue experiments measured on the B200 in round 2 and NOT adopted (tools/microbench.py, 512 segments): staging bias / column sums in
// shared memory (needs a 4-stage ring: proj+res 1.27 -> 1.21 ms but fc2 2.90 -> 3.14 ms and the LN_FOLD variants slower), and a single-MUFU
// GELU (Abramowitz-Stegun 7.1.28 instead of 7.1.26: fc1 3.29 vs 3.29 ms - the fc1 epilogue is not XU-bound).
// Tiles are walked n-fastest so the CTAs that share an A row-block run at the same time and hit it in L2;
// W (<= 4.7 MB) is L2-resident throughout.  M / N / K tails are handled by TMA zero fill + epilogue masking.
//
// Reference ops replaced: nn.Linear at vit_helper.py:105,156,393-396; modeling_ast.py:149-152,200,258,272;
// modules/transformer.py:62-64,74,87-92; nn.TransformerEncoderLayer linears (motionformer.py:329); sync_model.py:55-56.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace sfb {
namespace gemm {

using namespace sfb::tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kAccStages = 2;
constexpr int kEpiWarps = 16;  // 4 per TMEM lane quarter: the epilogue is latency-bound per warp, so it wants warps, not ILP
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kEpiWarpsTma = 8;                                  // EPI_F32_TMA: 2 per lane quarter, 8 KB of staging each (no per-thread global access)
constexpr int kThreadsTma = 64 + kEpiWarpsTma * 32;
constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;               // 16 KB: this CTA's 128 rows of A
constexpr uint32_t EPI_WARP_BYTES = 4096;                        // 32 rows x 128 B transpose buffer per epilogue warp
constexpr uint32_t TMEM_COLS = 512;
constexpr int kMaxStages = 6;

template <int CG>
struct Cfg {
    static constexpr int kStages = CG == 2 ? 5 : 3;
    static constexpr int B_ROWS = BLOCK_N / CG;                   // W rows staged by one CTA
    static constexpr uint32_t B_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;    // 32 KB (pair) / 48 KB (single)
    static constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + kEpiWarps * EPI_WARP_BYTES + 1024;  // + 1024-byte alignment slack
};

// experiment switches (env SFB_GEMM_DBG, measurement aid only: results are wrong when set): 1 = no global stores in the epilogue,
// 2 = every TMA load fetches tile (0, 0) (operands always L2-resident), 4 = no residual loads
constexpr int DBG_NOSTORE = 1 << 16, DBG_SAMETILE = 1 << 17, DBG_NORES = 1 << 18;

struct EpiParams {
    const float *bias;
    const float *residual;
    int64_t ldr;
    void *out;
    int64_t ldo;
    int M, N, K;
    int flags;
    int num_m_blocks, num_n_blocks;
    int m_fastest;   // tile walk order (experiment switch SFB_GEMM_ORDER=1); default n-fastest
    // LayerNorm fusion
    const float *ln_stats;    // LN_FOLD: [M][ln_parts][2] partial (sum, sum of squares) of the rows of A over its K columns
    const float *ln_colsum;   // LN_FOLD: [N] column sums of the folded bf16 weights
    int ln_parts;
    float ln_eps;
    float *emit_stats;        // EMIT_LN: [M][N / 64][2]
    __nv_bfloat16 *emit_bf16; // EMIT_LN: bf16 copy of the output rows, row stride ld_emit
    int64_t ld_emit;
    int tma_store;            // bf16 output tiles leave through cp.async.bulk.tensor (UTMASTG) instead of per-thread st.global
};

// ------------------------------------------------------------------------------------------ cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion may be signalled on the PEER CTA's mbarrier (pair leader's "full" barrier)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap *m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit of all prior MMAs of the pair, arriving on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
                 : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the bit that distinguishes the two CTAs of a pair in a shared-window address

// ------------------------------------------------------------------------------------------------ the kernel
// CG = CTAs per cluster (1, or a cta_group::2 pair).  EPI selects the epilogue at compile time (each variant gets its own register
// allocation under the 96-register cap of a 576-thread CTA):
//   EPI_BF16     bf16 output (bias, optional GELU)                 EPI_BF16_LN  the same behind a folded LayerNorm (SFB_GEMM_LN_FOLD)
//   EPI_F32      fp32 output (bias, optional GELU / residual)      EPI_F32_LN   the same + bf16 copy and row statistics (SFB_GEMM_EMIT_LN)
// (A third cluster shape - two pairs sharing every W tile through TMA multicast, 48 instead of 64 bytes of L2 reads per MMA clock - was
// built and measured 3-8 % slower in round 1: multicast removes L2 reads, not bytes entering the SM, and 4-CTA clusters strand SMs.)
//   EPI_F32_TMA  fp32 output (bias, optional GELU / residual, optional EMIT_LN) with NO per-thread global access: 8 epilogue warps, each
//                with a 4 KB fp32 tile (32 rows x 32 columns, 128B-swizzled = the TMA box layout) and a 4 KB bf16 tile.  The residual box
//                arrives by TMA load (its lines were pulled into L2 by the producer a tile earlier), every thread adds its accumulator
//                row segment IN PLACE (row-per-thread mapping: the LayerNorm statistics of EMIT_LN are plain per-thread sums, no
//                shuffles), and the tile leaves by TMA store; the bf16 copy of EMIT_LN leaves the same way.
enum { EPI_BF16 = 0, EPI_BF16_LN = 1, EPI_F32 = 2, EPI_F32_LN = 3, EPI_F32_TMA = 4 };

template <int CG, int EPI>
__global__ void __launch_bounds__(EPI == EPI_F32_TMA ? kThreadsTma : kThreads, 1)   // 18 warps are allocated as 20 (granularity 4): 96 registers per thread
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                         const __grid_constant__ CUtensorMap tmap_emit, const EpiParams p) {
    using C = Cfg<CG>;
    constexpr int CL = CG;
    constexpr int kStages = C::kStages;
    constexpr uint32_t STAGE_BYTES = C::STAGE_BYTES;
    constexpr int kEpiW = EPI == EPI_F32_TMA ? kEpiWarpsTma : kEpiWarps;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 2 * kAccStages + 2 * kEpiWarpsTma];    // ... + two "residual box landed" barriers per TMA-epilogue warp
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;     // rank in the cluster = rank in the pair; 0 is the MMA leader
    const uint32_t rank = crank;
    const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + kAccStages + a); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        if (p.tma_store) tma_prefetch_desc(&tmap_out);
        if (EPI == EPI_F32_TMA) {
            tma_prefetch_desc(&tmap_out), tma_prefetch_desc(&tmap_res), tma_prefetch_desc(&tmap_emit);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);     // the (leader's) producer arrive.expect_tx; TMA of both CTAs completes the bytes
            mbar_init(empty_bar(s), 1);    // the tcgen05.commit of the (pair's) MMA issuer
        }
        for (int a = 0; a < kAccStages; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), kEpiW * CG);   // epilogue warps of every CTA of the pair arrive on the leader's barrier
        }
        if (EPI == EPI_F32_TMA) {
            for (int w = 0; w < 2 * kEpiWarpsTma; ++w) mbar_init(bar_base + 8u * (2 * kMaxStages + 2 * kAccStages + w), 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {  // one full warp per CTA allocates all 512 TMEM columns (1 CTA per SM) and later frees them
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();       // barriers initialised + TMEM allocated in every CTA of the cluster
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int num_tiles = p.num_m_blocks * p.num_n_blocks;       // tiles of (128 * CG) x 256
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int first_tile = blockIdx.x / CL, tile_step = gridDim.x / CL;     // a tile is (128 * CG) rows x 256 columns

    if (warp == 0) {
        // ================================ TMA producer ================================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            int m0 = (p.m_fastest ? tile % p.num_m_blocks : tile / p.num_n_blocks) * (BLOCK_M * CL) + crank * BLOCK_M;
            int n0 = (p.m_fastest ? tile / p.num_m_blocks : tile % p.num_n_blocks) * BLOCK_N + rank * C::B_ROWS * (CG - 1);
            if (lane == 0) {
                // a box that lies completely outside the matrix (second CTA of a pair on a ragged edge) loads rows 0.. instead:
                // its products only reach accumulator rows / columns that the epilogue masks
                if (m0 >= p.M) m0 = 0;
                if (n0 >= p.N) n0 = 0;
                if (p.flags & DBG_SAMETILE) m0 = crank * BLOCK_M, n0 = rank * C::B_ROWS * (CG - 1);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = tiles_base + stage * STAGE_BYTES;
                    if (CG == 2) {
                        const uint32_t fb = full_bar(stage) & kPeerBitMask;                   // the pair leader's barrier
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);   // bytes landing in both CTAs of the pair
                        tma_load_2d_cg2(sa, &tmap_a, fb, kb * BLOCK_K, m0);
                        tma_load_2d_cg2(sa + A_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
                    } else {
                        mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
                        tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BLOCK_K, m0);
                        tma_load_2d(sa + A_BYTES, &tmap_w, full_bar(stage), kb * BLOCK_K, n0);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (pair leader only) ================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M * CG, BLOCK_N);
            const uint64_t desc0 = make_sw128_desc(tiles_base);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // every epilogue warp (of both CTAs) has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    // descriptors of a stage differ from `desc0` by constants (address field = bytes >> 4): the issue sequence is four
                    // tcgen05.mma with immediate-offset operands - a sequence that rebuilds descriptors from addresses issues ~3x slower
                    // (tools/ubench/mma_cost.cu: ~150 vs ~56 clocks per instruction, against 128 clocks of tensor-pipe work each)
                    const uint64_t da = desc0 + static_cast<uint64_t>(stage * (STAGE_BYTES >> 4)), db = da + (A_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint32_t acc_flag = k != 0 ? 1u : static_cast<uint32_t>(kb != 0);
                        if (CG == 2) umma_bf16_cg2(tmem_d, da + 2 * k, db + 2 * k, idesc, acc_flag);
                        else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, acc_flag);
                    }
                    if (CG == 2) umma_commit_cg2(empty_bar(stage), static_cast<uint16_t>(3u)); else umma_commit(empty_bar(stage));   // smem stage reusable
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (CG == 2) umma_commit_cg2(tfull_bar(acc), static_cast<uint16_t>(3u)); else umma_commit(tfull_bar(acc));   // accumulator complete
                if (++acc == kAccStages) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (EPI == EPI_F32_TMA) {
        // ================================ epilogue, fp32 tiles through TMA both ways ===
        // warp (q, half) owns accumulator rows [32q, 32q+32) x columns [128 half, 128 half + 128): four 32-column chunks per tile.
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;
        const bool gelu = (p.flags & SFB_GEMM_GELU) != 0;
        const bool has_res = (p.flags & SFB_GEMM_RESIDUAL) != 0 && !(p.flags & DBG_NORES);
        const bool emit = (p.flags & SFB_GEMM_EMIT_LN) != 0;
        const bool do_store = !(p.flags & DBG_NOSTORE);
        // per warp: two 4 KB tiles (32 rows x 128 B, 128B-swizzled = the TMA box layout).  Without EMIT_LN both hold fp32 chunks alternately
        // (the residual box of chunk g+2 is requested as soon as the store of chunk g has left its tile: two boxes in flight per warp);
        // with EMIT_LN the second tile collects the bf16 copy of two chunks (64 columns) and the fp32 chunks use the first tile only.
        uint8_t *xt = smem_raw + (tiles_base - smem_u32(smem_raw)) + kStages * STAGE_BYTES + (warp - 2) * 8192;
        const uint32_t xt_s = smem_u32(xt), bt_s = xt_s + 4096;
        uint8_t *brow = xt + 4096 + lane * 128;
        const int sw = lane & 7;                                       // 16-byte chunk k of row r lives at chunk k ^ (r & 7): SWIZZLE_128B
        const uint32_t rbar0 = bar_base + 8u * (2 * kMaxStages + 2 * kAccStages + 2 * (warp - 2));
        const int depth = emit ? 1 : 2;                                // residual boxes in flight
        uint32_t rphase = 0;                                           // bit b: parity of tile b's barrier
        int acc = 0, g = 0;                                            // g: valid chunks processed so far (its parity selects the tile)
        uint32_t acc_phase = 0;
        const uint32_t tempty_leader0 = CG == 2 ? mapa_u32(tempty_bar(0), 0u) : tempty_bar(0);
        auto tile_origin = [&](int tile, int &m0w, int &n0w) {
            m0w = (p.m_fastest ? tile % p.num_m_blocks : tile / p.num_n_blocks) * (BLOCK_M * CL) + crank * BLOCK_M + q * 32;
            n0w = (p.m_fastest ? tile / p.num_m_blocks : tile % p.num_n_blocks) * BLOCK_N + half * 128;
        };
        // request cursor: walks the chunks this warp will process (tiles in order, chunks 0..3, skipping rows >= M / columns >= N), `depth` ahead
        int q_tile = first_tile, q_c = -1, q_m = 0, q_col = 0, q_n = 0;      // q_n: boxes requested so far
        bool q_valid = has_res;
        auto request_next = [&]() {                                   // warp-uniform bookkeeping, lane 0 issues
            while (q_valid) {
                if (q_c < 3) ++q_c; else q_tile += tile_step, q_c = 0;
                if (q_tile >= num_tiles) { q_valid = false; break; }
                int m, n;
                tile_origin(q_tile, m, n);
                if (m < p.M && n + q_c * 32 < p.N) { q_m = m, q_col = n + q_c * 32; break; }
                q_c = 3;                                              // nothing (more) to do in this tile
            }
            if (q_valid) {
                const int bsel = emit ? 0 : (q_n & 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(rbar0 + 8u * bsel, 4096u);
                    tma_load_2d(xt_s + 4096u * bsel, &tmap_res, rbar0 + 8u * bsel, q_col, q_m);
                }
                ++q_n;
            }
        };
        for (int i = 0; i < depth; ++i) request_next();
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            int m0w, n0w;
            tile_origin(tile, m0w, n0w);
            const bool live = m0w < p.M;                               // warp-uniform
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N + half * 128);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            uint32_t r[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(tbase + c * 32, r[c]);
            tmem_ld_wait();
            tc_fence_before();                                         // the accumulator goes back to the MMA issuer before any of the epilogue math
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc); else mbar_arrive(tempty_bar(acc));
            }
            float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};              // EMIT_LN: (sum, sum of squares) of this thread's row per 64-column group
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = n0w + c * 32;
                if (live && col < p.N) {                               // warp-uniform
                    const int bsel = emit ? 0 : (g & 1);
                    uint8_t *xrow = xt + 4096 * bsel + lane * 128;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c][j]);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col + j < p.N) {
                                const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + col + j));
                                v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
                            }
                        }
                    }
                    if (gelu) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) gelu_erf_fast2(v[j], v[j + 1]);
                    }
                    if (has_res) {
                        mbar_wait(rbar0 + 8u * bsel, (rphase >> bsel) & 1u);
                        rphase ^= 1u << bsel;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 rv = *reinterpret_cast<const float4 *>(xrow + ((k ^ sw) << 4));
                            v[4 * k] += rv.x, v[4 * k + 1] += rv.y, v[4 * k + 2] += rv.z, v[4 * k + 3] += rv.w;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        *reinterpret_cast<float4 *>(xrow + ((k ^ sw) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    if (emit) {
                        float a1 = 0.f, a2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) a1 += v[j], a2 = fmaf(v[j], v[j], a2);
                        s1[c >> 1] += a1, s2[c >> 1] += a2;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            *reinterpret_cast<uint4 *>(brow + ((((c & 1) * 4 + k) ^ sw) << 4)) =
                                make_uint4(pack_bf16x2(v[8 * k], v[8 * k + 1]), pack_bf16x2(v[8 * k + 2], v[8 * k + 3]),
                                           pack_bf16x2(v[8 * k + 4], v[8 * k + 5]), pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
                    }
                    fence_proxy_async_smem();                          // generic-proxy writes -> visible to the TMA unit
                    __syncwarp();
                    if (lane == 0 && do_store) {
                        tma_store_2d(&tmap_out, xt_s + 4096u * bsel, col, m0w);          // rows >= M / columns >= N are clipped by the unit
                        if (emit && (c & 1)) tma_store_2d(&tmap_emit, bt_s, n0w + (c >> 1) * 64, m0w);
                        tma_store_commit();
                        tma_store_wait_read();                         // this tile may be overwritten: by the next residual box, or by chunk g + depth
                    }
                    __syncwarp();
                    request_next();                                    // the box of chunk g + depth goes into the tile that has just been stored
                    ++g;
                }
            }
            if (emit && live) {
                const int64_t grow = m0w + lane;
                if (grow < p.M && do_store) {
#pragma unroll
                    for (int g2 = 0; g2 < 2; ++g2) {
                        const int gcol = n0w + g2 * 64;
                        if (gcol < p.N) *reinterpret_cast<float2 *>(p.emit_stats + (grow * (p.N >> 6) + (gcol >> 6)) * 2) = make_float2(s1[g2], s2[g2]);
                    }
                }
            }
            if (++acc == kAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        if (lane == 0) tma_store_wait_all();                           // shared memory must outlive the bulk stores that read it
    } else {
        // ================================ epilogue ====================================
        // TMEM -> registers (one accumulator row per thread) -> bias / GELU -> per-warp 4 KB shared-memory transpose buffer
        // (16-byte chunks XOR-swizzled by row, conflict-free both ways) -> row-contiguous global access: every load/store
        // instruction covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes.
        const int q = warp & 3;               // TMEM lane quarter this warp may read: lanes [32q, 32q+32)
        const int quarter = (warp - 2) >> 2;  // which 64 of the 256 accumulator columns
        const bool gelu = (p.flags & SFB_GEMM_GELU) != 0;
        const bool has_res = (p.flags & SFB_GEMM_RESIDUAL) != 0 && !(p.flags & DBG_NORES);
        constexpr bool out_f32 = EPI == EPI_F32 || EPI == EPI_F32_LN;
        constexpr bool ln_fold = EPI == EPI_BF16_LN;
        constexpr bool emit_ln = EPI == EPI_F32_LN;
        const bool do_store = !(p.flags & DBG_NOSTORE);
        const float inv_k = 1.0f / static_cast<float>(p.K);
        uint8_t *stage = smem_raw + (tiles_base - smem_u32(smem_raw)) + kStages * STAGE_BYTES + (warp - 2) * EPI_WARP_BYTES;
        uint8_t *st_row = stage + lane * 128;                 // this thread's accumulator row in the transpose buffer
        const int rr = lane >> 3, cc = lane & 7;              // read-back mapping: rows 4i + rr, 16-byte chunk cc
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t tempty_leader0 = CG == 2 ? mapa_u32(tempty_bar(0), 0u) : tempty_bar(0);
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int m0 = (p.m_fastest ? tile % p.num_m_blocks : tile / p.num_n_blocks) * (BLOCK_M * CL) + crank * BLOCK_M + q * 32;
            const int n0 = (p.m_fastest ? tile / p.num_m_blocks : tile % p.num_n_blocks) * BLOCK_N + quarter * 64;
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N + quarter * 64);
            // The fp32 residual rows of this thread's read-back mapping do not depend on the accumulator: the first chunk's
            // are requested BEFORE waiting for the MMAs of this tile, so their HBM latency hides behind the main loop; the
            // second chunk's are requested as soon as the first chunk's registers are free (explicit registers: `out` may alias
            // `residual`, so the compiler will not hoist these loads above earlier stores by itself).
            float4 res[8];
            auto load_res = [&](int c) {
                const int gcol = n0 + c * 32 + cc * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t grow = m0 + 4 * i + rr;
                    res[i] = (grow < p.M && gcol < p.N) ? *reinterpret_cast<const float4 *>(p.residual + grow * p.ldr + gcol)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            if (out_f32 && has_res) {
                load_res(0);
                // chunk 1's residual rows are requested only after chunk 0's registers are free: pull their lines into L2 now
                // (no registers needed), so that second request is an L2 hit instead of a DRAM round trip
                if (p.ldr != 0 && n0 + 32 + cc * 4 < p.N) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int64_t grow = m0 + 4 * i + rr;
                        if (grow < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + grow * p.ldr + n0 + 32 + cc * 4));
                    }
                }
            }
            // LN_FOLD: mean / rstd of this thread's accumulator row from the partial sums the producer GEMM (or sfb_rowstats_cast) left
            float ln_mean = 0.f, ln_rstd = 1.f;
            if (ln_fold) {
                const int64_t grow = m0 + lane;
                if (grow < p.M) {
                    const float *sp = p.ln_stats + grow * p.ln_parts * 2;
                    float s1 = 0.f, s2 = 0.f;
                    if ((p.ln_parts & 1) == 0) {
                        for (int t = 0; t < p.ln_parts; t += 2) {
                            const float4 u = __ldg(reinterpret_cast<const float4 *>(sp + 2 * t));
                            s1 += u.x + u.z, s2 += u.y + u.w;
                        }
                    } else {
                        for (int t = 0; t < p.ln_parts; ++t) {
                            const float2 u = __ldg(reinterpret_cast<const float2 *>(sp + 2 * t));
                            s1 += u.x, s2 += u.y;
                        }
                    }
                    ln_mean = s1 * inv_k;
                    ln_rstd = rsqrtf(fmaxf(fmaf(-ln_mean, ln_mean, s2 * inv_k), 0.f) + p.ln_eps);
                }
            }
            float ln_ps = 0.f, ln_pq = 0.f;     // EMIT_LN: this lane's row (4 cc + rr) summed over the warp's 64 columns
            // The accumulator is handed back to the MMA issuer as soon as this warp's LAST tcgen05.ld of the tile has completed - the
            // bias / GELU / transpose / global stores that follow work on registers only.  (Releasing at the end of the epilogue, with a
            // release-ordered remote arrive behind the global stores, showed up as ~10 % of the stall samples and kept tile i+2's
            // main loop waiting for tile i's stores.)
            auto release_acc = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc); else mbar_arrive(tempty_bar(acc));
                }
            };
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (out_f32) {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int col0 = n0 + c * 32;
                    const int gcol = col0 + cc * 4;
                    uint32_t r[32];
                    tmem_ld32(tbase + c * 32, r);
                    tmem_ld_wait();
                    if (c == 1) release_acc();
                    if (col0 < p.N) {          // warp-uniform
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        if (p.bias != nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (col0 + j < p.N) {
                                    const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + col0 + j));
                                    v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
                                }
                            }
                        }
                        if (gelu) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) gelu_erf_fast2(v[j], v[j + 1]);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            *reinterpret_cast<float4 *>(st_row + ((k ^ (lane & 7)) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        __syncwarp();
                        float e1[8], e2[8];          // EMIT_LN: this lane's 4-column partial (sum, sum of squares) of rows 4 i + rr
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rloc = 4 * i + rr;
                            const int64_t grow = m0 + rloc;
                            float4 val = *reinterpret_cast<const float4 *>(stage + rloc * 128 + ((cc ^ (rloc & 7)) << 4));
                            const bool inb = grow < p.M && gcol < p.N;
                            if (inb) {
                                if (has_res) val.x += res[i].x, val.y += res[i].y, val.z += res[i].z, val.w += res[i].w;
                                if (do_store) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + grow * p.ldo + gcol) = val;
                            }
                            if (emit_ln) {          // compile-time
                                if (inb && do_store)
                                    *reinterpret_cast<uint2 *>(p.emit_bf16 + grow * p.ld_emit + gcol) = make_uint2(pack_bf16x2(val.x, val.y), pack_bf16x2(val.z, val.w));
                                e1[i] = inb ? (val.x + val.y) + (val.z + val.w) : 0.f;
                                e2[i] = inb ? fmaf(val.x, val.x, fmaf(val.y, val.y, fmaf(val.z, val.z, val.w * val.w))) : 0.f;
                            }
                        }
                        if (emit_ln) {
                            // The 8 lanes that share rr hold 4 columns each of rows 4 i + rr, i = 0..7.  Transposing reduction: every exchange
                            // halves the number of rows a lane is responsible for (7 shuffles per statistic instead of 24 butterflies), and lane
                            // cc ends up with the total of row 4 cc + rr.
                            const bool h4 = (cc & 4) != 0, h2 = (cc & 2) != 0, h1 = (cc & 1) != 0;
                            float b1[4], b2[4], c1[2], c2[2];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                b1[j] = (h4 ? e1[j + 4] : e1[j]) + __shfl_xor_sync(0xffffffffu, h4 ? e1[j] : e1[j + 4], 4);
                                b2[j] = (h4 ? e2[j + 4] : e2[j]) + __shfl_xor_sync(0xffffffffu, h4 ? e2[j] : e2[j + 4], 4);
                            }
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                c1[j] = (h2 ? b1[j + 2] : b1[j]) + __shfl_xor_sync(0xffffffffu, h2 ? b1[j] : b1[j + 2], 2);
                                c2[j] = (h2 ? b2[j + 2] : b2[j]) + __shfl_xor_sync(0xffffffffu, h2 ? b2[j] : b2[j + 2], 2);
                            }
                            ln_ps += (h1 ? c1[1] : c1[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c1[0] : c1[1], 1);
                            ln_pq += (h1 ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c2[0] : c2[1], 1);
                        }
                        __syncwarp();
                    }
                    if (has_res && c == 0) load_res(1);
                }
                if (emit_ln && n0 < p.N) {
                    const int64_t grow = m0 + 4 * cc + rr;
                    if (grow < p.M && do_store)
                        *reinterpret_cast<float2 *>(p.emit_stats + (grow * (p.N >> 6) + (n0 >> 6)) * 2) = make_float2(ln_ps, ln_pq);
                }
            } else {
                const int col0 = n0;
                if (p.tma_store) {        // the previous tile's TMA store must have finished READING this warp's staging buffer
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
                {
#pragma unroll
                    for (int hseg = 0; hseg < 2; ++hseg) {
                        uint32_t r[32];
                        tmem_ld32(tbase + hseg * 32, r);
                        tmem_ld_wait();
                        if (hseg == 1) release_acc();
                        if (col0 >= p.N) continue;       // warp-uniform
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        const int cb = col0 + hseg * 32;
                        if (ln_fold) {          // out = rstd (acc - mean colsum) + bias'
                            const float nm = -ln_mean;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (cb + j < p.N) {
                                    const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + cb + j));
                                    const float4 cs = __ldg(reinterpret_cast<const float4 *>(p.ln_colsum + cb + j));
                                    v[j] = fmaf(fmaf(nm, cs.x, v[j]), ln_rstd, b.x), v[j + 1] = fmaf(fmaf(nm, cs.y, v[j + 1]), ln_rstd, b.y);
                                    v[j + 2] = fmaf(fmaf(nm, cs.z, v[j + 2]), ln_rstd, b.z), v[j + 3] = fmaf(fmaf(nm, cs.w, v[j + 3]), ln_rstd, b.w);
                                }
                            }
                        } else if (p.bias != nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (cb + j < p.N) {
                                    const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + cb + j));
                                    v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
                                }
                            }
                        }
                        if (gelu) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) gelu_erf_fast2(v[j], v[j + 1]);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k)      // 16-byte chunk hseg*4 + k of this thread's 128-byte row
                            *reinterpret_cast<uint4 *>(st_row + (((hseg * 4 + k) ^ (lane & 7)) << 4)) =
                                make_uint4(pack_bf16x2(v[8 * k], v[8 * k + 1]), pack_bf16x2(v[8 * k + 2], v[8 * k + 3]),
                                           pack_bf16x2(v[8 * k + 4], v[8 * k + 5]), pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
                    }
                }
                if (col0 < p.N && p.tma_store) {
                    // the staging buffer IS the box layout of the output tensor map (32 rows x 128 B, 16-byte chunks XOR-swizzled by row):
                    // one elected lane hands it to the TMA unit, which clips rows >= M / columns >= N itself
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && do_store) {
                        tma_store_2d(&tmap_out, smem_u32(stage), col0, m0);
                        tma_store_commit();
                    }
                } else if (col0 < p.N) {
                    __syncwarp();
                    const int gcol = col0 + cc * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rloc = 4 * i + rr;
                        const int64_t grow = m0 + rloc;
                        const uint4 val = *reinterpret_cast<const uint4 *>(stage + rloc * 128 + ((cc ^ (rloc & 7)) << 4));
                        if (grow < p.M && gcol < p.N && do_store)
                            *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + grow * p.ldo + gcol) = val;
                    }
                    __syncwarp();
                }
            }
            if (++acc == kAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        if (!out_f32 && p.tma_store && lane == 0) tma_store_wait_all();      // smem must outlive the bulk stores that read it
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();       // nobody may still touch the peer's smem / TMEM / barriers
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// -------------------------------------------------------------------------- bring-up cross-check (CUDA cores)
// Plain shared-memory tiled GEMM with the same epilogue; used only by tests / SFB_GEMM_IMPL=1 to tell a tcgen05
// bug from a model-assembly bug.  Not on the product path.
__global__ void __launch_bounds__(256) gemm_bf16_simple_kernel(const __nv_bfloat16 *A, int64_t lda, const __nv_bfloat16 *W, EpiParams p) {
    __shared__ float sa[64][33];
    __shared__ float sw[64][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += 32) {
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {
            const int r = i >> 5, c = i & 31;
            const int gm = m0 + r, gn = n0 + r, gk = k0 + c;
            sa[r][c] = (gm < p.M && gk < p.K) ? __bfloat162float(A[static_cast<int64_t>(gm) * lda + gk]) : 0.f;
            sw[r][c] = (gn < p.N && gk < p.K) ? __bfloat162float(W[static_cast<int64_t>(gn) * p.K + gk]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sa[ty * 4 + i][k], b[i] = sw[tx * 4 + i][k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= p.M) continue;
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= p.N) continue;
            float v = acc[i][j];
            if (p.flags & SFB_GEMM_LN_FOLD) {
                float s1 = 0.f, s2 = 0.f;
                for (int t = 0; t < p.ln_parts; ++t) s1 += p.ln_stats[(static_cast<int64_t>(gm) * p.ln_parts + t) * 2], s2 += p.ln_stats[(static_cast<int64_t>(gm) * p.ln_parts + t) * 2 + 1];
                const float mean = s1 / p.K;
                const float rstd = rsqrtf(fmaxf(s2 / p.K - mean * mean, 0.f) + p.ln_eps);
                v = (v - mean * p.ln_colsum[gn]) * rstd;
            }
            v += p.bias ? p.bias[gn] : 0.f;
            if (p.flags & SFB_GEMM_GELU) v = gelu_erf(v);
            if (p.flags & SFB_GEMM_RESIDUAL) v += p.residual[static_cast<int64_t>(gm) * p.ldr + gn];
            if (p.flags & SFB_GEMM_OUT_F32)
                reinterpret_cast<float *>(p.out)[static_cast<int64_t>(gm) * p.ldo + gn] = v;
            else
                reinterpret_cast<__nv_bfloat16 *>(p.out)[static_cast<int64_t>(gm) * p.ldo + gn] = __float2bfloat16_rn(v);
        }
    }
}

// --------------------------------------------------------------------------------------------- host side
// 2D bf16 row-major (rows x cols, row stride ld elements) -> box (box_rows x 64 cols), 128B swizzle, zero OOB fill
static int make_tmap(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    return encode_tmap_bf16_2d(map, base, rows, cols, ld, box_rows, BLOCK_K);
}

}  // namespace gemm
}  // namespace sfb

extern "C" int sfb_gemm_bf16(const void *A, int64_t lda, const void *W, const float *bias, const float *residual, int64_t ldr,
                             void *out, int64_t ldo, int M, int N, int K, int flags, int impl, void *stream) {
    SFB_CHECK_ARG(!(flags & (SFB_GEMM_LN_FOLD | SFB_GEMM_EMIT_LN)), "sfb_gemm_bf16: the LayerNorm-fusion flags need sfb_gemm_bf16_ln");
    return sfb_gemm_bf16_ln(A, lda, W, bias, residual, ldr, out, ldo, M, N, K, flags, impl, nullptr, 0, nullptr, 0.f, nullptr, nullptr, 0, stream);
}

extern "C" int sfb_gemm_bf16_ln(const void *A, int64_t lda, const void *W, const float *bias, const float *residual, int64_t ldr,
                                void *out, int64_t ldo, int M, int N, int K, int flags, int impl, const float *ln_stats, int ln_parts,
                                const float *ln_colsum, float ln_eps, float *emit_stats, void *emit_bf16, int64_t ld_emit, void *stream) {
    using namespace sfb;
    using namespace sfb::gemm;
    SFB_CHECK_ARG(A && W && out, "sfb_gemm_bf16: null pointer");
    SFB_CHECK_ARG(M > 0 && N > 0 && K > 0, "sfb_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
    SFB_CHECK_ARG(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldo % 8 == 0 && lda >= K && ldo >= N,
                  "sfb_gemm_bf16: K, N, lda, ldo must be multiples of 8 (K=%d N=%d lda=%lld ldo=%lld)", K, N, (long long)lda,
                  (long long)ldo);
    SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "sfb_gemm_bf16: A, W, out must be 16-byte aligned");
    SFB_CHECK_ARG(impl == 1 || !(flags & SFB_GEMM_RESIDUAL) || (flags & SFB_GEMM_OUT_F32), "sfb_gemm_bf16: RESIDUAL requires OUT_F32 (fp32 residual stream)");
    if (flags & SFB_GEMM_RESIDUAL) {
        SFB_CHECK_ARG(residual != nullptr && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 && ldr % 4 == 0,
                      "sfb_gemm_bf16: residual must be non-null, 16-byte aligned, ldr %% 4 == 0");
    }
    SFB_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "sfb_gemm_bf16: bias must be 16-byte aligned");
    if (flags & SFB_GEMM_LN_FOLD) {
        SFB_CHECK_ARG(!(flags & SFB_GEMM_OUT_F32) || impl == 1, "sfb_gemm_bf16_ln: LN_FOLD writes bf16 (qkv / fc1 style consumers)");
        SFB_CHECK_ARG(ln_stats && ln_colsum && bias && ln_parts > 0 && ln_eps >= 0.f, "sfb_gemm_bf16_ln: LN_FOLD needs ln_stats, ln_colsum, bias, ln_parts > 0");
        SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(ln_stats) & 15) == 0 && (reinterpret_cast<uintptr_t>(ln_colsum) & 15) == 0,
                      "sfb_gemm_bf16_ln: ln_stats / ln_colsum must be 16-byte aligned");
    }
    if (flags & SFB_GEMM_EMIT_LN) {
        SFB_CHECK_ARG(impl != 1 && (flags & SFB_GEMM_OUT_F32), "sfb_gemm_bf16_ln: EMIT_LN belongs to the fp32-output (residual stream) epilogue of the tcgen05 kernel");
        SFB_CHECK_ARG(emit_stats && emit_bf16 && N % 64 == 0 && ld_emit >= N && ld_emit % 4 == 0 && (reinterpret_cast<uintptr_t>(emit_bf16) & 7) == 0 &&
                          (reinterpret_cast<uintptr_t>(emit_stats) & 7) == 0,
                      "sfb_gemm_bf16_ln: EMIT_LN needs emit_stats (M x N/64 x 2 floats), emit_bf16 (row stride ld_emit >= N, %% 4), N %% 64 == 0");
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    EpiParams p;
    p.bias = bias, p.residual = residual, p.ldr = ldr, p.out = out, p.ldo = ldo;
    p.M = M, p.N = N, p.K = K, p.flags = flags;
    p.ln_stats = ln_stats, p.ln_colsum = ln_colsum, p.ln_parts = ln_parts, p.ln_eps = ln_eps;
    p.emit_stats = emit_stats, p.emit_bf16 = reinterpret_cast<__nv_bfloat16 *>(emit_bf16), p.ld_emit = ld_emit;
    p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    p.num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    static const int order_env = getenv("SFB_GEMM_ORDER") ? atoi(getenv("SFB_GEMM_ORDER")) : 0;
    p.m_fastest = order_env;
    static const int dbg_env = getenv("SFB_GEMM_DBG") ? atoi(getenv("SFB_GEMM_DBG")) : 0;
    p.flags |= (dbg_env & 7) << 16;

    if (impl == 1) {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        gemm_bf16_simple_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(A), lda,
                                                       reinterpret_cast<const __nv_bfloat16 *>(W), p);
        SFB_CHECK_LAUNCH();
        return SFB_OK;
    }
    SFB_CHECK_ARG(impl == 0 || impl == 2 || impl == 3, "sfb_gemm_bf16: unknown impl %d (0 auto, 1 CUDA-core check, 2 single-CTA, 3 CTA-pair)", impl);
    // CTA pairs (256-row tiles) unless the problem has at most one 128-row block
    const int cg = impl == 2 ? 1 : impl == 3 ? 2 : (M > BLOCK_M ? 2 : 1);
    int epi = (flags & SFB_GEMM_OUT_F32) ? ((flags & SFB_GEMM_EMIT_LN) ? EPI_F32_LN : EPI_F32) : ((flags & SFB_GEMM_LN_FOLD) ? EPI_BF16_LN : EPI_BF16);
    // fp32 outputs go through the TMA epilogue whenever their rows form regular 2D tensors (SFB_GEMM_F32_TMA=0: per-thread epilogue, A/B aid);
    // a broadcast residual row (ldr == 0) and unaligned EMIT_LN copies stay on the per-thread epilogue
    const int f32_tma_env = getenv("SFB_GEMM_F32_TMA") ? atoi(getenv("SFB_GEMM_F32_TMA")) : 1;             // read per call: tools/gemm_ab.py flips it
    if ((flags & SFB_GEMM_OUT_F32) && f32_tma_env && ldo % 4 == 0 && (!(flags & SFB_GEMM_RESIDUAL) || ldr >= N) &&
        (!(flags & SFB_GEMM_EMIT_LN) || ((reinterpret_cast<uintptr_t>(emit_bf16) & 15) == 0 && ld_emit % 8 == 0)))
        epi = EPI_F32_TMA;

    CUtensorMap tmap_a, tmap_w;
    int rc = make_tmap(&tmap_a, A, M, K, lda, BLOCK_M);
    if (rc != SFB_OK) return rc;
    rc = make_tmap(&tmap_w, W, N, K, K, BLOCK_N / cg);           // rows of W fetched by one CTA per k-block
    if (rc != SFB_OK) return rc;
    // bf16 outputs leave through TMA stores: one 32-row x 64-column box per epilogue warp (SFB_GEMM_TMA_STORE=0: per-thread st.global, A/B aid)
    static const int tma_store_env = getenv("SFB_GEMM_TMA_STORE") ? atoi(getenv("SFB_GEMM_TMA_STORE")) : 1;
    CUtensorMap tmap_out = tmap_a, tmap_res = tmap_a, tmap_emit = tmap_a;
    p.tma_store = 0;
    if (tma_store_env && !(flags & SFB_GEMM_OUT_F32)) {
        rc = encode_tmap_bf16_2d(&tmap_out, out, M, N, ldo, 32, 64);
        if (rc != SFB_OK) return rc;
        p.tma_store = 1;
    }
    if (epi == EPI_F32_TMA) {
        rc = encode_tmap_f32_2d(&tmap_out, out, M, N, ldo, 32, 32, true);
        if (rc == SFB_OK && (flags & SFB_GEMM_RESIDUAL)) {
            rc = encode_tmap_f32_2d(&tmap_res, residual, M, N, ldr, 32, 32, true);
        }
        if (rc == SFB_OK && (flags & SFB_GEMM_EMIT_LN)) rc = encode_tmap_bf16_2d(&tmap_emit, emit_bf16, M, N, ld_emit, 32, 64);
        if (rc != SFB_OK) return rc;
    }

    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
#define SFB_GEMM_ATTR(CGV, EPIV) \
    SFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<CGV, EPIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<CGV>::SMEM_BYTES))
        SFB_GEMM_ATTR(1, EPI_BF16); SFB_GEMM_ATTR(1, EPI_BF16_LN); SFB_GEMM_ATTR(1, EPI_F32); SFB_GEMM_ATTR(1, EPI_F32_LN); SFB_GEMM_ATTR(1, EPI_F32_TMA);
        SFB_GEMM_ATTR(2, EPI_BF16); SFB_GEMM_ATTR(2, EPI_BF16_LN); SFB_GEMM_ATTR(2, EPI_F32); SFB_GEMM_ATTR(2, EPI_F32_LN); SFB_GEMM_ATTR(2, EPI_F32_TMA);
#undef SFB_GEMM_ATTR
    }
    p.num_m_blocks = (M + cg * BLOCK_M - 1) / (cg * BLOCK_M);           // 128-row (single CTA) or 256-row (pair) tiles
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(epi == EPI_F32_TMA ? kThreadsTma : kThreads);
    cfg.dynamicSmemBytes = cg == 2 ? Cfg<2>::SMEM_BYTES : Cfg<1>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    const int max_clusters = num_sms() / cg;
    const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
    cfg.gridDim = dim3(cg * clusters);
#define SFB_GEMM_LAUNCH(CGV, EPIV) \
    SFB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<CGV, EPIV>, tmap_a, tmap_w, tmap_out, tmap_res, tmap_emit, p))
    if (cg == 2) {
        switch (epi) {
            case EPI_BF16: SFB_GEMM_LAUNCH(2, EPI_BF16); break;
            case EPI_BF16_LN: SFB_GEMM_LAUNCH(2, EPI_BF16_LN); break;
            case EPI_F32: SFB_GEMM_LAUNCH(2, EPI_F32); break;
            case EPI_F32_TMA: SFB_GEMM_LAUNCH(2, EPI_F32_TMA); break;
            default: SFB_GEMM_LAUNCH(2, EPI_F32_LN); break;
        }
    } else {
        switch (epi) {
            case EPI_BF16: SFB_GEMM_LAUNCH(1, EPI_BF16); break;
            case EPI_BF16_LN: SFB_GEMM_LAUNCH(1, EPI_BF16_LN); break;
            case EPI_F32: SFB_GEMM_LAUNCH(1, EPI_F32); break;
            case EPI_F32_TMA: SFB_GEMM_LAUNCH(1, EPI_F32_TMA); break;
            default: SFB_GEMM_LAUNCH(1, EPI_F32_LN); break;
        }
    }
#undef SFB_GEMM_LAUNCH
    return SFB_OK;
}
