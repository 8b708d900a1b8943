// Micro-benchmark (experiment aid, not product): cost of short bursts of tcgen05.mma in the shapes the space-attention kernel issues.
// One CTA per SM slot; one thread issues `n_mma` MMAs of (M128, N, K16), commits, waits; clock64 around issue and around completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I synchformer_b200/csrc tools/ubench/mma_cost.cu -o gpurun_out/mma_cost
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tcgen05.cuh"
using namespace sfb::tc;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// mode 0: SS, B K-major      1: SS, B MN-major     2: TS (A from TMEM), B K-major     3: TS, B MN-major
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int n_mma, long long *out, int use_elect) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (uint32_t i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    if (threadIdx.x < 32 && (use_elect ? elect_one() : threadIdx.x == 0)) {
        const uint32_t sA = base, sB = base + 32768;
        const bool ts = mode == 2 || mode == 3 || mode == 5, mn = mode == 1 || mode == 3 || mode == 5;
        const uint32_t idesc = make_idesc_major(128, N, 0, mn ? 1 : 0);
        for (int rep = 0; rep < 4; ++rep) {
            const long long t0 = clock64();
            if (mode == 4) {
                const uint64_t da = make_sw128_desc(sA), db = make_sw128_desc(sB);
#pragma unroll
                for (int i = 0; i < 13; ++i) umma_bf16(tm + 256, da + (i & 3) * 2, db + (i & 3) * 2, idesc, i != 0);
            } else if (mode == 5) {
                const uint64_t db = make_sw128_mn_desc(sB, 26624);
#pragma unroll
                for (int i = 0; i < 13; ++i) umma_bf16_ts(tm + 256, tm + i * 8, db + i * 128, idesc, i != 0);
            } else
            for (int i = 0; i < n_mma; ++i) {
                const uint64_t bdesc = mn ? make_sw128_mn_desc(sB + (i % 13) * 2048, 26624) : make_sw128_desc(sB + (i % 4) * 32);
                if (ts) umma_bf16_ts(tm + 256, tm + (i % 13) * 8, bdesc, idesc, i != 0);
                else umma_bf16(tm + 256, make_sw128_desc(sA + (i % 4) * 32), bdesc, idesc, i != 0);
            }
            const long long t1 = clock64();
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), rep & 1);
            const long long t2 = clock64();
            if (blockIdx.x == 0) out[rep * 2] = t1 - t0, out[rep * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
    long long *d, h[8];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char *names[4] = {"SS K-major B", "SS MN-major B", "TS K-major B", "TS MN-major B"};
    for (int mode = 0; mode < 4; ++mode)
        for (int N : {64, 128, 208, 256})
            for (int n : {1, 4, 13, 26}) {
                if ((mode & 1) && N != 64) continue;       // MN-major B laid out for N = 64 only
                k<<<1, 128, 100 * 1024>>>(mode, N, n, d, 0);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                printf("%-14s N=%3d n_mma=%2d: issue %5lld clk, issue->commit seen %5lld clk  (floor %4d clk) %s\n", names[mode], N, n, h[6], h[7], n * 128 * N / 256,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    for (int el = 0; el < 2; ++el)
    for (int mode : {0, 3, 4, 5})
        for (int N : {64, 208}) {
            if ((mode == 5 || mode == 3) && N != 64) continue;
            k<<<1, 128, 100 * 1024>>>(mode, N, 13, d, el);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("elect=%d mode=%d (0 SS loop, 3 TS-MN loop, 4 SS unrolled, 5 TS-MN unrolled) %s N=%3d: issue %5lld clk, issue->commit seen %5lld clk  (floor %4d clk) %s\n", el, mode, "", N, h[6], h[7],
                   13 * 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
