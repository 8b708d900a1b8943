// Micro-benchmark (experiment aid): when does the mbarrier of a tcgen05.commit fire if MORE MMAs are issued right after it?
// Thread 0 issues group A (4 x M128 N208 K16, SS) + commit(barA), then group B (13 x M128 N64 K16, TS, MN-major B) + commit(barB).
// Warp 1 waits on barA, warp 2 on barB; all stamp clock64 relative to the start of the issue.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tcgen05.cuh"
using namespace sfb::tc;

__global__ void __launch_bounds__(128, 1) k(long long *out, int with_b) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    __shared__ long long t_start;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (uint32_t i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); t_start = 0; }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        if (threadIdx.x == 0) {
            const uint64_t dq = make_sw128_desc(base), dk = make_sw128_desc(base + 32768), dv = make_sw128_mn_desc(base + 65536, 26624);
            const uint32_t ids = make_idesc_major(128, 208, 0, 0), ido = make_idesc_major(128, 64, 0, 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) umma_bf16(tm, dq + 2 * i, dk + 2 * i, ids, i != 0);
            umma_commit(smem_u32(&bars[0]));
            const long long t1 = clock64();
            if (with_b) {
#pragma unroll
                for (int i = 0; i < 13; ++i) umma_bf16_ts(tm + 256 + 128, tm + 256 + i * 8, dv + 128 * i, ido, i != 0);
            }
            umma_commit(smem_u32(&bars[1]));
            const long long t2 = clock64();
            out[rep * 8 + 0] = t1 - t0, out[rep * 8 + 1] = t2 - t0;
        } else if (threadIdx.x == 32) {
            mbar_wait(smem_u32(&bars[0]), rep & 1);
            out[rep * 8 + 2] = clock64() - t0;
        } else if (threadIdx.x == 64) {
            mbar_wait(smem_u32(&bars[1]), rep & 1);
            out[rep * 8 + 3] = clock64() - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
    long long *d, h[24];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int with_b = 0; with_b < 2; ++with_b) {
        cudaMemset(d, 0, sizeof(h));
        k<<<1, 128, 100 * 1024>>>(d, with_b);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("group B %s: A issued+committed at %lld, B issued+committed at %lld, barA seen by its waiter at %lld, barB seen at %lld  %s\n", with_b ? "issued" : "EMPTY ",
               h[16], h[17], h[18], h[19], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
