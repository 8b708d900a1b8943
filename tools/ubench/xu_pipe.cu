// Micro-benchmark (experiment aid): which issue pipe do the instructions of the softmax inner loop share?
// One warp per SM sub-partition (128 threads, 1 CTA per SM), 16 independent chains per thread, clock64 around 64 rounds.
// Modes: 0 ex2 only | 1 cvt.rn.bf16x2.f32 only | 2 (fmul + ex2) x2 + cvt | 3 (fmul + ex2) x2 + integer truncation pack (PRMT)
//        4 ex2 x2 + cvt + ffma x2 + fadd x2 (full mix) | 5 same with PRMT pack | 6 polynomial 2^x on the FMA pipe only
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/xu_pipe tools/ubench/xu_pipe.cu && /tmp/xu_pipe
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_rn(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t pack_tr(float a, float b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); return r; }
// 2^x for x <= 0 on the FMA / ALU pipes: round-to-nearest split + degree-3 polynomial on [-0.5, 0.5], exponent added as an integer
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.f);
    const float r = x + 12582912.f;                  // 1.5 * 2^23: integer part in the low mantissa bits
    const float f = x - (r - 12582912.f);            // [-0.5, 0.5]
    float p = fmaf(f, 0.0555041086f, 0.2402265069f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long *out, float *sink, float seed) {
    float v[16];
    uint32_t acc = 0;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = seed * (threadIdx.x + j) - 3.f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 64; ++rep) {
        seed += 1e-4f;                                   // inputs change every round: nothing is loop-invariant
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            if (MODE == 0) { v[j] = ex2f(v[j]) - 1.5f; v[j + 1] = ex2f(v[j + 1]) - 1.5f; }
            if (MODE == 1) { acc ^= pack_rn(v[j], v[j + 1]); v[j] += 1.0f; }
            if (MODE == 2) { const float a = ex2f(v[j] * seed), b = ex2f(v[j + 1] * seed); acc ^= pack_rn(a, b); }
            if (MODE == 3) { const float a = ex2f(v[j] * seed), b = ex2f(v[j + 1] * seed); acc ^= pack_tr(a, b); }
            if (MODE == 4) { const float a = ex2f(fmaf(v[j], seed, -0.5f)), b = ex2f(fmaf(v[j + 1], seed, -0.5f)); sum += a + b; acc ^= pack_rn(a, b); }
            if (MODE == 5) { const float a = ex2f(fmaf(v[j], seed, -0.5f)), b = ex2f(fmaf(v[j + 1], seed, -0.5f)); sum += a + b; acc ^= pack_tr(a, b); }
            if (MODE == 6) { const float a = ex2_poly(fmaf(v[j], seed, -0.5f)), b = ex2_poly(fmaf(v[j + 1], seed, -0.5f)); sum += a + b; acc ^= pack_tr(a, b); }
            if (MODE == 7) {       // half MUFU, half polynomial
                const float a = ex2f(fmaf(v[j], seed, -0.5f)), b = ex2_poly(fmaf(v[j + 1], seed, -0.5f)); sum += a + b; acc ^= pack_tr(a, b); }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[MODE] = t1 - t0;
    float s = sum + __uint_as_float(acc);
#pragma unroll
    for (int j = 0; j < 16; ++j) s += v[j];
    if (s == 123.456f) sink[0] = s;
}

int main() {
    long long *out; float *sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4); cudaMemset(out, 0, 64);
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<148, 128>>>(out, sink, 0.01f); k<1><<<148, 128>>>(out, sink, 0.01f); k<2><<<148, 128>>>(out, sink, 0.01f); k<3><<<148, 128>>>(out, sink, 0.01f);
        k<4><<<148, 128>>>(out, sink, 0.01f); k<5><<<148, 128>>>(out, sink, 0.01f); k<6><<<148, 128>>>(out, sink, 0.01f); k<7><<<148, 128>>>(out, sink, 0.01f);
    }
    long long h[8];
    cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    const char *names[8] = {"ex2 only", "cvt.rn.bf16x2 only", "2 ex2 + cvt", "2 ex2 + prmt", "softmax mix (cvt)", "softmax mix (prmt)", "polynomial 2^x + prmt", "half ex2 half poly + prmt"};
    // per round and thread: 16 elements (8 pairs); clocks per ELEMENT per warp
    for (int m = 0; m < 8; ++m) printf("%-28s %8lld clk  %.2f clk/element/warp\n", names[m], h[m], h[m] / (64.0 * 16));
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    // accuracy of the polynomial
    return 0;
}
