// CPU SIMT emulator shim — TEST INFRASTRUCTURE ONLY (tests/emu/build_emu.py compiles the UNMODIFIED kernel sources
// synchformer_b200/csrc/train.cu and attention_train.cu against this header instead of the real common.cuh).
//
// Purpose: execute the exact device code of the N3 kernels (and the exact host launch code of their C-ABI entry points) on a box
// without a GPU, so that indexing, barrier placement, shared-memory layout and launch configuration are checked before the first
// hardware run.  It is not a performance model and knows nothing about sm_100a; it only gives CUDA's SIMT semantics to g++:
//   * every CUDA thread of a block is a ucontext coroutine; blocks run one after another
//   * __syncthreads / __syncwarp / __shfl_xor_sync are rendezvous points of the coroutine scheduler (exited threads do not block)
//   * __shared__ variables are function-level statics (one instance, blocks are sequential); dynamic shared memory is one buffer
//   * kernel<<<grid, block, smem, stream>>>(args) is rewritten by build_emu.py into emu::launch(grid, block, smem, [&]{ kernel(args); })
// A deadlock (a barrier that not every live thread reaches) aborts with a message instead of hanging.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static
#define __CUDA_NO_BFLOAT16_OPERATORS__

#include <vector_types.h>
#include <vector_functions.h>
#include <cuda_bf16.h>

#include "../../include/synchformer_b200.h"

// the CUDA headers above declare the runtime API (types are reused); the four calls the entry points make are redirected here
inline const char *emu_cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t emu_cudaGetLastError() { return cudaSuccess; }
template <typename K>
inline cudaError_t emu_cudaFuncSetAttribute(K, int, int bytes) { return bytes <= 227 * 1024 ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t emu_cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) {
    memset(p, v, n);
    return cudaSuccess;
}
template <typename K>
inline cudaError_t emu_cudaOccupancy(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor emu_cudaOccupancy
#define cudaGetErrorString emu_cudaGetErrorString
#define cudaGetLastError emu_cudaGetLastError
#define cudaFuncSetAttribute emu_cudaFuncSetAttribute
#define cudaMemsetAsync emu_cudaMemsetAsync

namespace emu {

// Context switch.  glibc's swapcontext makes a sigprocmask system call per switch (~1 us; a 600 x 3072 transpose is 470 k threads), so
// on x86-64 the coroutines switch with a 14-instruction callee-saved-register swap instead; elsewhere ucontext is the fallback.
#if defined(__x86_64__)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void **save_sp, void *load_sp);
#ifdef EMU_MAIN_TU                      // defined once, in the translation unit build_emu.py generates for the exports
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
#endif
#else
#define EMU_FAST_SWITCH 0
#endif

struct Thread {
#if EMU_FAST_SWITCH
    void *sp = nullptr;
#else
    ucontext_t ctx;
#endif
    char *stack = nullptr;
    int state = 0;   // 0 runnable, 1 at __syncthreads, 2 at a warp rendezvous, 3 exited
};

inline uint3 &tid() { static uint3 v; return v; }
inline uint3 &bid() { static uint3 v; return v; }
inline dim3 &bdim() { static dim3 v; return v; }
inline dim3 &gdim() { static dim3 v; return v; }
inline unsigned char *&dyn_smem() { static unsigned char *p = nullptr; return p; }
inline int &schedule() { static int m = 0; return m; }
inline uint64_t &rng() { static uint64_t r = 0x9E3779B97F4A7C15ull; return r; }

struct Sched {
    std::vector<Thread> threads;
    std::vector<uint64_t> slot;       // per-thread exchange slot for shuffles
    std::vector<uint32_t> wide;       // per-thread 8-word exchange area for ldmatrix / mma fragments
#if EMU_FAST_SWITCH
    void *main_sp = nullptr;
#else
    ucontext_t main_ctx;
#endif
    int cur = 0;
    const std::function<void()> *body = nullptr;
    long launches = 0;
};
inline Sched &S() { static Sched s; return s; }

constexpr size_t kStack = 256 * 1024;

#if EMU_FAST_SWITCH
inline void yield_to_scheduler() { emu_switch(&S().threads[S().cur].sp, S().main_sp); }
inline void resume(int t) { emu_switch(&S().main_sp, S().threads[t].sp); }
inline void thread_entry() {
    (*S().body)();
    S().threads[S().cur].state = 3;
    yield_to_scheduler();
    abort();                          // an exited thread is never resumed
}
inline void prepare(Thread &th) {
    // stack image that emu_switch "returns" into: six zeroed callee-saved registers, then the entry address in a 16-byte aligned slot
    uintptr_t top = (reinterpret_cast<uintptr_t>(th.stack) + kStack) & ~static_cast<uintptr_t>(15);
    void **slot = reinterpret_cast<void **>(top - 16);
    slot[0] = reinterpret_cast<void *>(&thread_entry);
    void **sp = slot - 6;
    for (int i = 0; i < 6; ++i) sp[i] = nullptr;
    th.sp = sp;
}
#else
inline void yield_to_scheduler() { swapcontext(&S().threads[S().cur].ctx, &S().main_ctx); }
inline void resume(int t) { swapcontext(&S().main_ctx, &S().threads[t].ctx); }
inline void thread_entry() {
    (*S().body)();
    S().threads[S().cur].state = 3;
}
inline void prepare(Thread &th) {
    getcontext(&th.ctx);
    th.ctx.uc_stack.ss_sp = th.stack;
    th.ctx.uc_stack.ss_size = kStack;
    th.ctx.uc_link = &S().main_ctx;
    makecontext(&th.ctx, reinterpret_cast<void (*)()>(thread_entry), 0);
}
#endif

inline void run_block(int n_threads, const std::function<void()> &fn) {
    Sched &s = S();
    if (static_cast<int>(s.threads.size()) < n_threads) {
        s.threads.resize(n_threads);
        s.slot.resize(n_threads);
    }
    s.body = &fn;
    for (int t = 0; t < n_threads; ++t) {
        Thread &th = s.threads[t];
        if (!th.stack) th.stack = static_cast<char *>(malloc(kStack));
        prepare(th);
        th.state = 0;
    }
    const dim3 bd = bdim();
    static std::vector<int> order;
    order.resize(n_threads);
    for (;;) {
        bool progressed = false, live = false;
        // run order of the runnable threads within a pass: forward, reverse or reshuffled every pass.  CUDA promises no order between
        // barriers, so a correctly synchronised kernel gives bit-identical results under all three (tests run them all).
        for (int t = 0; t < n_threads; ++t) order[t] = schedule() == 1 ? n_threads - 1 - t : t;
        if (schedule() == 2)
            for (int t = n_threads - 1; t > 0; --t) {
                rng() = rng() * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[t], order[static_cast<int>((rng() >> 33) % static_cast<uint64_t>(t + 1))]);
            }
        for (int k = 0; k < n_threads; ++k) {
            const int t = order[k];
            if (s.threads[t].state != 0) continue;
            s.cur = t;
            tid() = make_uint3(t % bd.x, (t / bd.x) % bd.y, t / (bd.x * bd.y));
            resume(t);
            progressed = true;
        }
        // warp rendezvous: every live thread of the warp is waiting
        for (int w = 0; w * 32 < n_threads; ++w) {
            int waiting = 0, alive = 0;
            for (int t = w * 32; t < std::min(n_threads, w * 32 + 32); ++t) {
                alive += s.threads[t].state != 3;
                waiting += s.threads[t].state == 2;
            }
            if (waiting && waiting == alive) {
                for (int t = w * 32; t < std::min(n_threads, w * 32 + 32); ++t)
                    if (s.threads[t].state == 2) s.threads[t].state = 0;
                progressed = true;
            }
        }
        int at_bar = 0, alive = 0;
        for (int t = 0; t < n_threads; ++t) {
            alive += s.threads[t].state != 3;
            at_bar += s.threads[t].state == 1;
            live |= s.threads[t].state != 3;
        }
        if (at_bar && at_bar == alive) {
            for (int t = 0; t < n_threads; ++t)
                if (s.threads[t].state == 1) s.threads[t].state = 0;
            progressed = true;
        }
        if (!live) break;
        if (!progressed) {
            fprintf(stderr, "emu: DEADLOCK in block (%u,%u): %d threads alive, %d at __syncthreads, the rest stuck at a warp rendezvous\n", bid().x,
                    bid().y, alive, at_bar);
            abort();
        }
    }
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &fn) {
    static std::vector<unsigned char> smem;
    if (smem.size() < smem_bytes + 1024) smem.resize(smem_bytes + 1024);
    // poison dynamic shared memory so that reads of never-written bytes show up as NaNs / huge values instead of zeros
    dyn_smem() = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem.data()) + 1023) & ~static_cast<uintptr_t>(1023));
    gdim() = grid;
    bdim() = block;
    S().launches++;
    const int n_threads = static_cast<int>(block.x * block.y * block.z);
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                memset(dyn_smem(), 0xFF, smem_bytes);
                bid() = make_uint3(x, y, z);
                run_block(n_threads, fn);
            }
}

}  // namespace emu

#define threadIdx (emu::tid())
#define blockIdx (emu::bid())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())

inline void __syncthreads() {
    emu::S().threads[emu::S().cur].state = 1;
    emu::yield_to_scheduler();
}
inline void __syncwarp(unsigned = 0xffffffffu) {
    emu::S().threads[emu::S().cur].state = 2;
    emu::yield_to_scheduler();
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    emu::Sched &s = emu::S();
    const int me = s.cur;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    s.slot[me] = bits;
    __syncwarp();
    const int src = (me & ~31) | ((me ^ lane_mask) & 31);
    uint64_t got = s.slot[src];
    __syncwarp();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    emu::Sched &s = emu::S();
    const int me = s.cur;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    s.slot[me] = bits;
    __syncwarp();
    uint64_t got = s.slot[(me & ~31) | (src_lane & 31)];
    __syncwarp();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T>
inline T __ldg(const T *p) { return *p; }

// ---- the few PTX instructions the hardware-verified mma.sync attention kernels use (build_emu.py rewrites their asm statements into
// these calls).  Shared-memory "addresses" are byte offsets into the dynamic shared memory of the running block. -------------------------
inline size_t __cvta_generic_to_shared(const void *p) { return static_cast<size_t>(reinterpret_cast<const unsigned char *>(p) - emu::dyn_smem()); }
namespace emu {
inline void warp_publish(const uint32_t *mine, int n) {          // every lane stores n words; visible to the warp after the rendezvous
    Sched &s = S();
    if (s.wide.size() < s.threads.size() * 8) s.wide.resize(s.threads.size() * 8);
    for (int k = 0; k < n; ++k) s.wide[static_cast<size_t>(s.cur) * 8 + k] = mine[k];
    __syncwarp();
}
inline uint32_t warp_peek(int lane, int k) { return S().wide[static_cast<size_t>((S().cur & ~31) | lane) * 8 + k]; }
inline void ptx_cp_async16(uint32_t dst, const void *src) { memcpy(dyn_smem() + dst, src, 16); }
inline void ptx_nop() {}
// ldmatrix .m8n8 .b16: lane 8 i + r supplies the address of row r of matrix i; lane T receives, per matrix, the 32-bit word holding
// elements (T / 4, 2 (T % 4)) and (T / 4, 2 (T % 4) + 1) - or, with .trans, elements (2 (T % 4), T / 4) and (2 (T % 4) + 1, T / 4)
inline void ldmatrix(bool trans, int n_mat, uint32_t addr, uint32_t *out) {
    warp_publish(&addr, 1);
    const int T = S().cur & 31;
    for (int i = 0; i < n_mat; ++i) {
        if (!trans) {
            const unsigned char *row = dyn_smem() + warp_peek(8 * i + T / 4, 0);
            memcpy(&out[i], row + 4 * (T % 4), 4);
        } else {
            uint16_t lo, hi;
            memcpy(&lo, dyn_smem() + warp_peek(8 * i + 2 * (T % 4), 0) + 2 * (T / 4), 2);
            memcpy(&hi, dyn_smem() + warp_peek(8 * i + 2 * (T % 4) + 1, 0) + 2 * (T / 4), 2);
            out[i] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
        }
    }
    __syncwarp();
}
inline void ptx_ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
    uint32_t o[4];
    ldmatrix(false, 4, addr, o);
    r0 = o[0], r1 = o[1], r2 = o[2], r3 = o[3];
}
inline void ptx_ldmatrix_x4_trans(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
    uint32_t o[4];
    ldmatrix(true, 4, addr, o);
    r0 = o[0], r1 = o[1], r2 = o[2], r3 = o[3];
}
inline void ptx_ldmatrix_x2(uint32_t &r0, uint32_t &r1, uint32_t addr) {
    uint32_t o[2];
    ldmatrix(false, 2, addr, o);
    r0 = o[0], r1 = o[1];
}
inline float bf16_half(uint32_t w, int hi) {
    const uint32_t bits = hi ? (w & 0xffff0000u) : (w << 16);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
// mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 (PTX ISA fragment layouts; g = lane / 4, t = lane % 4):
//   A: a0 (g, 2t..), a1 (g + 8, 2t..), a2 (g, 2t + 8..), a3 (g + 8, 2t + 8..);  B: b0 (k = 2t.., n = g), b1 (k = 2t + 8.., n = g)
//   C / D: c0 (g, 2t), c1 (g, 2t + 1), c2 (g + 8, 2t), c3 (g + 8, 2t + 1)
inline void ptx_mma_16816(float &c0, float &c1, float &c2, float &c3, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    const uint32_t mine[6] = {a0, a1, a2, a3, b0, b1};
    warp_publish(mine, 6);
    const int T = S().cur & 31, g = T / 4, t = T % 4;
    float *c[4] = {&c0, &c1, &c2, &c3};
    for (int q = 0; q < 4; ++q) {
        const int row = g + 8 * (q / 2), col = 2 * t + (q % 2);
        float acc = *c[q];
        for (int k = 0; k < 16; ++k) {
            const int a_lane = (row % 8) * 4 + (k % 8) / 2, a_reg = (row / 8) + 2 * (k / 8);
            const int b_lane = col * 4 + (k % 8) / 2, b_reg = 4 + k / 8;
            acc += bf16_half(warp_peek(a_lane, a_reg), k % 2) * bf16_half(warp_peek(b_lane, b_reg), k % 2);
        }
        *c[q] = acc;
    }
    __syncwarp();
}
inline void ptx_ex2(float &y, float x) { y = exp2f(x); }
}  // namespace emu
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
inline float __uint_as_float(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __expf(float x) { return expf(x); }
inline float __fdividef(float a, float b) { return a / b; }
using std::max;
using std::min;

namespace sfb {

constexpr int kD = 768;

inline char *err_buf() { static char b[512] = ""; return b; }
inline void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
}
inline int num_sms() { return 148; }
struct PerDeviceOnce {          // one emulated device
    bool done = false;
    bool first() { bool f = !done; done = true; return f; }
    void reset_current() { done = false; }
};

#define SFB_CHECK_ARG(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            sfb::set_error(__VA_ARGS__);  \
            return SFB_E_INVALID;         \
        }                                 \
    } while (0)
#define SFB_CHECK_CUDA(expr)                                  \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) {                             \
            sfb::set_error("%s failed (emulated)", #expr);    \
            return SFB_E_CUDA;                                \
        }                                                     \
    } while (0)
#define SFB_CHECK_LAUNCH() SFB_CHECK_CUDA(cudaGetLastError())

// the helpers of the real common.cuh that the N3 kernels use (the PTX-based ones are not needed by them)
inline float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
inline float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
inline float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
inline uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    uint32_t u;
    memcpy(&u, &t, 4);
    return u;
}
inline float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 t;
    memcpy(&t, &u, 4);
    return __bfloat1622float2(t);
}

}  // namespace sfb
