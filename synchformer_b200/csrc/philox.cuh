// Counter-based dropout masks for the training kernels (N3): Philox4x32-10 (Salmon et al., SC'11; the generator family torch's CUDA
// dropout uses, but NOT torch's stream: a mask is a pure function of (seed, site, element index), so the backward kernels regenerate it
// instead of storing it, and the CPU oracle (oracle/philox.py) reproduces it bit for bit).
//
//   keep(seed, site, e) = philox4x32_10(counter = (lo32(e >> 2), hi32(e >> 2), site, 0), key = (lo32(seed), hi32(seed)))[e & 3] >= thr
//   thr = min(floor(p * 2^32), 2^32 - 1);  kept elements are scaled by 1 / (1 - p)          (nn.Dropout semantics)
//
// `site` separates the dropout layers of one step (0 = embedding dropout sync_model.py:137, 1 + 3 i = attention probabilities of
// block i transformer.py:47,74, 2 + 3 i = residual dropout after attn.proj :48,76, 3 + 3 i = residual dropout after the MLP :92).
#pragma once
#include <stdint.h>

namespace sfb {

struct DropParams {
    uint32_t thr;       // 0 = dropout off
    float inv_keep;     // 1 / (1 - p)
    uint32_t seed_lo, seed_hi, site;
};

static inline DropParams make_drop_params(float p, uint64_t seed, uint32_t site) {
    DropParams dp;
    double t = static_cast<double>(p) * 4294967296.0;
    if (t > 4294967295.0) t = 4294967295.0;
    dp.thr = p > 0.f ? static_cast<uint32_t>(t) : 0u;
    dp.inv_keep = dp.thr != 0u ? 1.0f / (1.0f - p) : 1.0f;
    dp.seed_lo = static_cast<uint32_t>(seed & 0xffffffffull);
    dp.seed_hi = static_cast<uint32_t>(seed >> 32);
    dp.site = site;
    return dp;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// multiplier applied to element e: 0 if dropped, 1 / (1 - p) if kept (1 if dropout is off)
__device__ __forceinline__ float drop_scale(const DropParams &dp, uint64_t e) {
    if (dp.thr == 0u) return 1.0f;
    const uint64_t ctr = e >> 2;
    const uint4 r = philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), dp.site, 0u, dp.seed_lo, dp.seed_hi);
    const uint32_t lane = static_cast<uint32_t>(e & 3);
    const uint32_t u = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
    return u >= dp.thr ? dp.inv_keep : 0.0f;
}

}  // namespace sfb
