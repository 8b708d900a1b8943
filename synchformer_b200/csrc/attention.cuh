// Internal view of sfb_attn_desc shared by the attention kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfb {
namespace attn {

struct Desc {
    const __nv_bfloat16 *q, *k, *v, *kp, *vp;
    __nv_bfloat16 *out;
    int64_t q_outer, q_inner, q_row;
    int64_t kv_outer, kv_inner, kv_row;
    int64_t o_outer, o_inner, o_row;
    int64_t prefix_outer;
    int has_prefix;
    int n_outer, n_inner, n_heads, Lq, Lk;
    float scale;
    const __nv_bfloat16 *xq;      // optional extra query row per (outer, head)
    int64_t xq_outer;
    float *xpartial;
};

// tcgen05 / TMEM kernel for hd 64, 128 < Lq <= 256, Lk + prefix <= 256 (Motionformer space attention); attention_tc.cu
bool tc_supported(const Desc &d);
int launch_tc(const Desc &d, cudaStream_t st);

}  // namespace attn
}  // namespace sfb
