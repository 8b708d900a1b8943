// Host-side plumbing of the C-ABI: error text, device probing.
#include <stdarg.h>
#include <string.h>

#include <cuda.h>

#include "common.cuh"

namespace sfb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev & 63;
}

int num_sms() {
    static int n[64] = {};
    const int dev = current_device_slot();
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

bool PerDeviceOnce::first() {
    const int dev = current_device_slot();
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}
void PerDeviceOnce::reset_current() { done[current_device_slot()] = false; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// TMA descriptor for a 2D bf16 row-major matrix (rows x cols, row stride ld elements): box (box_rows x box_cols), 128B swizzle
// (box_cols * 2 bytes must be 128), zero fill outside the matrix.  The driver entry point is resolved through the runtime, so the
// library does not link libcuda.
int encode_tmap_bf16_2d(void *map, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return SFB_E_CUDA;
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld cols %lld ld %lld box %d x %d)", static_cast<int>(r),
                  static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld), box_rows, box_cols);
        return SFB_E_CUDA;
    }
    return SFB_OK;
}

// The same for a 2D fp32 row-major matrix: box (box_rows x box_cols); `swizzle128` = 128B swizzle (box_cols * 4 bytes must be 128, used for
// shared-memory tiles) or none (boxes that are only prefetched into L2: up to 256 columns).
int encode_tmap_f32_2d(void *map, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols, bool swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return SFB_E_CUDA;
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rows %lld cols %lld ld %lld box %d x %d)", static_cast<int>(r),
                  static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld), box_rows, box_cols);
        return SFB_E_CUDA;
    }
    return SFB_OK;
}

}  // namespace sfb

extern "C" int sfb_abi_version(void) { return 1; }

extern "C" const char *sfb_last_error(void) { return sfb::g_err; }

extern "C" int sfb_device_check(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        sfb::set_error("no CUDA device");
        return SFB_E_CUDA;
    }
    if (major != 10) {
        sfb::set_error("synchformer_b200 kernels are built for sm_100a only; device has compute capability %d.x", major);
        return SFB_E_UNSUPPORTED;
    }
    return SFB_OK;
}
