// Host-side plumbing of the C-ABI: error text, device probing.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sfb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
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
