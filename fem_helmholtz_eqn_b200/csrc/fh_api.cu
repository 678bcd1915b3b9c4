// Error plumbing, version and device queries of libfemhelmholtz_b200.
#include <stdarg.h>
#include <string.h>

#include "fh_common.cuh"

namespace fh {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    (void)cudaGetLastError();  // clear the sticky-free error so the next call starts clean
    return static_cast<int>(e);
}

static int query(int attr_kind) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceAttr a = attr_kind == 0 ? cudaDevAttrMultiProcessorCount
                                      : cudaDevAttrMaxSharedMemoryPerBlockOptin;
    if (cudaDeviceGetAttribute(&v, a, dev) != cudaSuccess) return 0;
    return v;
}

int sm_count() { return query(0); }
size_t smem_optin_bytes() { return static_cast<size_t>(query(1)); }

}  // namespace fh

extern "C" {

const char *fh_last_error(void) { return fh::g_err; }

int fh_abi_version(void) { return FH_ABI_VERSION; }

int fh_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host, size_t *smem_optin_host) {
    int dev = 0;
    FH_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count_host) {
        FH_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count_host = v;
    }
    if (cc_major_host) {
        FH_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major_host = v;
    }
    if (cc_minor_host) {
        FH_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor_host = v;
    }
    if (smem_optin_host) {
        FH_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        *smem_optin_host = static_cast<size_t>(v);
    }
    return FH_OK;
}

}  // extern "C"
