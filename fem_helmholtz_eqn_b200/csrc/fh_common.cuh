// Shared device/host helpers for libfemhelmholtz_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fem_helmholtz.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfemhelmholtz_b200 is written for sm_100a (B200) only"
#endif

namespace fh {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define FH_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return fh::cuda_fail(e__, #call);  \
    } while (0)

#define FH_REQUIRE(cond, ...)                  \
    do {                                       \
        if (!(cond)) {                         \
            fh::set_error(__VA_ARGS__);        \
            return FH_E_INVALID;               \
        }                                      \
    } while (0)

#define FH_LAUNCH_CHECK(name)                                              \
    do {                                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return fh::cuda_fail(e__, "launch " name); \
    } while (0)

int sm_count();                 // cached per current device
size_t smem_optin_bytes();      // cached per current device

// ---------------------------------------------------------------- complex128 arithmetic
// Plain struct of two doubles, 16-byte aligned so that loads/stores are single 128-bit
// transactions (LDG.128 / STG.128 / LDS.128).
struct __align__(16) cplx {
    double re, im;
};

__host__ __device__ __forceinline__ cplx mk(double re, double im) {
    cplx r;
    r.re = re;
    r.im = im;
    return r;
}
__host__ __device__ __forceinline__ cplx czero() { return mk(0.0, 0.0); }
__host__ __device__ __forceinline__ cplx cone() { return mk(1.0, 0.0); }
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx operator-(cplx a) { return mk(-a.re, -a.im); }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
// c - a*b  (4 FMAs)
__device__ __forceinline__ cplx nmsub(cplx c, cplx a, cplx b) {
    double re = fma(-a.re, b.re, c.re);
    re = fma(a.im, b.im, re);
    double im = fma(-a.re, b.im, c.im);
    im = fma(-a.im, b.re, im);
    return mk(re, im);
}
// c + a*b  (4 FMAs)
__device__ __forceinline__ cplx cmadd(cplx c, cplx a, cplx b) {
    double re = fma(a.re, b.re, c.re);
    re = fma(-a.im, b.im, re);
    double im = fma(a.re, b.im, c.im);
    im = fma(a.im, b.re, im);
    return mk(re, im);
}
// 1/a.  The solver's pivots are O(1/h) .. O(k^2 h); |a|^2 neither over- nor underflows
// in any configuration of interest, so the unscaled formula is used (cheaper than Smith's).
__device__ __forceinline__ cplx crecip(cplx a) {
    double s = 1.0 / fma(a.re, a.re, a.im * a.im);
    return mk(a.re * s, -a.im * s);
}
// 1/x to ~1 ulp: hardware seed (MUFU.RCP64H, relative error <= 2^-19.9 measured over 2^26 arguments,
// tools/rcp_probe.cu) + ONE cubically convergent step r (1 + e + e^2), e = 1 - x r: three dependent FMAs instead of the
// four of two Newton steps (every reciprocal sits on the critical path of an elimination chain); measured error
// <= 2.2e-16.  Non-finite / zero inputs give NaN or Inf, which the status check reports.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double p = fma(e, e, e);
    return fma(r, p, r);
}
__device__ __forceinline__ cplx frecip(cplx a) {
    const double s = fast_rcp(fma(a.re, a.re, a.im * a.im));
    return mk(a.re * s, -a.im * s);
}
__device__ __forceinline__ bool cfinite(cplx a) { return isfinite(a.re) && isfinite(a.im); }

// ---------------------------------------------------------------- streaming memory ops
__device__ __forceinline__ cplx ldg_stream(const cplx *p) {  // read-once global data: skip L1
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return mk(v.x, v.y);
}
__device__ __forceinline__ void stg_stream(cplx *p, cplx v) {  // write-once global data
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.re), "d"(v.im)
                 : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte shared-memory read that keeps its place in program order (volatile asm statements are not reordered
// against each other): used where the ORDER of reads and arithmetic is part of the design.
__device__ __forceinline__ cplx lds_volatile(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return mk(v.x, v.y);
}
__device__ __forceinline__ void sts_volatile(uint32_t addr, cplx v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.re), "d"(v.im) : "memory");
}
// 16-byte asynchronous global->shared copy (LDGSTS), L2 only.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

}  // namespace fh
