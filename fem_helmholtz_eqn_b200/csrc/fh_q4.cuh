// Helpers shared by the 2D bilinear-quad kernels (fh_q4.cu: assembly; fh_q4_post.cu: what comes before and after it).
#pragma once
#include "fh_common.cuh"

namespace fh {

constexpr int kQ4Threads = 128;

__device__ __forceinline__ double norm2_blas(double x, double y) {  // np.linalg.norm([x, y])
    return __dsqrt_rn(__fma_rn(y, y, __dmul_rn(x, x)));
}
__device__ __forceinline__ bool isclose_default(double a, double b) {  // np.isclose(a, b): rtol 1e-5, atol 1e-8
    return fabs(__dsub_rn(a, b)) <= __dadd_rn(1e-8, __dmul_rn(1e-5, fabs(b)));
}

struct Q4Shape {
    double N[4], dxi[4], deta[4];
};
__device__ __forceinline__ Q4Shape q4_shape(double xi, double eta) {  // base.py:32-55
    Q4Shape s;
    const double mx = __dsub_rn(1.0, xi), px = __dadd_rn(1.0, xi);
    const double me = __dsub_rn(1.0, eta), pe = __dadd_rn(1.0, eta);
    s.N[0] = __dmul_rn(__dmul_rn(0.25, mx), me);
    s.N[1] = __dmul_rn(__dmul_rn(0.25, px), me);
    s.N[2] = __dmul_rn(__dmul_rn(0.25, px), pe);
    s.N[3] = __dmul_rn(__dmul_rn(0.25, mx), pe);
    s.dxi[0] = __dmul_rn(-0.25, me), s.dxi[1] = __dmul_rn(0.25, me);
    s.dxi[2] = __dmul_rn(0.25, pe), s.dxi[3] = __dmul_rn(-0.25, pe);
    s.deta[0] = __dmul_rn(-0.25, mx), s.deta[1] = __dmul_rn(-0.25, px);
    s.deta[2] = __dmul_rn(0.25, px), s.deta[3] = __dmul_rn(0.25, mx);
    return s;
}

struct Gauss {
    double pt[4], wt[4];
    int n;
};

}  // namespace fh
