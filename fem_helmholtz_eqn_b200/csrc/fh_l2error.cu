// L2 error of the 1D FEM solution against the continuum solution -exp(i k x), with the reference's quirks
// (nb:259-281): the COMPLEX error is squared without abs(), each Gauss term is weighted by h * w (not h/2 * w), and
// the caller takes sqrt of the complex sum.  One CTA per system, one element per thread and step, fixed-shape tree
// reduction (deterministic; the reference's strictly sequential sum is reproduced to rounding, not bit for bit).
// Reads 16 B per DOF (+8 B of node coordinates), writes 16 B per system: bandwidth-bound.
#include "fh_common.cuh"

namespace fh {

constexpr int kL2Threads = 256;

struct GaussL2 {
    double pt[4], wt[4];
    int n;
};

__global__ void __launch_bounds__(kL2Threads)
l2_error_1d_kernel(const double *__restrict__ x, const cplx *__restrict__ u, const double *__restrict__ ks,
                   int64_t n_elem, int64_t batch, GaussL2 gq, cplx *__restrict__ out) {
    __shared__ double sre[kL2Threads], sim[kL2Threads];
    for (int64_t s = blockIdx.x; s < batch; s += gridDim.x) {
        const double k = ks[s];
        const cplx *us = u + s * (n_elem + 1);
        double are = 0.0, aim = 0.0;
        for (int64_t e = threadIdx.x; e < n_elem; e += kL2Threads) {
            const double x0 = x[e], x1 = x[e + 1];
            const double h = x1 - x0;
            const cplx u0 = us[e], u1 = us[e + 1];
            for (int g = 0; g < gq.n; ++g) {
                const double b0 = (1.0 - gq.pt[g]) / 2, b1 = (1.0 + gq.pt[g]) / 2;  // nb:65
                const double xg = b0 * x0 + b1 * x1;
                const double ure = b0 * u0.re + b1 * u1.re, uim = b0 * u0.im + b1 * u1.im;
                double sn, cs;
                sincos(k * xg, &sn, &cs);
                const double dre = -cs - ure, dim = -sn - uim;  // analytical_solution(x, k) - Ue, nb:219,277
                const double hw = h * gq.wt[g];
                are += (dre * dre - dim * dim) * hw;            // complex square, no abs()
                aim += (2.0 * dre * dim) * hw;
            }
        }
        sre[threadIdx.x] = are, sim[threadIdx.x] = aim;
        __syncthreads();
        for (int st = kL2Threads / 2; st > 0; st >>= 1) {
            if (threadIdx.x < st) sre[threadIdx.x] += sre[threadIdx.x + st], sim[threadIdx.x] += sim[threadIdx.x + st];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[s] = mk(sre[0], sim[0]);
        __syncthreads();
    }
}

}  // namespace fh

extern "C" int fh_l2_error_1d_c128(const double *nodes, const fh_c128 *u, const double *ks, int64_t n_elem,
                                   int64_t batch, const double *gauss_pts_host, const double *gauss_wts_host,
                                   int n_gauss, fh_c128 *sum_out, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_elem >= 1 && batch >= 0, "n_elem must be >= 1 and batch >= 0");
    if (batch == 0) return FH_OK;
    FH_REQUIRE(nodes && u && ks && gauss_pts_host && gauss_wts_host && sum_out, "NULL pointer");
    FH_REQUIRE(n_gauss >= 1 && n_gauss <= 4, "n_gauss must be in [1, 4]");
    GaussL2 g;
    g.n = n_gauss;
    for (int i = 0; i < 4; ++i) g.pt[i] = i < n_gauss ? gauss_pts_host[i] : 0.0, g.wt[i] = i < n_gauss ? gauss_wts_host[i] : 0.0;
    int64_t grid = 8ll * sm_count();
    if (grid > batch) grid = batch;
    l2_error_1d_kernel<<<(unsigned)grid, kL2Threads, 0, static_cast<cudaStream_t>(stream)>>>(
        nodes, reinterpret_cast<const cplx *>(u), ks, n_elem, batch, g, reinterpret_cast<cplx *>(sum_out));
    FH_LAUNCH_CHECK("l2_error_1d_kernel");
    return FH_OK;
}
