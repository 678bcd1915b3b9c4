// What surrounds the 2D assembly in the reference's drivers, on the device:
//
//   fh_q4_boundary_lists      <- HelmHoltz.generate_mesh, boundary node / element detection    helmholtz.py:95-118
//   fh_compact_flags_i32         (ascending index list of the flagged entries: np.where(...)[0])
//   fh_q4_neumann_load_c128   <- *.get_normal_derivative at the edge Gauss points              fem2d.py:12-27, :113-127
//   fh_q4_gauss_points        <- the (x, y) of base.py:141-148 for every element and Gauss point
//   fh_q4_l2_error_c128       <- BaseSolver.compute_L2_error                                   base.py:119-177
//
// The analytic solutions the L2 error is measured against are closed forms in Bessel functions evaluated by the caller
// (scipy, as the reference does) at the points fh_q4_gauss_points returns; the Neumann load needs J_m only, which CUDA's
// jn() provides -- to ~1e-11 absolute, not to the last bit of scipy's: that path is an option (default: host scipy
// values, bit-identical load vector), see solvers/base.py.
#include "fh_q4.cuh"

namespace fh {

// ------------------------------------------------------------------------------------------------ boundary detection
// node_flag[i] = 1 if np.isclose(np.linalg.norm(nodes[i]), radius, atol=atol)  (rtol = numpy's default 1e-5;
// norm over axis 1 is sqrt(x*x + y*y) with a plain add, not the BLAS ddot of a 1-d norm).
__global__ void __launch_bounds__(kQ4Threads)
q4_node_on_circle_kernel(const double *__restrict__ nodes, int64_t n, double radius, double atol,
                         int32_t *__restrict__ node_flag) {
    const int64_t i = (int64_t)blockIdx.x * kQ4Threads + threadIdx.x;
    if (i >= n) return;
    const double x = nodes[2 * i], y = nodes[2 * i + 1];
    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
    const bool on = fabs(__dsub_rn(dist, radius)) <= __dadd_rn(atol, __dmul_rn(1e-5, fabs(radius)));
    node_flag[i] = on ? 1 : 0;
}
// elem_flag[e] = 1 if any node of element e is flagged (helmholtz.py:114-118)
__global__ void __launch_bounds__(kQ4Threads)
q4_elem_touches_kernel(const int32_t *__restrict__ elements, int64_t ne, const int32_t *__restrict__ node_flag,
                       int32_t *__restrict__ elem_flag) {
    const int64_t e = (int64_t)blockIdx.x * kQ4Threads + threadIdx.x;
    if (e >= ne) return;
    int f = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) f |= node_flag[elements[4 * e + a]];
    elem_flag[e] = f;
}

// Ascending list of the indices i with flags[i] != 0, and their number: one CTA, tiles of 1024 entries, ballot + one
// scan of the 32 warp totals per tile (lists of boundary entities are short; the input is read once, coalesced).
constexpr int kListThreads = 1024;
__global__ void __launch_bounds__(kListThreads)
compact_flags_kernel(const int32_t *__restrict__ flags, int64_t n, int32_t *__restrict__ list, int32_t *__restrict__ count) {
    __shared__ int warp_tot[kListThreads / 32];
    __shared__ int warp_off[kListThreads / 32 + 1];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int base = 0;
    for (int64_t s0 = 0; s0 < n; s0 += kListThreads) {
        const int64_t i = s0 + t;
        const int f = (i < n && flags[i] != 0) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const int v = warp_tot[lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            warp_off[lane] = incl - v;
            if (lane == 31) warp_off[32] = incl;
        }
        __syncthreads();
        if (f) list[base + warp_off[warp] + __popc(bal & ((1u << lane) - 1u))] = (int32_t)i;
        base += warp_off[32];
        __syncthreads();
    }
    if (t == 0) *count = base;
}

// ------------------------------------------------------------------------------------------------ Neumann load
// g[q][edge][gauss] = sum_{m=-nf}^{nf} i^m k J_m'(k r) exp(i m theta) at the Gauss points of the flagged edges of the
// listed elements (fem2d.py:12-27 evaluated where fem2d.py:113-127 asks for it);  J_m' = (J_{m-1} - J_{m+1}) / 2.
__device__ __forceinline__ double bessel_j(int m, double x) {  // integer order of either sign
    const double v = jn(m < 0 ? -m : m, x);
    return (m < 0 && (m & 1)) ? -v : v;
}
__global__ void __launch_bounds__(kQ4Threads)
q4_neumann_load_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ elements,
                       const int32_t *__restrict__ ids, int64_t nb, const uint8_t *__restrict__ on_edge, double k,
                       int n_fourier, Gauss gq, cplx *__restrict__ g) {
    const int64_t o = (int64_t)blockIdx.x * kQ4Threads + threadIdx.x;  // (element, edge, gauss point)
    const int64_t per = 4 * gq.n;
    if (o >= nb * per) return;
    const int64_t q = o / per;
    const int ed = (int)((o - q * per) / gq.n), gi = (int)(o - q * per - (int64_t)ed * gq.n);
    cplx out = czero();
    if (on_edge[4 * q + ed]) {
        const int64_t e = ids[q];
        const int64_t n0 = elements[4 * e + ed], n1 = elements[4 * e + ((ed + 1) & 3)];
        const double xi = gq.pt[gi];
        const double N0 = __dmul_rn(0.5, __dsub_rn(1.0, xi)), N1 = __dmul_rn(0.5, __dadd_rn(1.0, xi));  // base.py:57-63
        const double x = __dadd_rn(__dadd_rn(0.0, __dmul_rn(N0, nodes[2 * n0])), __dmul_rn(N1, nodes[2 * n1]));
        const double y = __dadd_rn(__dadd_rn(0.0, __dmul_rn(N0, nodes[2 * n0 + 1])), __dmul_rn(N1, nodes[2 * n1 + 1]));
        const double r = sqrt(x * x + y * y), theta = atan2(y, x);
        for (int m = -n_fourier; m <= n_fourier; ++m) {
            const double jp = 0.5 * (bessel_j(m - 1, k * r) - bessel_j(m + 1, k * r));
            double sn, cs;
            sincos((double)m * theta, &sn, &cs);
            cplx term = mk(k * jp * cs, k * jp * sn);
            switch (((m % 4) + 4) % 4) {  // i^m
                case 1: term = mk(-term.im, term.re); break;
                case 2: term = mk(-term.re, -term.im); break;
                case 3: term = mk(term.im, -term.re); break;
                default: break;
            }
            out = out + term;
        }
    }
    g[o] = out;
}

// ------------------------------------------------------------------------------------------------ L2 error
// Physical coordinates of the Gauss points, x = ((((0 + N0 x0) + N1 x1) + N2 x2) + N3 x3) as base.py:141-148 adds them up:
// xy[e][i][j] = (x, y) at (xi_i, eta_j).
__global__ void __launch_bounds__(kQ4Threads)
q4_gauss_points_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ elements, int64_t ne, Gauss gq,
                       double *__restrict__ xy) {
    const int64_t e = (int64_t)blockIdx.x * kQ4Threads + threadIdx.x;
    if (e >= ne) return;
    double x[4], y[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t nd = elements[4 * e + a];
        x[a] = nodes[2 * nd], y[a] = nodes[2 * nd + 1];
    }
    for (int i = 0; i < gq.n; ++i)
        for (int j = 0; j < gq.n; ++j) {
            const Q4Shape s = q4_shape(gq.pt[i], gq.pt[j]);
            double px = 0.0, py = 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a) px = __dadd_rn(px, __dmul_rn(s.N[a], x[a])), py = __dadd_rn(py, __dmul_rn(s.N[a], y[a]));
            double *o = xy + 2 * ((e * gq.n + i) * gq.n + j);
            o[0] = px, o[1] = py;
        }
}

constexpr int kL2CtasMax = 1024;
struct Q4L2Ws {
    double part[kL2CtasMax];
    unsigned int done;
};
// sum over elements and Gauss points of |u_h - u_exact|^2 det J w_i w_j (base.py:127-175); out[0] = the sum (the caller
// takes the square root, base.py:177).  Element sums in the reference's order; the sum over elements is a fixed-shape tree.
__global__ void __launch_bounds__(kQ4Threads)
q4_l2_error_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ elements, int64_t ne,
                   const cplx *__restrict__ u, const cplx *__restrict__ u_exact, Gauss gq, Q4L2Ws *__restrict__ ws,
                   double *__restrict__ out) {
    __shared__ double red[kQ4Threads];
    __shared__ bool last;
    double acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * kQ4Threads + threadIdx.x; e < ne; e += (int64_t)gridDim.x * kQ4Threads) {
        double x[4], y[4];
        cplx ue[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int64_t nd = elements[4 * e + a];
            x[a] = nodes[2 * nd], y[a] = nodes[2 * nd + 1];
            ue[a] = u[nd];
        }
        double elem = 0.0;
        for (int i = 0; i < gq.n; ++i)
            for (int j = 0; j < gq.n; ++j) {
                const Q4Shape s = q4_shape(gq.pt[i], gq.pt[j]);
                double ur = 0.0, ui = 0.0, J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    ur = __dadd_rn(ur, __dmul_rn(s.N[a], ue[a].re));
                    ui = __dadd_rn(ui, __dmul_rn(s.N[a], ue[a].im));
                    J00 = __dadd_rn(J00, __dmul_rn(s.dxi[a], x[a]));
                    J01 = __dadd_rn(J01, __dmul_rn(s.dxi[a], y[a]));
                    J10 = __dadd_rn(J10, __dmul_rn(s.deta[a], x[a]));
                    J11 = __dadd_rn(J11, __dmul_rn(s.deta[a], y[a]));
                }
                const double detJ = __dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J01, J10));
                const cplx ex = u_exact[(e * gq.n + i) * gq.n + j];
                const double a = hypot(__dsub_rn(ur, ex.re), __dsub_rn(ui, ex.im));  // abs(u_num - u_exact)
                const double err = __dmul_rn(a, a);
                elem = __dadd_rn(elem, __dmul_rn(__dmul_rn(__dmul_rn(err, detJ), gq.wt[i]), gq.wt[j]));
            }
        acc += elem;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int st = kQ4Threads / 2; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ws->part[blockIdx.x] = red[0];
        __threadfence();
        last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double v = 0.0;
        for (unsigned c = 0; c < gridDim.x; ++c) v += reinterpret_cast<const volatile double *>(ws->part)[c];
        out[0] = v;
        ws->done = 0;
    }
}

static Gauss gauss_from_host(const double *pts, const double *wts, int n) {
    Gauss g;
    g.n = n;
    for (int i = 0; i < 4; ++i) g.pt[i] = i < n ? pts[i] : 0.0, g.wt[i] = (i < n && wts != nullptr) ? wts[i] : 0.0;
    return g;
}

}  // namespace fh

extern "C" {

int fh_q4_boundary_lists(const double *nodes, int64_t n_nodes, const int32_t *elements, int64_t n_elements, double radius,
                         double atol, int32_t *node_flag, int32_t *elem_flag, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_nodes >= 0 && n_elements >= 0, "sizes must be >= 0");
    FH_REQUIRE((n_nodes == 0 || (nodes && node_flag)) && (n_elements == 0 || (elements && elem_flag && node_flag)),
               "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_nodes > 0) {
        q4_node_on_circle_kernel<<<(unsigned)((n_nodes + kQ4Threads - 1) / kQ4Threads), kQ4Threads, 0, st>>>(
            nodes, n_nodes, radius, atol, node_flag);
        FH_LAUNCH_CHECK("q4_node_on_circle_kernel");
    }
    if (n_elements > 0) {
        q4_elem_touches_kernel<<<(unsigned)((n_elements + kQ4Threads - 1) / kQ4Threads), kQ4Threads, 0, st>>>(
            elements, n_elements, node_flag, elem_flag);
        FH_LAUNCH_CHECK("q4_elem_touches_kernel");
    }
    return FH_OK;
}

int fh_compact_flags_i32(const int32_t *flags, int64_t n, int32_t *list, int32_t *count, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
    FH_REQUIRE(count && (n == 0 || (flags && list)), "NULL device pointer");
    compact_flags_kernel<<<1, kListThreads, 0, static_cast<cudaStream_t>(stream)>>>(flags, n, list, count);
    FH_LAUNCH_CHECK("compact_flags_kernel");
    return FH_OK;
}

int fh_q4_neumann_load_c128(const double *nodes, const int32_t *elements, const int32_t *elem_ids, int64_t n_ids,
                            const uint8_t *on_edge, double k, int n_fourier, const double *gauss_pts_host, int n_gauss,
                            fh_c128 *g_vals, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_ids >= 0 && n_fourier >= 0, "n_ids and n_fourier must be >= 0");
    if (n_ids == 0) return FH_OK;
    FH_REQUIRE(nodes && elements && elem_ids && on_edge && gauss_pts_host && g_vals, "NULL pointer");
    FH_REQUIRE(n_gauss >= 1 && n_gauss <= 4, "n_gauss must be in [1, 4]");
    const Gauss g = gauss_from_host(gauss_pts_host, nullptr, n_gauss);
    const int64_t total = n_ids * 4 * n_gauss;
    q4_neumann_load_kernel<<<(unsigned)((total + kQ4Threads - 1) / kQ4Threads), kQ4Threads, 0,
                             static_cast<cudaStream_t>(stream)>>>(nodes, elements, elem_ids, n_ids, on_edge, k, n_fourier,
                                                                  g, reinterpret_cast<cplx *>(g_vals));
    FH_LAUNCH_CHECK("q4_neumann_load_kernel");
    return FH_OK;
}

int fh_q4_gauss_points(const double *nodes, const int32_t *elements, int64_t n_elements, const double *gauss_pts_host,
                       int n_gauss, double *xy, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_elements >= 0, "n_elements must be >= 0");
    if (n_elements == 0) return FH_OK;
    FH_REQUIRE(nodes && elements && gauss_pts_host && xy, "NULL pointer");
    FH_REQUIRE(n_gauss >= 1 && n_gauss <= 4, "n_gauss must be in [1, 4]");
    const Gauss g = gauss_from_host(gauss_pts_host, nullptr, n_gauss);
    q4_gauss_points_kernel<<<(unsigned)((n_elements + kQ4Threads - 1) / kQ4Threads), kQ4Threads, 0,
                             static_cast<cudaStream_t>(stream)>>>(nodes, elements, n_elements, g, xy);
    FH_LAUNCH_CHECK("q4_gauss_points_kernel");
    return FH_OK;
}

size_t fh_q4_l2_error_workspace_bytes(void) { return sizeof(fh::Q4L2Ws); }

int fh_q4_l2_error_c128(const double *nodes, const int32_t *elements, int64_t n_elements, const fh_c128 *u,
                        const fh_c128 *u_exact, const double *gauss_pts_host, const double *gauss_wts_host, int n_gauss,
                        double *sum_out, void *workspace, size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_elements >= 1, "n_elements must be >= 1");
    FH_REQUIRE(nodes && elements && u && u_exact && gauss_pts_host && gauss_wts_host && sum_out && workspace, "NULL pointer");
    FH_REQUIRE(n_gauss >= 1 && n_gauss <= 4, "n_gauss must be in [1, 4]");
    FH_REQUIRE(workspace_bytes >= sizeof(Q4L2Ws), "workspace too small (fh_q4_l2_error_workspace_bytes)");
    const Gauss g = gauss_from_host(gauss_pts_host, gauss_wts_host, n_gauss);
    int64_t grid = (n_elements + kQ4Threads - 1) / kQ4Threads;
    if (grid > kL2CtasMax) grid = kL2CtasMax;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Q4L2Ws *ws = static_cast<Q4L2Ws *>(workspace);
    FH_CUDA(cudaMemsetAsync(&ws->done, 0, sizeof(unsigned int), st));
    q4_l2_error_kernel<<<(unsigned)grid, kQ4Threads, 0, st>>>(nodes, elements, n_elements,
                                                             reinterpret_cast<const cplx *>(u),
                                                             reinterpret_cast<const cplx *>(u_exact), g, ws, sum_out);
    FH_LAUNCH_CHECK("q4_l2_error_kernel");
    return FH_OK;
}

}  // extern "C"
