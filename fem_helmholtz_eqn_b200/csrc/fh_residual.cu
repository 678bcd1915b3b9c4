// Residual / parity metrics of a batch of tridiagonal systems, on device (gate G3):
// out[s] = { ||b - A u||_2^2, ||u||_2^2, ||b||_2^2, ||A||_inf }.  One CTA per system (grid-stride),
// coalesced streaming reads (96 B per row), fixed-shape tree reductions => deterministic results.
#include "fh_common.cuh"

namespace fh {

constexpr int kResThreads = 256;
__device__ __forceinline__ double abs1c(cplx z) { return fabs(z.re) + fabs(z.im); }  // LAPACK's cabs1

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(kResThreads)
residual_kernel(const cplx *__restrict__ dl, const cplx *__restrict__ d, const cplx *__restrict__ du,
                const cplx *__restrict__ b, const cplx *__restrict__ u, int64_t n, int64_t batch,
                double *__restrict__ out) {
    __shared__ double red[4][kResThreads / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int64_t s = blockIdx.x; s < batch; s += gridDim.x) {
        const int64_t base = s * n;
        double r2 = 0.0, u2 = 0.0, b2 = 0.0, amax = 0.0;
        for (int64_t j = t; j < n; j += kResThreads) {
            const cplx uj = u[base + j], bj = ldg_stream(b + base + j), dj = ldg_stream(d + base + j);
            cplx r = nmsub(bj, dj, uj);
            double rowsum = hypot(dj.re, dj.im);
            if (j > 0) {
                const cplx lj = ldg_stream(dl + base + j);
                r = nmsub(r, lj, u[base + j - 1]);
                rowsum += hypot(lj.re, lj.im);
            }
            if (j < n - 1) {
                const cplx hj = ldg_stream(du + base + j);
                r = nmsub(r, hj, u[base + j + 1]);
                rowsum += hypot(hj.re, hj.im);
            }
            r2 += r.re * r.re + r.im * r.im;
            u2 += uj.re * uj.re + uj.im * uj.im;
            b2 += bj.re * bj.re + bj.im * bj.im;
            amax = fmax(amax, rowsum);
        }
        r2 = warp_sum(r2), u2 = warp_sum(u2), b2 = warp_sum(b2), amax = warp_max(amax);
        if (lane == 0) red[0][warp] = r2, red[1][warp] = u2, red[2][warp] = b2, red[3][warp] = amax;
        __syncthreads();
        if (warp == 0) {
            const int nw = kResThreads / 32;
            double v0 = lane < nw ? red[0][lane] : 0.0, v1 = lane < nw ? red[1][lane] : 0.0;
            double v2 = lane < nw ? red[2][lane] : 0.0, v3 = lane < nw ? red[3][lane] : 0.0;
            v0 = warp_sum(v0), v1 = warp_sum(v1), v2 = warp_sum(v2), v3 = warp_max(v3);
            if (lane == 0) out[4 * s + 0] = v0, out[4 * s + 1] = v1, out[4 * s + 2] = v2, out[4 * s + 3] = v3;
        }
        __syncthreads();
    }
}

// r = b - A u, element-wise (the right-hand side of one step of iterative refinement)
__global__ void __launch_bounds__(kResThreads)
residual_vector_kernel(const cplx *__restrict__ dl, const cplx *__restrict__ d, const cplx *__restrict__ du,
                       const cplx *__restrict__ b, const cplx *__restrict__ u, int64_t n, int64_t total,
                       cplx *__restrict__ r) {
    for (int64_t o = (int64_t)blockIdx.x * kResThreads + threadIdx.x; o < total; o += (int64_t)gridDim.x * kResThreads) {
        const int64_t j = o % n;
        cplx v = nmsub(b[o], d[o], u[o]);
        if (j > 0) v = nmsub(v, dl[o], u[o - 1]);
        if (j < n - 1) v = nmsub(v, du[o], u[o + 1]);
        r[o] = v;
    }
}

}  // namespace fh

extern "C" int fh_tridiag_residual_vector_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                               const fh_c128 *u, int64_t n, int64_t batch, fh_c128 *r,
                                               fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && batch >= 0, "n and batch must be >= 0");
    if (n == 0 || batch == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u && r, "NULL device pointer");
    const int64_t total = n * batch;
    int64_t grid = (total + kResThreads - 1) / kResThreads;
    const int64_t cap = 16ll * sm_count();
    if (grid > cap) grid = cap;
    residual_vector_kernel<<<(unsigned)grid, kResThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d), reinterpret_cast<const cplx *>(du),
        reinterpret_cast<const cplx *>(b), reinterpret_cast<const cplx *>(u), n, total, reinterpret_cast<cplx *>(r));
    FH_LAUNCH_CHECK("residual_vector_kernel");
    return FH_OK;
}

extern "C" int fh_tridiag_residual_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                        const fh_c128 *u, int64_t n, int64_t batch, double *out4,
                                        fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && batch >= 0, "n and batch must be >= 0");
    if (n == 0 || batch == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u && out4, "NULL device pointer");
    int64_t grid = 8ll * sm_count();
    if (grid > batch) grid = batch;
    if (grid < 1) grid = 1;
    residual_kernel<<<(unsigned)grid, kResThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d), reinterpret_cast<const cplx *>(du),
        reinterpret_cast<const cplx *>(b), reinterpret_cast<const cplx *>(u), n, batch, out4);
    FH_LAUNCH_CHECK("residual_kernel");
    return FH_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Selective iterative refinement, entirely on the device (no host round trip): the batched solver flags the systems
// whose interface-row probe found a defect (status bit 1, value 2); exactly those get one step  r = b - A u,
// A e = r, u += e.  Four stream-ordered launches:
//   1. compact_flagged_kernel   status -> ascending list of flagged systems + their number (one CTA, block scan)
//   2. residual_subset_kernel   r_i = b - A u of system list[i]                (compact rows)
//   3. the batched solver       A(list[i]) e_i = r_i, matrices read in place   (fh_tridiag_batched.cu)
//   4. correct_subset_kernel    u(list[i]) += e_i; status bit 1 cleared, bit 0 of the correction solve merged in
// Launch sizes depend on the capacity of the workspace only; the kernels read the count from device memory.
namespace fh {

int batched_solve_subset(const cplx *dl, const cplx *d, const cplx *du, const cplx *r, cplx *e, int64_t n, int64_t batch,
                         const int32_t *list, const int32_t *count, int64_t cap, int32_t *status, cudaStream_t stream);

constexpr int kCompactThreads = 1024;

// header[0] = number of flagged systems, header[1] = number refined by this call (min(flagged, cap)).  One CTA walks
// status[] in tiles of 1024 consecutive systems (so the list comes out ascending); the flags of 64 tiles are fetched
// up front with coalesced loads and kept as one bit mask per thread, so the tile loop itself touches no global memory:
// ballot inside the warp, one shuffle scan over the 32 warp totals, two block barriers per tile.
__global__ void __launch_bounds__(kCompactThreads)
compact_flagged_kernel(const int32_t *__restrict__ status, int64_t batch, int32_t mask, int32_t cap,
                       int32_t *__restrict__ list, int32_t *__restrict__ header) {
    __shared__ int warp_tot[kCompactThreads / 32];
    __shared__ int warp_off[kCompactThreads / 32 + 1];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int base = 0;
    for (int64_t s0 = 0; s0 < batch; s0 += 64 * kCompactThreads) {
        unsigned long long bits = 0;
#pragma unroll 8
        for (int q = 0; q < 64; ++q) {
            const int64_t s = s0 + (int64_t)q * kCompactThreads + t;
            if (s < batch && (status[s] & mask) != 0) bits |= 1ull << q;
        }
        const int tiles = (int)((batch - s0 + kCompactThreads - 1) / kCompactThreads < 64
                                    ? (batch - s0 + kCompactThreads - 1) / kCompactThreads : 64);
        for (int q = 0; q < tiles; ++q) {
            const int f = (int)((bits >> q) & 1ull);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) warp_tot[warp] = __popc(bal);
            __syncthreads();
            if (warp == 0) {  // exclusive scan of the 32 warp totals
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
            if (f) {
                const int pos = base + warp_off[warp] + __popc(bal & ((1u << lane) - 1u));
                if (pos < cap) list[pos] = (int32_t)(s0 + (int64_t)q * kCompactThreads + t);
            }
            base += warp_off[32];
        }
    }
    if (t == 0) {
        header[0] = base;
        header[1] = base < cap ? base : cap;
    }
}

__global__ void __launch_bounds__(kResThreads)
residual_subset_kernel(const cplx *__restrict__ dl, const cplx *__restrict__ d, const cplx *__restrict__ du,
                       const cplx *__restrict__ b, const cplx *__restrict__ u, int64_t n,
                       const int32_t *__restrict__ list, const int32_t *__restrict__ header, cplx *__restrict__ r) {
    const int64_t total = (int64_t)header[1] * n;
    for (int64_t o = (int64_t)blockIdx.x * kResThreads + threadIdx.x; o < total; o += (int64_t)gridDim.x * kResThreads) {
        const int64_t i = o / n, j = o - i * n;
        const int64_t g = (int64_t)list[i] * n + j;
        cplx v = nmsub(ldg_stream(b + g), ldg_stream(d + g), u[g]);
        if (j > 0) v = nmsub(v, ldg_stream(dl + g), u[g - 1]);
        if (j < n - 1) v = nmsub(v, ldg_stream(du + g), u[g + 1]);
        r[o] = v;
    }
}

__global__ void __launch_bounds__(kResThreads)
correct_subset_kernel(cplx *__restrict__ u, int64_t n, const int32_t *__restrict__ list,
                      const int32_t *__restrict__ header, const cplx *__restrict__ e,
                      const int32_t *__restrict__ status_e, int32_t mask, int32_t *__restrict__ status) {
    const int64_t m = header[1], total = m * n;
    for (int64_t o = (int64_t)blockIdx.x * kResThreads + threadIdx.x; o < total; o += (int64_t)gridDim.x * kResThreads) {
        const int64_t i = o / n, j = o - i * n;
        const int64_t g = (int64_t)list[i] * n + j;
        u[g] = u[g] + ldg_stream(e + o);
        if (j == 0) status[list[i]] = (status[list[i]] & ~mask) | (status_e[i] & 1);
    }
}

struct RefinePlan {  // carving of the workspace
    int64_t cap;
    size_t off_header, off_list, off_status, off_r, off_e, total;
};
static RefinePlan refine_plan(int64_t n, int64_t batch, size_t workspace_bytes) {
    RefinePlan p;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    p.off_header = 0;
    p.off_list = up(2 * sizeof(int32_t));
    p.off_status = p.off_list + up(sizeof(int32_t) * (size_t)batch);
    const size_t fixed = p.off_status + up(sizeof(int32_t) * (size_t)batch);
    const size_t per_system = 2 * sizeof(cplx) * (size_t)n;
    int64_t cap = workspace_bytes > fixed ? (int64_t)((workspace_bytes - fixed) / per_system) : 0;
    if (cap > batch) cap = batch;
    p.cap = cap;
    p.off_r = fixed;
    p.off_e = fixed + sizeof(cplx) * (size_t)n * (size_t)cap;
    p.total = fixed + per_system * (size_t)cap;
    return p;
}
}  // namespace fh

// Workspace for refining up to `max_systems` flagged systems per call.
extern "C" size_t fh_tridiag_refine_flagged_workspace_bytes(int64_t n, int64_t batch, int64_t max_systems) {
    if (n <= 0 || batch <= 0) return 0;
    if (max_systems > batch) max_systems = batch;
    if (max_systems < 1) max_systems = 1;
    const fh::RefinePlan p = fh::refine_plan(n, batch, (size_t)0);
    return p.off_r + 2 * sizeof(fh::cplx) * (size_t)n * (size_t)max_systems;
}

extern "C" int fh_tridiag_refine_flagged_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                              fh_c128 *u, int64_t n, int64_t batch, int32_t *status, void *workspace,
                                              size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && batch >= 0, "n and batch must be >= 0");
    if (n == 0 || batch == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u && status && workspace, "NULL device pointer");
    FH_REQUIRE(batch < (1ll << 31), "batch too large");
    if (n > fh_tridiag_batched_max_n()) {
        set_error("refinement is implemented for the batched solver (n <= %lld)", (long long)fh_tridiag_batched_max_n());
        return FH_E_UNSUPPORTED;
    }
    const RefinePlan p = refine_plan(n, batch, workspace_bytes);
    FH_REQUIRE(p.cap >= 1, "workspace too small (see fh_tridiag_refine_flagged_workspace_bytes)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace);
    int32_t *header = reinterpret_cast<int32_t *>(ws + p.off_header);
    int32_t *list = reinterpret_cast<int32_t *>(ws + p.off_list);
    int32_t *status_e = reinterpret_cast<int32_t *>(ws + p.off_status);
    cplx *r = reinterpret_cast<cplx *>(ws + p.off_r), *e = reinterpret_cast<cplx *>(ws + p.off_e);
    const int32_t mask = 2;  // FH_STATUS_SMALL_PIVOT
    compact_flagged_kernel<<<1, kCompactThreads, 0, st>>>(status, batch, mask, (int32_t)p.cap, list, header);
    FH_LAUNCH_CHECK("compact_flagged_kernel");
    const unsigned grid = (unsigned)(8 * sm_count());
    residual_subset_kernel<<<grid, kResThreads, 0, st>>>(
        reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d), reinterpret_cast<const cplx *>(du),
        reinterpret_cast<const cplx *>(b), reinterpret_cast<const cplx *>(u), n, list, header, r);
    FH_LAUNCH_CHECK("residual_subset_kernel");
    int rc = batched_solve_subset(reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d),
                                  reinterpret_cast<const cplx *>(du), r, e, n, batch, list, header + 1, p.cap, status_e, st);
    if (rc != FH_OK) return rc;
    correct_subset_kernel<<<grid, kResThreads, 0, st>>>(reinterpret_cast<cplx *>(u), n, list, header, e, status_e, mask,
                                                       status);
    FH_LAUNCH_CHECK("correct_subset_kernel");
    return FH_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Pivoted fallback.  The fast solvers are unpivoted; a system on which they produce a non-finite value (an
// exactly vanishing leading minor, e.g. k*h = 2 on the first element) is re-solved here the way LAPACK zgtsv
// does it: Gaussian elimination with partial pivoting between adjacent rows, one thread per system, serial in
// n.  It only ever sees the handful of systems the status flags name, so its speed does not matter; what
// matters is that numpy.linalg.LinAlgError is raised for genuinely singular systems only, as the reference does.
namespace fh {
__global__ void pivoted_gtsv_kernel(const cplx *__restrict__ dl, const cplx *__restrict__ d, const cplx *__restrict__ du,
                                    const cplx *__restrict__ b, cplx *__restrict__ u, int64_t n, int64_t batch,
                                    cplx *__restrict__ work, int32_t *__restrict__ status) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= batch) return;
    const cplx *L = dl + s * n, *D0 = d + s * n, *U0 = du + s * n, *B = b + s * n;
    cplx *x = u + s * n;                 // right-hand side, then the solution
    cplx *dd = work + (3 * s) * n;       // diagonal
    cplx *u1 = dd + n, *u2 = u1 + n;     // first and second super-diagonal of the factor
    for (int64_t i = 0; i < n; ++i) dd[i] = D0[i], u1[i] = i < n - 1 ? U0[i] : czero(), u2[i] = czero(), x[i] = B[i];
    bool singular = false;
    for (int64_t i = 0; i + 1 < n; ++i) {
        const cplx sub = L[i + 1];
        if (abs1c(dd[i]) >= abs1c(sub)) {  // no interchange
            if (abs1c(dd[i]) == 0.0) { singular = true; break; }
            const cplx m = sub * crecip(dd[i]);
            dd[i + 1] = nmsub(dd[i + 1], m, u1[i]);
            x[i + 1] = nmsub(x[i + 1], m, x[i]);
        } else {                           // swap rows i and i+1
            const cplx m = dd[i] * crecip(sub);
            const cplx di1 = dd[i + 1], ui = u1[i], xi = x[i];
            dd[i] = sub;
            u1[i] = di1;
            dd[i + 1] = nmsub(ui, m, di1);
            if (i + 2 < n) {
                u2[i] = u1[i + 1];
                u1[i + 1] = -(m * u1[i + 1]);
            }
            x[i] = x[i + 1];
            x[i + 1] = nmsub(xi, m, x[i + 1]);
        }
    }
    if (!singular && abs1c(dd[n - 1]) == 0.0) singular = true;
    if (!singular) {
        x[n - 1] = x[n - 1] * crecip(dd[n - 1]);
        if (n > 1) x[n - 2] = nmsub(x[n - 2], u1[n - 2], x[n - 1]) * crecip(dd[n - 2]);
        for (int64_t i = n - 3; i >= 0; --i)
            x[i] = nmsub(nmsub(x[i], u1[i], x[i + 1]), u2[i], x[i + 2]) * crecip(dd[i]);
        for (int64_t i = 0; i < n && !singular; ++i) singular = !cfinite(x[i]);
    }
    if (status != nullptr) status[s] = singular ? 1 : 0;
}
}  // namespace fh

extern "C" size_t fh_tridiag_solve_pivoted_workspace_bytes(int64_t n, int64_t batch) {
    return sizeof(fh::cplx) * 3 * (size_t)n * (size_t)batch;
}

extern "C" int fh_tridiag_solve_pivoted_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                             fh_c128 *u, int64_t n, int64_t batch, int32_t *status, void *workspace,
                                             size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && batch >= 0, "n and batch must be >= 0");
    if (n == 0 || batch == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u && workspace, "NULL device pointer");
    FH_REQUIRE(workspace_bytes >= fh_tridiag_solve_pivoted_workspace_bytes(n, batch), "workspace too small");
    pivoted_gtsv_kernel<<<(unsigned)((batch + 31) / 32), 32, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d), reinterpret_cast<const cplx *>(du),
        reinterpret_cast<const cplx *>(b), reinterpret_cast<cplx *>(u), n, batch, reinterpret_cast<cplx *>(workspace),
        status);
    FH_LAUNCH_CHECK("pivoted_gtsv_kernel");
    return FH_OK;
}
