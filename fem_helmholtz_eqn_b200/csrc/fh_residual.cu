// Residual / parity metrics of a batch of tridiagonal systems, on device (gate G3):
// out[s] = { ||b - A u||_2^2, ||u||_2^2, ||b||_2^2, ||A||_inf }.  One CTA per system (grid-stride),
// coalesced streaming reads (96 B per row), fixed-shape tree reductions => deterministic results.
#include "fh_helm1d.cuh"

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
// The same metrics for GENERATED rows (the fused / partitioned solvers never store A): rows [row_begin, row_begin + n_rows)
// of the Helmholtz system of wavenumber k are re-generated from (h, k) -- bit-identical to fh_assemble_1d_c128 -- and
// applied to the slab of the solution; 16 B per row of HBM traffic (u is read once; the neighbours come from L1/L2).
// The two unknowns next to the slab (previous rank's last, next rank's first) are passed as pointers.  Every CTA leaves
// its partial sums in the workspace, the last CTA to finish folds them in a fixed order (deterministic) into
// out4 = { sum |b - A u|^2, sum |u|^2, sum |b|^2, max row sum of |A| } of this slab: partial sums of a rank; the caller
// adds (max for the fourth) over the ranks.
namespace fh {
constexpr int kGenResCtasMax = 2048;
struct GenResWs {
    double part[kGenResCtasMax][4];
    unsigned int done;
};

__global__ void __launch_bounds__(kResThreads)
residual_generated_kernel(Problem1d p, double k, double k2, const double *__restrict__ kk, int64_t row_begin,
                          int64_t n_rows, const cplx *__restrict__ u, const cplx *__restrict__ u_left,
                          const cplx *__restrict__ u_right, GenResWs *__restrict__ ws, double *__restrict__ out4) {
    __shared__ double red[4][kResThreads / 32];
    __shared__ bool last;
    if (kk != nullptr) k = __ldg(kk), k2 = __ldg(kk + 1);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool scalar = p.h_mode == FH_H_SCALAR && p.n_elem >= 2;
    cplx il = czero(), id = czero(), iu = czero(), ib = czero();
    if (scalar) row_values(p, row_geom(p, 1), 1, k, k2, il, id, iu, ib);
    const double irow = hypot(il.re, il.im) + hypot(id.re, id.im) + hypot(iu.re, iu.im);
    double r2 = 0.0, u2 = 0.0, b2 = 0.0, amax = 0.0;
    for (int64_t o = (int64_t)blockIdx.x * kResThreads + t; o < n_rows; o += (int64_t)gridDim.x * kResThreads) {
        const int64_t j = row_begin + o;
        cplx lo = il, di = id, hi = iu, rh = ib;
        double rowsum = irow;
        if (!scalar || j == 0 || j == p.n_elem) {
            row_values(p, row_geom(p, j), j, k, k2, lo, di, hi, rh);
            rowsum = hypot(di.re, di.im) + (j > 0 ? hypot(lo.re, lo.im) : 0.0) + (j < p.n_elem ? hypot(hi.re, hi.im) : 0.0);
        }
        const cplx uj = u[o];
        cplx r = nmsub(rh, di, uj);
        if (j > 0) r = nmsub(r, lo, o > 0 ? u[o - 1] : (u_left != nullptr ? *u_left : czero()));
        if (j < p.n_elem) r = nmsub(r, hi, o + 1 < n_rows ? u[o + 1] : (u_right != nullptr ? *u_right : czero()));
        r2 += r.re * r.re + r.im * r.im;
        u2 += uj.re * uj.re + uj.im * uj.im;
        b2 += rh.re * rh.re + rh.im * rh.im;
        amax = fmax(amax, rowsum);
    }
    r2 = warp_sum(r2), u2 = warp_sum(u2), b2 = warp_sum(b2), amax = warp_max(amax);
    if (lane == 0) red[0][warp] = r2, red[1][warp] = u2, red[2][warp] = b2, red[3][warp] = amax;
    __syncthreads();
    if (t == 0) {
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        for (int w = 0; w < kResThreads / 32; ++w) v0 += red[0][w], v1 += red[1][w], v2 += red[2][w], v3 = fmax(v3, red[3][w]);
        double *mine = ws->part[blockIdx.x];
        mine[0] = v0, mine[1] = v1, mine[2] = v2, mine[3] = v3;
        __threadfence();
        last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && t == 0) {  // (<= 2048 partial sums: a fixed-order serial fold keeps the result reproducible)
        __threadfence();
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        for (unsigned c = 0; c < gridDim.x; ++c) {
            const volatile double *q = ws->part[c];
            v0 += q[0], v1 += q[1], v2 += q[2], v3 = fmax(v3, q[3]);
        }
        out4[0] = v0, out4[1] = v1, out4[2] = v2, out4[3] = v3;
        ws->done = 0;  // ready for the next launch (graph replays)
    }
}
}  // namespace fh

extern "C" size_t fh_helmholtz1d_residual_workspace_bytes(void) { return sizeof(fh::GenResWs); }

extern "C" int fh_helmholtz1d_residual_c128(const fh_problem1d *prob_host, double k, double k2, const double *k_k2_dev,
                                            int64_t row_begin, int64_t n_rows, const fh_c128 *u, const fh_c128 *u_left,
                                            const fh_c128 *u_right, double *out4, void *workspace,
                                            size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    Problem1d p;
    int rc = problem_from_host(prob_host, p);
    if (rc != FH_OK) return rc;
    FH_REQUIRE(row_begin >= 0 && n_rows >= 1 && row_begin + n_rows <= p.n_elem + 1,
               "row range [%lld, %lld) outside the mesh of %lld nodes", (long long)row_begin,
               (long long)(row_begin + n_rows), (long long)(p.n_elem + 1));
    FH_REQUIRE(u && out4 && workspace, "NULL device pointer");
    FH_REQUIRE(workspace_bytes >= sizeof(GenResWs), "workspace too small (fh_helmholtz1d_residual_workspace_bytes)");
    FH_REQUIRE(row_begin == 0 || u_left != nullptr, "u_left is needed unless the slab starts at mesh row 0");
    FH_REQUIRE(row_begin + n_rows == p.n_elem + 1 || u_right != nullptr, "u_right is needed unless the slab ends the mesh");
    int64_t grid = 8ll * sm_count();
    const int64_t need = (n_rows + kResThreads - 1) / kResThreads;
    if (grid > need) grid = need;
    if (grid > kGenResCtasMax) grid = kGenResCtasMax;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GenResWs *ws = static_cast<GenResWs *>(workspace);
    // (the counter is zero after every completed launch; the first use of a fresh workspace needs it cleared)
    FH_CUDA(cudaMemsetAsync(&ws->done, 0, sizeof(unsigned int), st));
    residual_generated_kernel<<<(unsigned)grid, kResThreads, 0, st>>>(
        p, k, k2, k_k2_dev, row_begin, n_rows, reinterpret_cast<const cplx *>(u), reinterpret_cast<const cplx *>(u_left),
        reinterpret_cast<const cplx *>(u_right), ws, out4);
    FH_LAUNCH_CHECK("residual_generated_kernel");
    return FH_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Selective iterative refinement, entirely on the device (no host round trip): the batched solver flags the systems
// whose interface-row probe found a defect (status bit 1, value 2); exactly those get one step  r = b - A u,
// A e = r, u += e.  Two stream-ordered launches:
//   1. compact_flagged_kernel   status -> ascending list of flagged systems + their number (one CTA, ballots + scan)
//   2. the batched solver as tridiag_pipe_kernel<., BandRefineSource> (fh_tridiag_pipe.cuh): the residual is formed
//      while the rows of system list[i] are read, the correction is added to u by a TMA reduce-add, status bit 1 is
//      cleared and bit 0 of the correction solve merged in -- 96 B of traffic per refined row.
// Round 1 / early round 2 ran this as four launches (compaction, a residual kernel into a compact r, the solver on
// compact r -> e, an update kernel: 240 B per refined row, 0.32 ms for the 1,732 flagged systems of the 65,536-wavenumber
// sweep against 0.19 ms now, tools/refine_probe.py) -- the same bits (tests/test_gpu_1d.py spells the step out through
// the C ABI).
// Launch sizes depend on the batch only; the kernels read the count from device memory.
namespace fh {

int batched_refine_subset(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n, int64_t batch,
                          const int32_t *list, const int32_t *count, int64_t cap, int32_t *status, cudaStream_t stream);

constexpr int kCompactThreads = 1024;
constexpr int kCompactWords = 2;  // 32-system flag words per thread and tile: a tile is 65,536 systems

// header[0] = number of flagged systems, header[1] = number refined by this call (min(flagged, cap)).  One CTA walks
// the status array in tiles of 65,536 systems: every warp reads its share with coalesced loads (32 consecutive systems
// per instruction, eight instructions in flight) and turns each into one ballot word in shared memory; every thread then
// owns two consecutive words, counts their bits, one block-wide exclusive scan places the threads, and the set bits are
// written out as indices -- the list comes out ascending.  Three barriers per tile.  (The first versions read
// status[] one int per thread and run: 52, then 46 us for 65,536 systems, a seventh of the whole refinement.)
__global__ void __launch_bounds__(kCompactThreads)
compact_flagged_kernel(const int32_t *__restrict__ status, int64_t batch, int32_t mask, int32_t cap,
                       int32_t *__restrict__ list, int32_t *__restrict__ header) {
    __shared__ uint32_t bits[kCompactThreads * kCompactWords];
    __shared__ int warp_tot[kCompactThreads / 32];
    __shared__ int tile_total;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int kWarpWords = 32 * kCompactWords;
    constexpr int64_t kTile = (int64_t)kCompactThreads * kCompactWords * 32;
    int running = 0;
    for (int64_t base = 0; base < batch; base += kTile) {
#pragma unroll 8
        for (int q = 0; q < kWarpWords; ++q) {
            const int64_t s = base + ((int64_t)(warp * kWarpWords + q) << 5) + lane;
            const bool f = s < batch && (status[s] & mask) != 0;
            const uint32_t w = __ballot_sync(0xffffffffu, f);
            if (lane == (q & 31)) bits[warp * kWarpWords + q] = w;
        }
        __syncwarp();
        uint32_t w[kCompactWords];
        int count = 0;
#pragma unroll
        for (int j = 0; j < kCompactWords; ++j) w[j] = bits[t * kCompactWords + j], count += __popc(w[j]);
        int incl = count;  // inclusive scan inside the warp ...
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {  // ... and over the 32 warp totals
            const int v = warp_tot[lane];
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int q = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += q;
            }
            warp_tot[lane] = x - v;  // exclusive
            if (lane == 31) tile_total = x;
        }
        __syncthreads();
        int pos = running + warp_tot[warp] + incl - count;
#pragma unroll
        for (int j = 0; j < kCompactWords; ++j) {
            const int64_t s0 = base + ((int64_t)(t * kCompactWords + j) << 5);
            for (uint32_t m = w[j]; m != 0; m &= m - 1, ++pos)
                if (pos < cap) list[pos] = (int32_t)(s0 + __ffs((int)m) - 1);
        }
        running += tile_total;
        __syncthreads();  // (warp_tot / tile_total are rewritten by the next tile)
    }
    if (t == 0) {
        header[0] = running;
        header[1] = running < cap ? running : cap;
    }
}

struct RefinePlan {  // carving of the workspace: {flagged, refined} + the list of flagged systems
    int64_t cap;
    size_t off_header, off_list, total;
};
static RefinePlan refine_plan(int64_t batch, size_t workspace_bytes) {
    RefinePlan p;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    p.off_header = 0;
    p.off_list = up(2 * sizeof(int32_t));
    p.total = p.off_list + up(sizeof(int32_t) * (size_t)batch);
    p.cap = workspace_bytes >= p.total ? batch : 0;  // every flagged system fits
    return p;
}
}  // namespace fh

// Workspace for refining up to `max_systems` flagged systems per call.
extern "C" size_t fh_tridiag_refine_flagged_workspace_bytes(int64_t n, int64_t batch, int64_t max_systems) {
    (void)n, (void)max_systems;  // (what the four-launch version sized its r / e arrays by)
    if (batch <= 0) return 0;
    return fh::refine_plan(batch, (size_t)0).total;
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
    const RefinePlan p = refine_plan(batch, workspace_bytes);
    FH_REQUIRE(p.cap >= 1, "workspace too small (see fh_tridiag_refine_flagged_workspace_bytes)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace);
    int32_t *header = reinterpret_cast<int32_t *>(ws + p.off_header);
    int32_t *list = reinterpret_cast<int32_t *>(ws + p.off_list);
    const int32_t mask = 2;  // FH_STATUS_SMALL_PIVOT
    compact_flagged_kernel<<<1, kCompactThreads, 0, st>>>(status, batch, mask, (int32_t)p.cap, list, header);
    FH_LAUNCH_CHECK("compact_flagged_kernel");
    return batched_refine_subset(reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d),
                                 reinterpret_cast<const cplx *>(du), reinterpret_cast<const cplx *>(b),
                                 reinterpret_cast<cplx *>(u), n, batch, list, header + 1, p.cap, status, st);
}

// ---------------------------------------------------------------------------------------------------------
// Pivoted fallback.  The fast solvers are unpivoted; a system on which they break down (an exactly vanishing
// leading minor, e.g. k*h = 2 on the first element, or a resonance of the partition that refinement does not repair)
// is re-solved here the way LAPACK zgtsv does it: Gaussian elimination with partial pivoting between adjacent rows.
// That recurrence is serial in n, so the kernel is built around its dependent chain:
//   * one CTA per system; the elimination runs from BOTH ends at once ("burn at both ends"): lane 0 of warp 0 sweeps
//     rows 0 .. m downwards, lane 0 of warp 1 rows n-1 .. m+1 upwards (the reversed system is tridiagonal too, with the
//     sub- and super-diagonal exchanged); the two sweeps meet in a 2 x 2 system for (x_m, x_{m+1}), solved with
//     pivoting, and both lanes back-substitute outwards.  Half the chain length, no loss of robustness.
//   * the other threads are the memory system of the two sweepers: rows arrive in shared-memory tiles by cp.async one
//     tile ahead (coalesced 16-byte requests), so the sweepers never wait on a global load; the factor rows
//     (1/pivot, u1, u2, rhs -- 64 B per row) leave by plain stores to a per-CTA workspace slot and come back through
//     the same tiles for the back-substitution.
// Per row the chain is  |p| compare -> 1/p -> multiplier -> new diagonal  (~110 cycles); n = 4097: ~0.16 ms per system
// (the one-thread-per-system kernel of round 1 took 5 ms), any n (tiles stream: n = 1e6 takes ~40 ms).
// What matters beyond speed is that numpy.linalg.LinAlgError is raised for genuinely singular systems only, as the
// reference does (nb:204): a zero pivot after the interchange gives Inf -> NaN, reported in status.
namespace fh {

constexpr int kPivThreads = 256;
constexpr int kPivTile = 128;  // rows per side and tile
constexpr int kPivSlotsMax = 256;

struct PivSmem {
    cplx tile[2][2][kPivTile][4];  // [side][buffer][row][4 numbers]  (forward: lo, di, hi, rhs; backward: the factor row)
    cplx mid[2][3];                // the two modified rows that meet in the middle: (diag, off-diagonal, rhs)
    cplx xmid[2];                  // x_m, x_{m+1}
    double red[4][kPivThreads / 32];
    int flag;
};

// normwise backward error ingredients of one system, by the whole CTA (same arithmetic as residual_kernel)
__device__ __forceinline__ void cta_residual(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, const cplx *u,
                                             int64_t n, double (*red)[kPivThreads / 32], double out[4]) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double r2 = 0.0, u2 = 0.0, b2 = 0.0, amax = 0.0;
    for (int64_t j = t; j < n; j += kPivThreads) {
        const cplx uj = u[j], bj = ldg_stream(b + j), dj = ldg_stream(d + j);
        cplx r = nmsub(bj, dj, uj);
        double rowsum = hypot(dj.re, dj.im);
        if (j > 0) {
            const cplx lj = ldg_stream(dl + j);
            r = nmsub(r, lj, u[j - 1]);
            rowsum += hypot(lj.re, lj.im);
        }
        if (j < n - 1) {
            const cplx hj = ldg_stream(du + j);
            r = nmsub(r, hj, u[j + 1]);
            rowsum += hypot(hj.re, hj.im);
        }
        r2 += r.re * r.re + r.im * r.im;
        u2 += uj.re * uj.re + uj.im * uj.im;
        b2 += bj.re * bj.re + bj.im * bj.im;
        amax = fmax(amax, rowsum);
    }
    r2 = warp_sum(r2), u2 = warp_sum(u2), b2 = warp_sum(b2), amax = warp_max(amax);
    __syncthreads();  // (red[] may still be read from a previous call)
    if (lane == 0) red[0][warp] = r2, red[1][warp] = u2, red[2][warp] = b2, red[3][warp] = amax;
    __syncthreads();
    constexpr int nw = kPivThreads / 32;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
    for (int w = 0; w < nw; ++w) v0 += red[0][w], v1 += red[1][w], v2 += red[2][w], v3 = fmax(v3, red[3][w]);
    out[0] = v0, out[1] = v1, out[2] = v2, out[3] = v3;  // identical on every thread
}

// One system, one CTA.  L, D, U, B: the system's bands and right-hand side (global); x: the solution (global, may not
// alias B); fac: workspace slot of n * 4 complex numbers.  Returns true (on every thread) if the system is singular to
// working precision (a non-finite value in the solution).
__device__ bool cta_pivoted_solve(const cplx *L, const cplx *D, const cplx *U, const cplx *B, cplx *x, int64_t n,
                                  cplx *fac, PivSmem &sm) {
    const int t = threadIdx.x;
    if (n == 1) {
        if (t == 0) {
            const cplx v = B[0] * crecip(D[0]);
            x[0] = v;
            sm.flag = cfinite(v) ? 0 : 1;
        }
        __syncthreads();
        const bool bad1 = sm.flag != 0;
        __syncthreads();
        return bad1;
    }
    // rows consumed by the two sweeps: top 0 .. m (m + 1 rows), bottom n-1 .. m+1 (mp + 1 rows); they eliminate
    // x_0 .. x_{m-1} and x_{n-1} .. x_{m+2}
    const int64_t m = (n - 2) / 2, mp = n - 2 - m;
    const int64_t cnt[2] = {m + 1, mp + 1};
    const int64_t tiles = ((cnt[0] > cnt[1] ? cnt[0] : cnt[1]) + kPivTile - 1) / kPivTile;
    // side s, local row j <-> global row (s == 0 ? j : n - 1 - j); local "lo" couples to local row j - 1
    auto load_tile = [&](int64_t it, int buf) {
        for (int e = t; e < 2 * kPivTile; e += kPivThreads) {
            const int s = e / kPivTile, r = e - s * kPivTile;
            const int64_t j = it * kPivTile + r;
            if (j < cnt[s]) {
                const int64_t g = s == 0 ? j : n - 1 - j;
                cp_async16(&sm.tile[s][buf][r][0], (s == 0 ? L : U) + g);
                cp_async16(&sm.tile[s][buf][r][1], D + g);
                cp_async16(&sm.tile[s][buf][r][2], (s == 0 ? U : L) + g);
                cp_async16(&sm.tile[s][buf][r][3], B + g);
            }
        }
        cp_async_commit();
    };
    const bool sweeper = (t & 31) == 0 && t < 64;
    const int side = t >> 5;  // (meaningful for the sweepers)
    cplx dd = czero(), u1 = czero(), rr = czero();  // the sweeper's current row: dd x_j + u1 x_{j+1} = rr
    load_tile(0, 0);
    for (int64_t it = 0; it < tiles; ++it) {
        const int buf = (int)(it & 1);
        cp_async_wait_all();
        __syncthreads();  // tile `it` has landed; everybody is done with the other buffer
        if (it + 1 < tiles) load_tile(it + 1, buf ^ 1);
        if (sweeper) {
            const int64_t j0 = it * kPivTile;
            int64_t j1 = j0 + kPivTile;
            if (j1 > cnt[side]) j1 = cnt[side];
            const cplx(*rows)[4] = sm.tile[side][buf];
            cplx nlo = rows[0][0], ndi = rows[0][1], nhi = rows[0][2], nrh = rows[0][3];
            for (int64_t j = j0; j < j1; ++j) {
                const cplx lo = nlo, di = ndi, hi = nhi, rh = nrh;
                if (j + 1 < j1) {  // the next row's reads are issued ahead of this row's dependent chain
                    const int r = (int)(j + 1 - j0);
                    nlo = rows[r][0], ndi = rows[r][1], nhi = rows[r][2], nrh = rows[r][3];
                }
                if (j == 0) {  // the first row of the sweep: nothing above it
                    dd = di, u1 = hi, rr = rh;
                    continue;
                }
                // eliminate x_{j-1} between the current row (dd, u1, 0 | rr) and the next one (lo, di, hi | rh):
                // the row with the larger leading entry (LAPACK's cabs1) is the pivot row
                const bool swap = abs1c(dd) < abs1c(lo);
                const cplx p0 = swap ? lo : dd, p1 = swap ? di : u1, p2 = swap ? hi : czero(), pr = swap ? rh : rr;
                const cplx o0 = swap ? dd : lo, o1 = swap ? u1 : di, o2 = swap ? czero() : hi, orr = swap ? rr : rh;
                const cplx inv = frecip(p0);
                const cplx mlt = o0 * inv;
                cplx *f = fac + 4 * (side == 0 ? j - 1 : n - j);  // factor row of the eliminated unknown
                f[0] = inv, f[1] = p1, f[2] = p2, f[3] = pr;
                dd = nmsub(o1, mlt, p1);
                u1 = nmsub(o2, mlt, p2);
                rr = nmsub(orr, mlt, pr);
            }
        }
    }
    if (sweeper) sm.mid[side][0] = dd, sm.mid[side][1] = u1, sm.mid[side][2] = rr;
    __syncthreads();
    if (t == 0) {
        // top:    a00 x_m + a01 x_{m+1} = r0      bottom (reversed): a11 x_{m+1} + a10 x_m = r1
        cplx a00 = sm.mid[0][0], a01 = sm.mid[0][1], r0 = sm.mid[0][2];
        cplx a11 = sm.mid[1][0], a10 = sm.mid[1][1], r1 = sm.mid[1][2];
        cplx xm, xm1;
        if (abs1c(a00) >= abs1c(a10)) {
            const cplx inv = crecip(a00), w = a10 * inv;
            xm1 = nmsub(r1, w, r0) * crecip(nmsub(a11, w, a01));
            xm = nmsub(r0, a01, xm1) * inv;
        } else {
            const cplx inv = crecip(a10), w = a00 * inv;
            xm1 = nmsub(r0, w, r1) * crecip(nmsub(a01, w, a11));
            xm = nmsub(r1, a11, xm1) * inv;
        }
        sm.xmid[0] = xm, sm.xmid[1] = xm1;
        x[m] = xm, x[m + 1] = xm1;
        sm.flag = (cfinite(xm) && cfinite(xm1)) ? 0 : 1;
    }
    __syncthreads();  // (also: the factor rows written by the sweepers are visible to the whole CTA)
    // ---- back-substitution, outwards from the middle: local rows cnt - 2 .. 0 of each side ----
    auto load_fac = [&](int64_t it, int buf) {
        for (int e = t; e < 2 * kPivTile * 4; e += kPivThreads) {
            const int s = e / (4 * kPivTile), q = e - s * 4 * kPivTile, r = q >> 2, c = q & 3;
            const int64_t j = it * kPivTile + r;
            if (j + 1 < cnt[s]) cp_async16(&sm.tile[s][buf][r][c], fac + 4 * (s == 0 ? j : n - 1 - j) + c);
        }
        cp_async_commit();
    };
    const int64_t ftiles = ((cnt[0] > cnt[1] ? cnt[0] : cnt[1]) - 1 + kPivTile - 1) / kPivTile;
    cplx xa = czero(), xb = czero();  // the two unknowns beyond the current one (local j + 1, j + 2)
    if (sweeper) xa = sm.xmid[side], xb = sm.xmid[side ^ 1];
    bool bad = false;
    if (ftiles > 0) load_fac(ftiles - 1, (int)((ftiles - 1) & 1));
    for (int64_t it = ftiles - 1; it >= 0; --it) {
        const int buf = (int)(it & 1);
        cp_async_wait_all();
        __syncthreads();
        if (it > 0) load_fac(it - 1, buf ^ 1);
        if (sweeper) {
            const int64_t j0 = it * kPivTile;
            int64_t j1 = j0 + kPivTile;
            if (j1 > cnt[side] - 1) j1 = cnt[side] - 1;
            const cplx(*rows)[4] = sm.tile[side][buf];
            cplx ninv = czero(), nf1 = czero(), nf2 = czero(), nfr = czero();
            if (j1 > j0) ninv = rows[j1 - 1 - j0][0], nf1 = rows[j1 - 1 - j0][1], nf2 = rows[j1 - 1 - j0][2], nfr = rows[j1 - 1 - j0][3];
            for (int64_t j = j1 - 1; j >= j0; --j) {
                const cplx inv = ninv, f1 = nf1, f2 = nf2, fr = nfr;
                if (j > j0) {
                    const int r = (int)(j - 1 - j0);
                    ninv = rows[r][0], nf1 = rows[r][1], nf2 = rows[r][2], nfr = rows[r][3];
                }
                const cplx v = nmsub(nmsub(fr, f2, xb), f1, xa) * inv;
                bad |= !cfinite(v);
                x[side == 0 ? j : n - 1 - j] = v;
                xb = xa, xa = v;
            }
        }
    }
    if (sweeper && bad) atomicOr(&sm.flag, 1);
    __syncthreads();
    const bool singular = sm.flag != 0;
    __syncthreads();
    return singular;
}

// Direct entry: every system of the batch (grid-stride over the workspace slots).
__global__ void __launch_bounds__(kPivThreads)
pivoted_gtsv_kernel(const cplx *__restrict__ dl, const cplx *__restrict__ d, const cplx *__restrict__ du,
                    const cplx *__restrict__ b, cplx *__restrict__ u, int64_t n, int64_t batch, cplx *__restrict__ work,
                    int32_t *__restrict__ status) {
    __shared__ PivSmem sm;
    cplx *fac = work + (size_t)blockIdx.x * 4 * (size_t)n;
    for (int64_t s = blockIdx.x; s < batch; s += gridDim.x) {
        const bool singular = cta_pivoted_solve(dl + s * n, d + s * n, du + s * n, b + s * n, u + s * n, n, fac, sm);
        if (threadIdx.x == 0 && status != nullptr) status[s] = singular ? 1 : 0;
    }
}

// Repair of the systems whose status carries the breakdown bit (bit 0), decided and carried out on the device.  Every
// CTA scans its own contiguous share of status[] (no compaction pass, no list: on the usual step nothing is flagged and
// the kernel is a 256 KB read) and deals with what it finds there:
//   per system: normwise backward error of the solution that is there (refinement may already have repaired it);
//               finite and <= accept_tol -> kept, bit 0 cleared;   otherwise -> re-solved with partial pivoting,
//               status = 1 only if that solution is not finite either (singular to working precision)
//   header[0] += flagged, header[2] += kept, header[3] += re-solved, header[4] += singular
__global__ void __launch_bounds__(kPivThreads)
repair_flagged_kernel(const cplx *__restrict__ dl, const cplx *__restrict__ d, const cplx *__restrict__ du,
                      const cplx *__restrict__ b, cplx *__restrict__ u, int64_t n, int64_t batch, int64_t per_cta,
                      int32_t *__restrict__ header, double accept_tol, cplx *__restrict__ work,
                      int32_t *__restrict__ status) {
    __shared__ PivSmem sm;
    __shared__ int found[kPivThreads], sorted[kPivThreads];
    __shared__ int n_found;
    cplx *fac = work + (size_t)blockIdx.x * 4 * (size_t)n;
    const int64_t s_begin = (int64_t)blockIdx.x * per_cta;
    const int64_t s_end = s_begin + per_cta < batch ? s_begin + per_cta : batch;
    for (int64_t s0 = s_begin; s0 < s_end; s0 += kPivThreads) {
        const int64_t mine = s0 + threadIdx.x;
        const bool flagged = mine < s_end && (status[mine] & 1) != 0;
        if (__syncthreads_or(flagged ? 1 : 0) == 0) continue;   // (the usual case)
        if (threadIdx.x == 0) n_found = 0;
        __syncthreads();
        if (flagged) found[atomicAdd(&n_found, 1)] = (int)(mine - s0);
        __syncthreads();
        const int count = n_found;
        if ((int)threadIdx.x < count) {   // ascending order, whatever order the atomics came in
            const int v = found[threadIdx.x];
            int rank = 0;
            for (int w = 0; w < count; ++w) rank += found[w] < v ? 1 : 0;
            sorted[rank] = v;
        }
        __syncthreads();
        for (int i = 0; i < count; ++i) {
            const int64_t s = s0 + sorted[i], o = s * n;
            double m4[4];
            cta_residual(dl + o, d + o, du + o, b + o, u + o, n, sm.red, m4);
            const double be = sqrt(m4[0]) / (m4[3] * sqrt(m4[1]) + sqrt(m4[2]));
            if (threadIdx.x == 0) atomicAdd(&header[0], 1);
            if (be <= accept_tol) {  // (false for NaN)
                if (threadIdx.x == 0) {
                    status[s] &= ~1;
                    atomicAdd(&header[2], 1);
                }
                continue;
            }
            const bool singular = cta_pivoted_solve(dl + o, d + o, du + o, b + o, u + o, n, fac, sm);
            if (threadIdx.x == 0) {
                status[s] = singular ? 1 : 0;
                atomicAdd(&header[3], 1);
                if (singular) atomicAdd(&header[4], 1);
            }
        }
        __syncthreads();
    }
}

static int64_t pivoted_slots(int64_t n, int64_t batch) {
    int64_t slots = batch < kPivSlotsMax ? batch : kPivSlotsMax;
    const int64_t by_mem = (int64_t)((2ull << 30) / (sizeof(cplx) * 4 * (size_t)n));  // <= 2 GB of factor rows
    if (slots > by_mem) slots = by_mem;
    return slots < 1 ? 1 : slots;
}
}  // namespace fh

extern "C" size_t fh_tridiag_solve_pivoted_workspace_bytes(int64_t n, int64_t batch) {
    if (n <= 0 || batch <= 0) return 0;
    return sizeof(fh::cplx) * 4 * (size_t)n * (size_t)fh::pivoted_slots(n, batch);
}

extern "C" int fh_tridiag_solve_pivoted_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                             fh_c128 *u, int64_t n, int64_t batch, int32_t *status, void *workspace,
                                             size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && batch >= 0, "n and batch must be >= 0");
    if (n == 0 || batch == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u && workspace, "NULL device pointer");
    FH_REQUIRE(u != b, "the solution must not alias the right-hand side");
    int64_t slots = (int64_t)(workspace_bytes / (sizeof(cplx) * 4 * (size_t)n));
    FH_REQUIRE(slots >= 1, "workspace too small (see fh_tridiag_solve_pivoted_workspace_bytes)");
    if (slots > batch) slots = batch;
    if (slots > kPivSlotsMax) slots = kPivSlotsMax;
    pivoted_gtsv_kernel<<<(unsigned)slots, kPivThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d), reinterpret_cast<const cplx *>(du),
        reinterpret_cast<const cplx *>(b), reinterpret_cast<cplx *>(u), n, batch, reinterpret_cast<cplx *>(workspace),
        status);
    FH_LAUNCH_CHECK("pivoted_gtsv_kernel");
    return FH_OK;
}

namespace fh {
struct RepairPlan {
    int64_t slots;
    size_t off_header, off_list, off_work, total;
};
static RepairPlan repair_plan(int64_t n, int64_t batch, int64_t slots) {
    RepairPlan p;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    p.slots = slots;
    p.off_header = 0;
    p.off_list = up(8 * sizeof(int32_t));
    p.off_work = p.off_list + up(sizeof(int32_t) * (size_t)batch);
    p.total = p.off_work + sizeof(cplx) * 4 * (size_t)n * (size_t)slots;
    return p;
}
}  // namespace fh

// Workspace for repairing with up to `max_concurrent` systems in flight (every one of them is repaired either way).
extern "C" size_t fh_tridiag_repair_flagged_workspace_bytes(int64_t n, int64_t batch, int64_t max_concurrent) {
    if (n <= 0 || batch <= 0) return 0;
    if (max_concurrent < 1) max_concurrent = 1;
    const int64_t cap = fh::pivoted_slots(n, batch);
    return fh::repair_plan(n, batch, max_concurrent < cap ? max_concurrent : cap).total;
}

extern "C" int fh_tridiag_repair_flagged_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                              fh_c128 *u, int64_t n, int64_t batch, int32_t *status, double accept_tol,
                                              void *workspace, size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && batch >= 0, "n and batch must be >= 0");
    if (n == 0 || batch == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u && status && workspace, "NULL device pointer");
    FH_REQUIRE(batch < (1ll << 31), "batch too large");
    const RepairPlan fixed = repair_plan(n, batch, 0);
    FH_REQUIRE(workspace_bytes > fixed.total, "workspace too small (see fh_tridiag_repair_flagged_workspace_bytes)");
    int64_t slots = (int64_t)((workspace_bytes - fixed.total) / (sizeof(cplx) * 4 * (size_t)n));
    FH_REQUIRE(slots >= 1, "workspace too small (see fh_tridiag_repair_flagged_workspace_bytes)");
    if (slots > kPivSlotsMax) slots = kPivSlotsMax;
    if (slots > batch) slots = batch;
    const RepairPlan p = repair_plan(n, batch, slots);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace);
    int32_t *header = reinterpret_cast<int32_t *>(ws + p.off_header);
    FH_CUDA(cudaMemsetAsync(header, 0, 8 * sizeof(int32_t), st));
    const int64_t per_cta = (batch + slots - 1) / slots;
    repair_flagged_kernel<<<(unsigned)slots, kPivThreads, 0, st>>>(
        reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d), reinterpret_cast<const cplx *>(du),
        reinterpret_cast<const cplx *>(b), reinterpret_cast<cplx *>(u), n, batch, per_cta, header, accept_tol,
        reinterpret_cast<cplx *>(ws + p.off_work), status);
    FH_LAUNCH_CHECK("repair_flagged_kernel");
    return FH_OK;
}
