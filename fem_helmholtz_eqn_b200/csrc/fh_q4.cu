// 2D bilinear-quad (Q4) Helmholtz assembly on the annulus meshes of the reference's src/solvers/*:
// element matrices, Robin / Neumann boundary-edge terms, deterministic CSR assembly, Dirichlet rows.
//
//   fh_q4_element_matrices_c128   <- BaseSolver.get_element_matrices            base.py:65-117
//   fh_q4_robin_edge_matrices_c128<- *.sommerfeld_element_matrices              fem2d.py:29-80, fem2d_dirichlet_sommerfeld.py:26-78
//   fh_q4_neumann_edge_vectors_c128<- *.neumann_element_matrices (given g)      fem2d.py:82-137
//   fh_q4_scatter_keys + fh_sorted_segments + fh_segment_sum_c128 <- the `+=` scatter loops of *.assemble()
//                                                                               fem2d.py:139-171 etc.
//   fh_q4_boundary_flags + fh_csr_apply_dirichlet_c128 <- apply_boundary_conditions  fem2d_dirichlet.py:11-29 etc.
//
// Bit parity: every element-level expression is evaluated with explicitly rounded intrinsics in the
// reference's order (Python evaluates left to right, numpy float64 scalars are IEEE doubles).  The one
// third-party quirk is np.linalg.norm of a 2-vector, which goes through BLAS ddot and rounds as
// fma(y, y, x*x) (checked against numpy 2.3.5 / OpenBLAS 0.3.30); edge lengths reproduce that.
//
// Determinism: the scatter is a gather.  Every (element, local slot) contribution gets the key row*n+col;
// a STABLE radix sort (CUB) groups equal keys while keeping element order, so one thread per non-zero adds
// its contributions in exactly the order of the reference's `K[i, j] += K_e[i, j]` loop -- no atomics.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "fh_q4.cuh"

namespace fh {

// One thread per element: 4 node coordinates gathered through the connectivity, 2x2 Gauss quadrature in
// registers, 16 complex entries out (row-major 4x4).  base.py:65-117.
__global__ void __launch_bounds__(kQ4Threads)
q4_element_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ elements, int64_t ne, double k2,
                  Gauss gq, cplx *__restrict__ Ke) {
    const int64_t e = (int64_t)blockIdx.x * kQ4Threads + threadIdx.x;
    if (e >= ne) return;
    double x[4], y[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t nd = elements[4 * e + a];
        const double2 c = *reinterpret_cast<const double2 *>(nodes + 2 * nd);
        x[a] = c.x, y[a] = c.y;
    }
    double K[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) K[q] = 0.0;
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j) {
            const Q4Shape s = q4_shape(gq.pt[i], gq.pt[j]);
            double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                J00 = __dadd_rn(J00, __dmul_rn(s.dxi[a], x[a]));
                J01 = __dadd_rn(J01, __dmul_rn(s.dxi[a], y[a]));
                J10 = __dadd_rn(J10, __dmul_rn(s.deta[a], x[a]));
                J11 = __dadd_rn(J11, __dmul_rn(s.deta[a], y[a]));
            }
            const double detJ = __dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J01, J10));
            const double i00 = __ddiv_rn(J11, detJ), i01 = __ddiv_rn(-J01, detJ);
            const double i10 = __ddiv_rn(-J10, detJ), i11 = __ddiv_rn(J00, detJ);
            double dx[4], dy[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                dx[a] = __dadd_rn(__dmul_rn(i00, s.dxi[a]), __dmul_rn(i01, s.deta[a]));
                dy[a] = __dadd_rn(__dmul_rn(i10, s.dxi[a]), __dmul_rn(i11, s.deta[a]));
            }
            const double wd = __dmul_rn(__dmul_rn(gq.wt[i], gq.wt[j]), detJ);
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const double grad = __dadd_rn(__dmul_rn(dx[m], dx[n]), __dmul_rn(dy[m], dy[n]));
                    const double mass = __dmul_rn(__dmul_rn(k2, s.N[m]), s.N[n]);
                    K[4 * m + n] = __dadd_rn(K[4 * m + n], __dmul_rn(wd, __dadd_rn(-grad, mass)));
                }
        }
    }
    cplx *out = Ke + 16 * e;
#pragma unroll
    for (int q = 0; q < 16; ++q) out[q] = mk(K[q], 0.0);
}

// Which of the 4 edges of each listed element lie on the circle of the given radius, and their lengths.
__global__ void q4_edge_geometry_kernel(const double *__restrict__ nodes, const int32_t *__restrict__ elements,
                                        const int32_t *__restrict__ ids, int64_t nb, double radius,
                                        uint8_t *__restrict__ on_edge, double *__restrict__ length) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    const int64_t e = ids[q];
    double x[4], y[4];
    bool on[4];
    for (int a = 0; a < 4; ++a) {
        const int64_t nd = elements[4 * e + a];
        x[a] = nodes[2 * nd], y[a] = nodes[2 * nd + 1];
        on[a] = isclose_default(norm2_blas(x[a], y[a]), radius);
    }
    for (int ed = 0; ed < 4; ++ed) {
        const int p0 = ed, p1 = (ed + 1) & 3;  // edges (0,1) (1,2) (2,3) (3,0)   fem2d.py:37-42
        on_edge[4 * q + ed] = (on[p0] && on[p1]) ? 1 : 0;
        length[4 * q + ed] = norm2_blas(__dsub_rn(x[p1], x[p0]), __dsub_rn(y[p1], y[p0]));
    }
}

// Robin (Sommerfeld / ABC) edge mass matrix per listed element: S_e[gm,gn] += N_m N_n * coef * (len/2) * w.
__global__ void q4_robin_kernel(const uint8_t *__restrict__ on_edge, const double *__restrict__ length, int64_t nb,
                                cplx coef, Gauss gq, cplx *__restrict__ Se) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    cplx S[16];
    for (int i = 0; i < 16; ++i) S[i] = czero();
    for (int ed = 0; ed < 4; ++ed) {
        if (!on_edge[4 * q + ed]) continue;
        const int nd[2] = {ed, (ed + 1) & 3};
        const double ds = __ddiv_rn(length[4 * q + ed], 2.0);
        for (int g = 0; g < gq.n; ++g) {
            const double N[2] = {__dmul_rn(0.5, __dsub_rn(1.0, gq.pt[g])), __dmul_rn(0.5, __dadd_rn(1.0, gq.pt[g]))};
            for (int lm = 0; lm < 2; ++lm)
                for (int ln = 0; ln < 2; ++ln) {
                    const double r = __dmul_rn(N[lm], N[ln]);
                    const double re = __dmul_rn(__dmul_rn(__dmul_rn(r, coef.re), ds), gq.wt[g]);
                    const double im = __dmul_rn(__dmul_rn(__dmul_rn(r, coef.im), ds), gq.wt[g]);
                    cplx &t = S[4 * nd[lm] + nd[ln]];
                    t = mk(__dadd_rn(t.re, re), __dadd_rn(t.im, im));
                }
        }
    }
    for (int i = 0; i < 16; ++i) Se[16 * q + i] = S[i];
}

// Neumann load per listed element: N_e[gm] += N_m * g * (len/2) * w with g given at the edge Gauss points.
__global__ void q4_neumann_kernel(const uint8_t *__restrict__ on_edge, const double *__restrict__ length, int64_t nb,
                                  const cplx *__restrict__ g_vals, Gauss gq, cplx *__restrict__ Ne) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    cplx V[4] = {czero(), czero(), czero(), czero()};
    for (int ed = 0; ed < 4; ++ed) {
        if (!on_edge[4 * q + ed]) continue;
        const int nd[2] = {ed, (ed + 1) & 3};
        const double ds = __ddiv_rn(length[4 * q + ed], 2.0);
        for (int g = 0; g < gq.n; ++g) {
            const double N[2] = {__dmul_rn(0.5, __dsub_rn(1.0, gq.pt[g])), __dmul_rn(0.5, __dadd_rn(1.0, gq.pt[g]))};
            const cplx gv = g_vals[(4 * q + ed) * gq.n + g];
            for (int lm = 0; lm < 2; ++lm) {
                const double re = __dmul_rn(__dmul_rn(__dmul_rn(N[lm], gv.re), ds), gq.wt[g]);
                const double im = __dmul_rn(__dmul_rn(__dmul_rn(N[lm], gv.im), ds), gq.wt[g]);
                V[nd[lm]] = mk(__dadd_rn(V[nd[lm]].re, re), __dadd_rn(V[nd[lm]].im, im));
            }
        }
    }
    for (int i = 0; i < 4; ++i) Ne[4 * q + i] = V[i];
}

// Scatter keys: matrix entries (16 per element, key = row * n_nodes + col) or vector entries (4 per element).
__global__ void q4_keys_kernel(const int32_t *__restrict__ elements, const int32_t *__restrict__ ids, int64_t count,
                               int64_t n_nodes, int matrix, int64_t *__restrict__ keys) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    const int64_t e = ids ? ids[q] : q;
    int64_t nd[4];
    for (int a = 0; a < 4; ++a) nd[a] = elements[4 * e + a];
    if (matrix) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) keys[16 * q + 4 * i + j] = nd[i] * n_nodes + nd[j];
    } else {
        for (int i = 0; i < 4; ++i) keys[4 * q + i] = nd[i];
    }
}

__global__ void iota_kernel(int32_t *p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

// rowptr[r] = first unique key >= r * n_cols  (binary search; unique keys are sorted)
__global__ void csr_rowptr_kernel(const int64_t *__restrict__ ukeys, const int64_t *__restrict__ nnz_dev,
                                  int64_t n_rows, int64_t n_cols, int64_t *__restrict__ rowptr,
                                  int32_t *__restrict__ colidx) {
    const int64_t nnz = *nnz_dev;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_rows) {
        const int64_t target = i * n_cols;
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ukeys[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        rowptr[i] = lo;
    }
    if (colidx != nullptr)
        for (int64_t q = i; q < nnz; q += (int64_t)gridDim.x * blockDim.x) colidx[q] = (int32_t)(ukeys[q] % n_cols);
}

// out[s] = (sum of contributions with id < split, in order) + (sum of the others, in order)
__global__ void segment_sum_kernel(const cplx *__restrict__ contrib, const int32_t *__restrict__ perm,
                                   const int64_t *__restrict__ seg, const int64_t *__restrict__ nseg_dev,
                                   int64_t split, cplx *__restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= *nseg_dev) return;
    double are = 0.0, aim = 0.0, bre = 0.0, bim = 0.0;
    bool has_b = false;
    for (int64_t q = seg[s]; q < seg[s + 1]; ++q) {
        const int32_t id = perm[q];
        const cplx v = contrib[id];
        if (id < split) are = __dadd_rn(are, v.re), aim = __dadd_rn(aim, v.im);
        else bre = __dadd_rn(bre, v.re), bim = __dadd_rn(bim, v.im), has_b = true;
    }
    // K += S touches every entry of the dense reference matrix; adding an all-zero S changes nothing
    out[s] = has_b ? mk(__dadd_rn(are, bre), __dadd_rn(aim, bim)) : mk(are, aim);
}

__global__ void boundary_flags_kernel(const double *__restrict__ nodes, int64_t n, double radius, double tol,
                                      uint8_t *__restrict__ flags, uint8_t bit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = nodes[2 * i], y = nodes[2 * i + 1];
    const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));  // np.linalg.norm(nodes, axis=1)
    if (fabs(__dsub_rn(r, radius)) < tol) flags[i] |= bit;
}

// Row replacement: A[row] = 0; A[row,row] = 1; b[row] = value[flag]    (flag bit 1 = inner, bit 2 = outer)
__global__ void csr_dirichlet_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                     cplx *__restrict__ vals, cplx *__restrict__ F, const uint8_t *__restrict__ flags,
                                     int64_t n, uint8_t mask, cplx g_inner, cplx g_outer) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t f = flags[r] & mask;
    if (!f) return;
    for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) vals[q] = colidx[q] == r ? cone() : czero();
    // inner rows are written first, outer rows afterwards (fem2d_dirichlet.py:20-27)
    F[r] = (f & 2) ? g_outer : g_inner;
}

__global__ void csr_to_dense_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                    const cplx *__restrict__ vals, int64_t n, cplx *__restrict__ A) {
    const int64_t r = blockIdx.x;
    for (int64_t q = rowptr[r] + threadIdx.x; q < rowptr[r + 1]; q += blockDim.x) A[r * n + colidx[q]] = vals[q];
}

static Gauss gauss_from_host(const double *pts, const double *wts, int n) {
    Gauss g;
    g.n = n;
    for (int i = 0; i < 4; ++i) g.pt[i] = i < n ? pts[i] : 0.0, g.wt[i] = i < n ? wts[i] : 0.0;
    return g;
}
static unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace fh

extern "C" {

int fh_q4_element_matrices_c128(const double *nodes, const int32_t *elements, int64_t n_elements, double k_squared,
                                const double *gauss_pts_host, const double *gauss_wts_host, fh_c128 *Ke,
                                fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_elements >= 0, "n_elements must be >= 0");
    if (n_elements == 0) return FH_OK;
    FH_REQUIRE(nodes && elements && gauss_pts_host && gauss_wts_host && Ke, "NULL pointer");
    q4_element_kernel<<<blocks_for(n_elements, kQ4Threads), kQ4Threads, 0, static_cast<cudaStream_t>(stream)>>>(
        nodes, elements, n_elements, k_squared, gauss_from_host(gauss_pts_host, gauss_wts_host, 2),
        reinterpret_cast<cplx *>(Ke));
    FH_LAUNCH_CHECK("q4_element_kernel");
    return FH_OK;
}

int fh_q4_edge_geometry(const double *nodes, const int32_t *elements, const int32_t *elem_ids, int64_t n_ids,
                        double radius, uint8_t *on_edge, double *length, fh_stream_t stream) {
    using namespace fh;
    if (n_ids <= 0) return FH_OK;
    FH_REQUIRE(nodes && elements && elem_ids && on_edge && length, "NULL pointer");
    q4_edge_geometry_kernel<<<blocks_for(n_ids, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        nodes, elements, elem_ids, n_ids, radius, on_edge, length);
    FH_LAUNCH_CHECK("q4_edge_geometry_kernel");
    return FH_OK;
}

int fh_q4_robin_edge_matrices_c128(const uint8_t *on_edge, const double *length, int64_t n_ids, double coef_re,
                                   double coef_im, const double *gauss_pts_host, const double *gauss_wts_host,
                                   int n_gauss, fh_c128 *Se, fh_stream_t stream) {
    using namespace fh;
    if (n_ids <= 0) return FH_OK;
    FH_REQUIRE(on_edge && length && gauss_pts_host && gauss_wts_host && Se, "NULL pointer");
    FH_REQUIRE(n_gauss >= 1 && n_gauss <= 4, "n_gauss must be in [1, 4]");
    q4_robin_kernel<<<blocks_for(n_ids, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        on_edge, length, n_ids, mk(coef_re, coef_im), gauss_from_host(gauss_pts_host, gauss_wts_host, n_gauss),
        reinterpret_cast<cplx *>(Se));
    FH_LAUNCH_CHECK("q4_robin_kernel");
    return FH_OK;
}

int fh_q4_neumann_edge_vectors_c128(const uint8_t *on_edge, const double *length, int64_t n_ids, const fh_c128 *g_vals,
                                    const double *gauss_pts_host, const double *gauss_wts_host, int n_gauss,
                                    fh_c128 *Ne, fh_stream_t stream) {
    using namespace fh;
    if (n_ids <= 0) return FH_OK;
    FH_REQUIRE(on_edge && length && g_vals && gauss_pts_host && gauss_wts_host && Ne, "NULL pointer");
    FH_REQUIRE(n_gauss >= 1 && n_gauss <= 4, "n_gauss must be in [1, 4]");
    q4_neumann_kernel<<<blocks_for(n_ids, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        on_edge, length, n_ids, reinterpret_cast<const cplx *>(g_vals),
        gauss_from_host(gauss_pts_host, gauss_wts_host, n_gauss), reinterpret_cast<cplx *>(Ne));
    FH_LAUNCH_CHECK("q4_neumann_kernel");
    return FH_OK;
}

int fh_q4_scatter_keys(const int32_t *elements, const int32_t *elem_ids, int64_t count, int64_t n_nodes, int matrix,
                       int64_t *keys, fh_stream_t stream) {
    using namespace fh;
    if (count <= 0) return FH_OK;
    FH_REQUIRE(elements && keys, "NULL pointer");
    q4_keys_kernel<<<blocks_for(count, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(elements, elem_ids, count,
                                                                                           n_nodes, matrix, keys);
    FH_LAUNCH_CHECK("q4_keys_kernel");
    return FH_OK;
}

// Symbolic phase.  keys[n_entries] -> stable order `perm`, unique sorted keys `ukeys`, segment offsets
// `seg[n_unique + 1]`, `n_unique_dev`.  Workspace from fh_sorted_segments_workspace_bytes().
size_t fh_sorted_segments_workspace_bytes(int64_t n_entries) {
    size_t sort_b = 0, rle_b = 0, scan_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const int64_t *)nullptr, (int64_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, n_entries);
    cub::DeviceRunLengthEncode::Encode(nullptr, rle_b, (const int64_t *)nullptr, (int64_t *)nullptr, (int64_t *)nullptr,
                                       (int64_t *)nullptr, n_entries);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (const int64_t *)nullptr, (int64_t *)nullptr, n_entries + 1);
    size_t tmp = sort_b > rle_b ? sort_b : rle_b;
    if (scan_b > tmp) tmp = scan_b;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // layout: [sorted keys][iota][counts][cub temp]
    return up(sizeof(int64_t) * n_entries) + up(sizeof(int32_t) * n_entries) + up(sizeof(int64_t) * (n_entries + 1)) +
           up(tmp) + 256;
}

int fh_sorted_segments(const int64_t *keys, int64_t n_entries, int32_t *perm, int64_t *ukeys, int64_t *seg,
                       int64_t *n_unique_dev, void *workspace, size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_entries > 0 && n_entries < (1ll << 31), "n_entries must be in (0, 2^31)");
    FH_REQUIRE(keys && perm && ukeys && seg && n_unique_dev && workspace, "NULL pointer");
    FH_REQUIRE(workspace_bytes >= fh_sorted_segments_workspace_bytes(n_entries), "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    char *w = static_cast<char *>(workspace);
    int64_t *skeys = reinterpret_cast<int64_t *>(w);
    w += up(sizeof(int64_t) * n_entries);
    int32_t *iota = reinterpret_cast<int32_t *>(w);
    w += up(sizeof(int32_t) * n_entries);
    int64_t *counts = reinterpret_cast<int64_t *>(w);
    w += up(sizeof(int64_t) * (n_entries + 1));
    size_t tmp_bytes = workspace_bytes - (size_t)(w - static_cast<char *>(workspace));
    iota_kernel<<<blocks_for(n_entries, 256), 256, 0, st>>>(iota, n_entries);
    FH_LAUNCH_CHECK("iota_kernel");
    FH_CUDA(cub::DeviceRadixSort::SortPairs(w, tmp_bytes, keys, skeys, iota, perm, n_entries, 0, 64, st));
    FH_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (n_entries + 1), st));
    FH_CUDA(cub::DeviceRunLengthEncode::Encode(w, tmp_bytes, skeys, ukeys, counts, n_unique_dev, n_entries, st));
    // exclusive scan over n_entries + 1 counts (trailing zeros) gives seg[0 .. n_unique]
    FH_CUDA(cub::DeviceScan::ExclusiveSum(w, tmp_bytes, counts, seg, n_entries + 1, st));
    return FH_OK;
}

int fh_csr_from_keys(const int64_t *ukeys, const int64_t *n_unique_dev, int64_t n_rows, int64_t n_cols, int64_t *rowptr,
                     int32_t *colidx, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(ukeys && n_unique_dev && rowptr && colidx, "NULL pointer");
    int64_t work = n_rows + 1;
    csr_rowptr_kernel<<<blocks_for(work > 65536 ? work : 65536, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ukeys, n_unique_dev, n_rows, n_cols, rowptr, colidx);
    FH_LAUNCH_CHECK("csr_rowptr_kernel");
    return FH_OK;
}

int fh_segment_sum_c128(const fh_c128 *contrib, const int32_t *perm, const int64_t *seg, const int64_t *n_seg_dev,
                        int64_t n_seg_max, int64_t split, fh_c128 *out, fh_stream_t stream) {
    using namespace fh;
    if (n_seg_max <= 0) return FH_OK;
    FH_REQUIRE(contrib && perm && seg && n_seg_dev && out, "NULL pointer");
    segment_sum_kernel<<<blocks_for(n_seg_max, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const cplx *>(contrib), perm, seg, n_seg_dev, split, reinterpret_cast<cplx *>(out));
    FH_LAUNCH_CHECK("segment_sum_kernel");
    return FH_OK;
}

int fh_q4_boundary_flags(const double *nodes, int64_t n_nodes, double radius, double tol, int bit, uint8_t *flags,
                         fh_stream_t stream) {
    using namespace fh;
    if (n_nodes <= 0) return FH_OK;
    FH_REQUIRE(nodes && flags && (bit == 1 || bit == 2), "bad arguments");
    boundary_flags_kernel<<<blocks_for(n_nodes, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        nodes, n_nodes, radius, tol, flags, (uint8_t)bit);
    FH_LAUNCH_CHECK("boundary_flags_kernel");
    return FH_OK;
}

int fh_csr_apply_dirichlet_c128(const int64_t *rowptr, const int32_t *colidx, fh_c128 *vals, fh_c128 *F,
                                const uint8_t *flags, int64_t n, int mask, double g_inner_re, double g_inner_im,
                                double g_outer_re, double g_outer_im, fh_stream_t stream) {
    using namespace fh;
    if (n <= 0) return FH_OK;
    FH_REQUIRE(rowptr && colidx && vals && F && flags, "NULL pointer");
    csr_dirichlet_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, colidx, reinterpret_cast<cplx *>(vals), reinterpret_cast<cplx *>(F), flags, n, (uint8_t)mask,
        mk(g_inner_re, g_inner_im), mk(g_outer_re, g_outer_im));
    FH_LAUNCH_CHECK("csr_dirichlet_kernel");
    return FH_OK;
}

int fh_csr_to_dense_c128(const int64_t *rowptr, const int32_t *colidx, const fh_c128 *vals, int64_t n, fh_c128 *A,
                         fh_stream_t stream) {
    using namespace fh;
    if (n <= 0) return FH_OK;
    FH_REQUIRE(rowptr && colidx && vals && A, "NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FH_CUDA(cudaMemsetAsync(A, 0, sizeof(cplx) * (size_t)n * (size_t)n, st));
    csr_to_dense_kernel<<<(unsigned)n, 32, 0, st>>>(rowptr, colidx, reinterpret_cast<const cplx *>(vals), n,
                                                    reinterpret_cast<cplx *>(A));
    FH_LAUNCH_CHECK("csr_to_dense_kernel");
    return FH_OK;
}

}  // extern "C"
