// 1D P1 Helmholtz assembly into tridiagonal bands (replaces nb:134-187) and the tiny element
// routines behind the reference's shape_functions / element_matrices / boundary_matrices.
//
// Layout: dl, d, du, b are complex128[n_k][N+1], batch-major.  One thread owns one mesh row j
// (= node j): it evaluates the two adjacent element matrices in registers, sums them in element
// order (deterministic, no atomics), then streams that row for every wavenumber of its k-slab.
// Consecutive threads write consecutive 16-byte entries -> fully coalesced 512 B per warp store.
//
// Roofline: HBM-bound, write-only.  Algorithmic bytes: 64 B per DOF (4 complex128), + 8 B per
// DOF of node coordinates per k-slab in FH_H_NODAL mode.
#include "fh_helm1d.cuh"

namespace fh {

constexpr int kAsmThreads = 256;

__global__ void __launch_bounds__(kAsmThreads)
assemble1d_kernel(Problem1d p, const double *__restrict__ ks, const double *__restrict__ k2s,
                  int64_t n_k, int64_t k_per_block, cplx *__restrict__ dl, cplx *__restrict__ d,
                  cplx *__restrict__ du, cplx *__restrict__ b) {
    // The wavenumbers of the block's slab are staged in shared memory 256 at a time, so the store loop
    // has no global-load latency on its critical path (v1 waited on two __ldg per row: 62% of HBM peak).
    __shared__ double sk[2][kAsmThreads];
    const int64_t n = p.n_elem + 1;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * kAsmThreads + threadIdx.x;
    const bool live = j < n;
    RowGeom g = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (live) g = row_geom(p, j);
    const int64_t k_begin = static_cast<int64_t>(blockIdx.y) * k_per_block;
    const int64_t k_end = min(n_k, k_begin + k_per_block);
    for (int64_t base = k_begin; base < k_end; base += kAsmThreads) {
        const int cnt = (int)min((int64_t)kAsmThreads, k_end - base);
        __syncthreads();
        if (threadIdx.x < cnt) {
            sk[0][threadIdx.x] = __ldg(ks + base + threadIdx.x);
            sk[1][threadIdx.x] = __ldg(k2s + base + threadIdx.x);
        }
        __syncthreads();
        if (!live) continue;
        int64_t o = base * n + j;
#pragma unroll 4
        for (int i = 0; i < cnt; ++i, o += n) {
            cplx vl, vd, vu, vb;
            row_values(p, g, j, sk[0][i], sk[1][i], vl, vd, vu, vb);
            stg_stream(dl + o, vl);
            stg_stream(d + o, vd);
            stg_stream(du + o, vu);
            stg_stream(b + o, vb);
        }
    }
}

// Scalar-h fast path (the reference's only mode, nb:158: one element matrix for the whole mesh).  All
// interior rows of a system hold the same three numbers, so the kernel is a structured fill: every CTA owns
// a contiguous slab of wavenumbers and streams the four arrays front to back -- four long sequential
// write streams per CTA instead of 4 KB pieces 64 KB apart (which held the row-tiled kernel to 4.2 TB/s,
// against 7.4 TB/s for a plain fill on the same GPU; tools/write_bw_probe.py).
__global__ void __launch_bounds__(kAsmThreads)
assemble1d_scalar_kernel(Problem1d p, const double *__restrict__ ks, const double *__restrict__ k2s,
                         int64_t n_k, int64_t per_block, cplx *__restrict__ dl, cplx *__restrict__ d,
                         cplx *__restrict__ du, cplx *__restrict__ b) {
    // every CTA owns a contiguous range of the flattened (wavenumber, row) index space, whatever the aspect
    // ratio of the batch is (65,536 short systems or one system of 2^27 rows)
    const int64_t n = p.n_elem + 1;
    const int64_t o_begin = static_cast<int64_t>(blockIdx.x) * per_block;
    const int64_t o_end = min(n_k * n, o_begin + per_block);
    if (o_begin >= o_end) return;
    const RowGeom g_first = row_geom(p, 0);
    const RowGeom g_mid = row_geom(p, p.n_elem > 1 ? 1 : 0);
    const RowGeom g_last = row_geom(p, p.n_elem);
    int64_t o = o_begin + threadIdx.x;
    int64_t s = o / n;      // wavenumber index; (s, j) follow o incrementally from here on
    int64_t j = o - s * n;  // row
    int64_t s_cached = -1;
    cplx il = czero(), id = czero(), iu = czero(), ib = czero();  // interior row of the cached wavenumber
    double k = 0.0, k2 = 0.0;
    for (; o < o_end; o += kAsmThreads) {
        if (s != s_cached) {
            k = __ldg(ks + s), k2 = __ldg(k2s + s);
            row_values(p, g_mid, 1, k, k2, il, id, iu, ib);
            s_cached = s;
        }
        cplx vl = il, vd = id, vu = iu, vb = ib;
        if (j == 0 || j == p.n_elem) row_values(p, j == 0 ? g_first : g_last, j, k, k2, vl, vd, vu, vb);
        stg_stream(dl + o, vl);
        stg_stream(d + o, vd);
        stg_stream(du + o, vu);
        stg_stream(b + o, vb);
        j += kAsmThreads;
        while (j >= n) j -= n, ++s;
    }
}

// Non-uniform meshes (FH_H_NODAL), many wavenumbers: the same structured fill, with the k-independent geometry of every
// row (six numbers, IEEE divisions) evaluated ONCE into a small table (48 B per row: L2-resident) instead of once per
// CTA row-tile, so that every CTA can stream contiguous slabs of the four arrays like the scalar-h kernel does.
__global__ void __launch_bounds__(kAsmThreads)
row_geometry_kernel(Problem1d p, double *__restrict__ geom /* [6][n] */) {
    const int64_t n = p.n_elem + 1;
    const int64_t j = static_cast<int64_t>(blockIdx.x) * kAsmThreads + threadIdx.x;
    if (j >= n) return;
    const RowGeom g = row_geom(p, j);
    geom[j] = g.Kd, geom[n + j] = g.Md, geom[2 * n + j] = g.Kl, geom[3 * n + j] = g.Ml;
    geom[4 * n + j] = g.Ku, geom[5 * n + j] = g.Mu;
}

__global__ void __launch_bounds__(kAsmThreads)
assemble1d_table_kernel(Problem1d p, const double *__restrict__ geom, const double *__restrict__ ks,
                        const double *__restrict__ k2s, int64_t n_k, int64_t per_block, cplx *__restrict__ dl,
                        cplx *__restrict__ d, cplx *__restrict__ du, cplx *__restrict__ b) {
    const int64_t n = p.n_elem + 1;
    const int64_t o_begin = static_cast<int64_t>(blockIdx.x) * per_block;
    const int64_t o_end = min(n_k * n, o_begin + per_block);
    if (o_begin >= o_end) return;
    int64_t o = o_begin + threadIdx.x;
    int64_t s = o / n;      // wavenumber index; (s, j) follow o incrementally from here on
    int64_t j = o - s * n;  // row
    int64_t s_cached = -1;
    double k = 0.0, k2 = 0.0;
    for (; o < o_end; o += kAsmThreads) {
        if (s != s_cached) k = __ldg(ks + s), k2 = __ldg(k2s + s), s_cached = s;
        RowGeom g;
        g.Kd = __ldg(geom + j), g.Md = __ldg(geom + n + j), g.Kl = __ldg(geom + 2 * n + j);
        g.Ml = __ldg(geom + 3 * n + j), g.Ku = __ldg(geom + 4 * n + j), g.Mu = __ldg(geom + 5 * n + j);
        cplx vl, vd, vu, vb;
        row_values(p, g, j, k, k2, vl, vd, vu, vb);
        stg_stream(dl + o, vl);
        stg_stream(d + o, vd);
        stg_stream(du + o, vu);
        stg_stream(b + o, vb);
        j += kAsmThreads;
        while (j >= n) j -= n, ++s;
    }
}

// One wavenumber, a contiguous range of mesh rows [row_begin, row_begin + n_rows): the slab of one rank of the partitioned
// solve (BASELINE config 4 with A materialised), or a window of a long mesh.  Same structured fill; (k, k**2) may live in
// device memory (kk != NULL) so that a captured graph can be replayed for another wavenumber.
__global__ void __launch_bounds__(kAsmThreads)
assemble1d_slab_kernel(Problem1d p, double k, double k2, const double *__restrict__ kk, int64_t row_begin,
                       int64_t n_rows, int64_t per_block, cplx *__restrict__ dl, cplx *__restrict__ d,
                       cplx *__restrict__ du, cplx *__restrict__ b) {
    const int64_t o_begin = static_cast<int64_t>(blockIdx.x) * per_block;
    const int64_t o_end = min(n_rows, o_begin + per_block);
    if (o_begin >= o_end) return;
    if (kk != nullptr) k = __ldg(kk), k2 = __ldg(kk + 1);
    const bool scalar = p.h_mode == FH_H_SCALAR && p.n_elem >= 2;
    cplx il = czero(), id = czero(), iu = czero(), ib = czero();
    if (scalar) row_values(p, row_geom(p, 1), 1, k, k2, il, id, iu, ib);
    for (int64_t o = o_begin + threadIdx.x; o < o_end; o += kAsmThreads) {
        const int64_t j = row_begin + o;
        cplx vl = il, vd = id, vu = iu, vb = ib;
        if (!scalar || j == 0 || j == p.n_elem) row_values(p, row_geom(p, j), j, k, k2, vl, vd, vu, vb);
        stg_stream(dl + o, vl);
        stg_stream(d + o, vd);
        stg_stream(du + o, vu);
        stg_stream(b + o, vb);
    }
}

__global__ void shape1d_kernel(double xi, double *phi, double *dphi) {
    double a[2], c[2];
    shape1d(xi, a, c);
    phi[0] = a[0], phi[1] = a[1], dphi[0] = c[0], dphi[1] = c[1];
}

__global__ void elem1d_kernel(double h, double xi, double w, double *Ke, double *Me) {
    const Elem1d e = elem1d(h, xi, w);
    Ke[0] = e.k00, Ke[1] = e.k01, Ke[2] = e.k10, Ke[3] = e.k11;
    Me[0] = e.m00, Me[1] = e.m01, Me[2] = e.m10, Me[3] = e.m11;
}

// nb:99-131.  Ne_i = phi_i(-1) * (-i k cos(theta)),  Se_ij = i k phi_i(1) phi_j(1)
__global__ void boundary1d_kernel(double k, double cos_theta, cplx *Ne, cplx *Se) {
    double pl[2], pr[2], dd[2];
    shape1d(-1.0, pl, dd);
    shape1d(1.0, pr, dd);
    const double flux_im = -__dmul_rn(k, cos_theta);
    for (int i = 0; i < 2; ++i) {
        Ne[i] = mk(0.0, __dmul_rn(pl[i], flux_im));
        for (int jj = 0; jj < 2; ++jj) Se[2 * i + jj] = mk(0.0, __dmul_rn(__dmul_rn(k, pr[i]), pr[jj]));
    }
}

}  // namespace fh

extern "C" {

int fh_shape_functions_1d(double xi, double *phi2, double *dphi2, fh_stream_t stream) {
    FH_REQUIRE(phi2 && dphi2, "output pointers must not be NULL");
    fh::shape1d_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(xi, phi2, dphi2);
    FH_LAUNCH_CHECK("shape1d_kernel");
    return FH_OK;
}

int fh_element_matrices_1d(double h, double xi, double w, double *Ke4, double *Me4, fh_stream_t stream) {
    FH_REQUIRE(Ke4 && Me4, "output pointers must not be NULL");
    fh::elem1d_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(h, xi, w, Ke4, Me4);
    FH_LAUNCH_CHECK("elem1d_kernel");
    return FH_OK;
}

int fh_boundary_matrices_1d(double k, double cos_theta, fh_c128 *Ne2, fh_c128 *Se4, fh_stream_t stream) {
    FH_REQUIRE(Ne2 && Se4, "output pointers must not be NULL");
    fh::boundary1d_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(
        k, cos_theta, reinterpret_cast<fh::cplx *>(Ne2), reinterpret_cast<fh::cplx *>(Se4));
    FH_LAUNCH_CHECK("boundary1d_kernel");
    return FH_OK;
}

static int assemble_impl(const fh_problem1d *prob_host, const double *ks, const double *k2s, int64_t n_k, fh_c128 *dl,
                         fh_c128 *d, fh_c128 *du, fh_c128 *b, void *workspace, size_t workspace_bytes,
                         fh_stream_t stream) {
    fh::Problem1d p;
    int rc = fh::problem_from_host(prob_host, p);
    if (rc != FH_OK) return rc;
    FH_REQUIRE(n_k >= 0, "n_k must be >= 0");
    if (n_k == 0) return FH_OK;
    FH_REQUIRE(ks && k2s && dl && d && du && b, "NULL device pointer");
    const int64_t n = p.n_elem + 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p.h_mode == FH_H_SCALAR && p.n_elem >= 2) {  // structured-fill fast path
        const int64_t total = n_k * n;
        int64_t slabs = 8ll * fh::sm_count();
        int64_t per = (total + slabs - 1) / slabs;
        per = ((per + fh::kAsmThreads - 1) / fh::kAsmThreads) * fh::kAsmThreads;  // whole passes of the CTA
        slabs = (total + per - 1) / per;
        fh::assemble1d_scalar_kernel<<<(unsigned)slabs, fh::kAsmThreads, 0, st>>>(
            p, ks, k2s, n_k, per, reinterpret_cast<fh::cplx *>(dl), reinterpret_cast<fh::cplx *>(d),
            reinterpret_cast<fh::cplx *>(du), reinterpret_cast<fh::cplx *>(b));
        FH_LAUNCH_CHECK("assemble1d_scalar_kernel");
        return FH_OK;
    }
    const int64_t row_blocks = (n + fh::kAsmThreads - 1) / fh::kAsmThreads;
    FH_REQUIRE(row_blocks < (1ll << 31), "mesh too large for one launch");
    if (workspace != nullptr && workspace_bytes >= sizeof(double) * 6 * (size_t)n) {
        // caller-supplied scratch: geometry table (48 B per row, evaluated once) + structured fill
        double *geom = static_cast<double *>(workspace);
        fh::row_geometry_kernel<<<(unsigned)row_blocks, fh::kAsmThreads, 0, st>>>(p, geom);
        FH_LAUNCH_CHECK("row_geometry_kernel");
        const int64_t total = n_k * n;
        int64_t slabs = 8ll * fh::sm_count();
        int64_t per = (total + slabs - 1) / slabs;
        per = ((per + fh::kAsmThreads - 1) / fh::kAsmThreads) * fh::kAsmThreads;
        slabs = (total + per - 1) / per;
        fh::assemble1d_table_kernel<<<(unsigned)slabs, fh::kAsmThreads, 0, st>>>(
            p, geom, ks, k2s, n_k, per, reinterpret_cast<fh::cplx *>(dl), reinterpret_cast<fh::cplx *>(d),
            reinterpret_cast<fh::cplx *>(du), reinterpret_cast<fh::cplx *>(b));
        FH_LAUNCH_CHECK("assemble1d_table_kernel");
        return FH_OK;
    }
    // No scratch: the row-tiled kernel (every thread owns a mesh row and evaluates its geometry itself).  Split the k
    // range so that the grid covers every SM several times even for short meshes, while each thread still amortises
    // its geometry over a slab of wavenumbers.
    const int64_t want_blocks = 8ll * fh::sm_count();
    int64_t k_splits = (want_blocks + row_blocks - 1) / row_blocks;
    if (k_splits < 1) k_splits = 1;
    if (k_splits > n_k) k_splits = n_k;
    if (k_splits > 65535) k_splits = 65535;
    const int64_t k_per_block = (n_k + k_splits - 1) / k_splits;
    k_splits = (n_k + k_per_block - 1) / k_per_block;
    dim3 grid(static_cast<unsigned>(row_blocks), static_cast<unsigned>(k_splits));
    fh::assemble1d_kernel<<<grid, fh::kAsmThreads, 0, st>>>(
        p, ks, k2s, n_k, k_per_block, reinterpret_cast<fh::cplx *>(dl), reinterpret_cast<fh::cplx *>(d),
        reinterpret_cast<fh::cplx *>(du), reinterpret_cast<fh::cplx *>(b));
    FH_LAUNCH_CHECK("assemble1d_kernel");
    return FH_OK;
}

int fh_assemble_1d_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s, int64_t n_k,
                        fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b, fh_stream_t stream) {
    return assemble_impl(prob_host, ks, k2s, n_k, dl, d, du, b, nullptr, 0, stream);
}

size_t fh_assemble_1d_workspace_bytes(int64_t n_elem, int h_mode, int64_t n_k) {
    // worth it from a handful of wavenumbers on: the table costs one pass over the mesh
    return (h_mode == FH_H_NODAL && n_k >= 8 && n_elem >= 1) ? sizeof(double) * 6 * (size_t)(n_elem + 1) : 0;
}

int fh_assemble_1d_ws_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s, int64_t n_k,
                           fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b, void *workspace, size_t workspace_bytes,
                           fh_stream_t stream) {
    return assemble_impl(prob_host, ks, k2s, n_k, dl, d, du, b, workspace, workspace_bytes, stream);
}

int fh_assemble_1d_slab_c128(const fh_problem1d *prob_host, double k, double k2, const double *k_k2_dev,
                             int64_t row_begin, int64_t n_rows, fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b,
                             fh_stream_t stream) {
    fh::Problem1d p;
    int rc = fh::problem_from_host(prob_host, p);
    if (rc != FH_OK) return rc;
    FH_REQUIRE(row_begin >= 0 && n_rows >= 0 && row_begin + n_rows <= p.n_elem + 1,
               "row range [%lld, %lld) outside the mesh of %lld nodes", (long long)row_begin,
               (long long)(row_begin + n_rows), (long long)(p.n_elem + 1));
    if (n_rows == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b, "NULL device pointer");
    int64_t slabs = 8ll * fh::sm_count();
    int64_t per = (n_rows + slabs - 1) / slabs;
    per = ((per + fh::kAsmThreads - 1) / fh::kAsmThreads) * fh::kAsmThreads;
    slabs = (n_rows + per - 1) / per;
    fh::assemble1d_slab_kernel<<<(unsigned)slabs, fh::kAsmThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        p, k, k2, k_k2_dev, row_begin, n_rows, per, reinterpret_cast<fh::cplx *>(dl), reinterpret_cast<fh::cplx *>(d),
        reinterpret_cast<fh::cplx *>(du), reinterpret_cast<fh::cplx *>(b));
    FH_LAUNCH_CHECK("assemble1d_slab_kernel");
    return FH_OK;
}

}  // extern "C"
