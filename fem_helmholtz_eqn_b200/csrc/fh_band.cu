// Sparse direct solve of the 2D systems: np.linalg.solve(K, F) of the four solver
// classes (fem2d.py:176, fem2d_dirichlet.py:51, fem2d_dirichlet_sommerfeld.py:109, fem2d_neumann_dirichlet.py:127)
// without ever forming the dense matrix.
//
// The CSR matrix is scattered, under an optional bandwidth-reducing node permutation, into LAPACK band storage
// (kl sub-, ku + kl super-diagonals: the kl extra ones take the fill of the row interchanges), factorised by
// Gaussian elimination with PARTIAL PIVOTING -- the same pivot rule (largest |re| + |im| in the column) and the same
// elimination order as LAPACK's zgbtf2 / the reference's zgesv, so accuracy matches np.linalg.solve -- and solved in
// place.  O(n (kl + ku) kl) flops and O(n (2 kl + ku)) bytes instead of O(n^3) / O(n^2); the Dirichlet row
// replacement makes K non-symmetric, which an LU with row pivoting does not care about.
//
// One persistent CTA of 1024 threads walks the columns: the band of the reference's meshes (n <~ 10^4, bandwidth
// <~ 10^2) is a few megabytes and lives in L2, each column step is four block barriers (pivot search, row swap,
// multipliers, rank-1 update of the kl x (ku + kl) window), the back-substitution one.  No atomics in the
// arithmetic: results are deterministic.
#include "fh_common.cuh"

namespace fh {

constexpr int kBandThreads = 1024;

// kl = max(p(i) - p(j)), ku = max(p(j) - p(i)) over the stored entries (p = perm, or the identity)
__global__ void csr_bandwidth_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                     const int32_t *__restrict__ perm, int64_t n, int32_t *__restrict__ klku) {
    int kl = 0, ku = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int pi = perm ? perm[i] : (int)i;
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) {
            const int pj = perm ? perm[colidx[q]] : colidx[q];
            kl = max(kl, pi - pj);
            ku = max(ku, pj - pi);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        kl = max(kl, __shfl_xor_sync(0xffffffffu, kl, o));
        ku = max(ku, __shfl_xor_sync(0xffffffffu, ku, o));
    }
    if ((threadIdx.x & 31) == 0) {  // integer max: order-independent
        atomicMax(&klku[0], kl);
        atomicMax(&klku[1], ku);
    }
}

// AB(kv + i - j, j) = A(i, j), column-major with leading dimension ld = 2 kl + ku + 1, kv = kl + ku; rhs permuted
__global__ void csr_to_band_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                   const cplx *__restrict__ vals, const cplx *__restrict__ b,
                                   const int32_t *__restrict__ perm, int64_t n, int kl, int ku, cplx *__restrict__ AB,
                                   cplx *__restrict__ rhs) {
    const int ld = 2 * kl + ku + 1, kv = kl + ku;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int pi = perm ? perm[i] : (int)i;
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) {
            const int pj = perm ? perm[colidx[q]] : colidx[q];
            AB[(int64_t)pj * ld + kv + pi - pj] = vals[q];
        }
        rhs[pi] = b[i];
    }
}

__device__ __forceinline__ double cabs1(cplx z) { return fabs(z.re) + fabs(z.im); }  // LAPACK izamax measure

__global__ void __launch_bounds__(kBandThreads)
band_lu_solve_kernel(cplx *__restrict__ AB, cplx *__restrict__ rhs, int64_t n, int kl, int ku,
                     const int32_t *__restrict__ perm, cplx *__restrict__ x, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    cplx *lcol = reinterpret_cast<cplx *>(band_smem);  // multipliers of the current column [kl]
    __shared__ double wbest[kBandThreads / 32];
    __shared__ int warg[kBandThreads / 32];
    __shared__ int s_jp;
    __shared__ int s_bad;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ld = 2 * kl + ku + 1, kv = kl + ku;
    if (t == 0) s_bad = 0;

    for (int64_t j = 0; j < n; ++j) {
        cplx *col = AB + j * ld;
        const int km = (int)min((int64_t)kl, n - 1 - j);         // rows below the diagonal in this column
        const int cw = (int)min((int64_t)kv, n - 1 - j);         // columns to the right that row interchanges can touch
        // ---- pivot search: first index of the largest |re| + |im| among rows j .. j + km --------------------
        double best = -1.0;
        int arg = 0;
        for (int i = t; i <= km; i += kBandThreads) {
            const double m = cabs1(col[kv + i]);
            if (m > best) best = m, arg = i;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) best = ob, arg = oa;
        }
        if (lane == 0) wbest[warp] = best, warg[warp] = arg;
        __syncthreads();
        if (warp == 0) {
            best = lane < kBandThreads / 32 ? wbest[lane] : -1.0;
            arg = lane < kBandThreads / 32 ? warg[lane] : 0;
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob > best || (ob == best && oa < arg)) best = ob, arg = oa;
            }
            if (lane == 0) {
                s_jp = arg;
                if (!(best > 0.0) || !isfinite(best)) s_bad = 1;  // exactly singular (or non-finite) column
            }
        }
        __syncthreads();
        const int jp = s_jp;
        // ---- interchange rows j and j + jp over columns j .. j + cw, and in the right-hand side -------------
        if (jp != 0) {
            for (int c = t; c <= cw; c += kBandThreads) {
                cplx *p = AB + (j + c) * ld + kv - c;  // A(j, j + c); A(j + jp, j + c) is jp entries further down
                const cplx tmp = p[0];
                p[0] = p[jp];
                p[jp] = tmp;
            }
            if (t == kBandThreads - 1) {
                const cplx tmp = rhs[j];
                rhs[j] = rhs[j + jp];
                rhs[j + jp] = tmp;
            }
        }
        __syncthreads();
        // ---- multipliers (kept in shared memory; L is applied to the right-hand side at once, not stored) ----
        const cplx inv = crecip(col[kv]);
        for (int i = 1 + t; i <= km; i += kBandThreads) lcol[i - 1] = col[kv + i] * inv;
        __syncthreads();
        // ---- rank-1 update of the window rows j+1 .. j+km, columns j+1 .. j+cw; rows fastest (contiguous) ----
        const int cells = km * cw;
        for (int q = t; q < cells; q += kBandThreads) {
            const int c = 1 + q / km, i = 1 + q % km;
            cplx *p = AB + (j + c) * ld + kv - c;  // A(j, j + c)
            p[i] = nmsub(p[i], lcol[i - 1], p[0]);
        }
        for (int i = 1 + t; i <= km; i += kBandThreads) rhs[j + i] = nmsub(rhs[j + i], lcol[i - 1], rhs[j]);
        __syncthreads();
    }
    // ---- back-substitution, column oriented: x_k = b_k / U(k,k), then b_i -= U(i,k) x_k for the kv rows above ----
    for (int64_t k = n - 1; k >= 0; --k) {
        const cplx *col = AB + k * ld;
        const cplx xk = rhs[k] * crecip(col[kv]);
        const int up = (int)min((int64_t)kv, k);
        __syncthreads();  // everybody has read rhs[k] before it is overwritten
        if (t == 0) {
            rhs[k] = xk;
            if (!cfinite(xk)) s_bad = 1;
        }
        for (int i = 1 + t; i <= up; i += kBandThreads) rhs[k - i] = nmsub(rhs[k - i], col[kv - i], xk);
        __syncthreads();
    }
    for (int64_t i = t; i < n; i += kBandThreads) x[i] = rhs[perm ? perm[i] : i];
    if (t == 0 && status != nullptr) *status = s_bad;
}

}  // namespace fh

extern "C" int fh_csr_bandwidth(const int64_t *rowptr, const int32_t *colidx, const int32_t *perm, int64_t n,
                                int32_t *klku, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0, "n must be >= 0");
    FH_REQUIRE(klku != nullptr, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FH_CUDA(cudaMemsetAsync(klku, 0, 2 * sizeof(int32_t), st));
    if (n == 0) return FH_OK;
    FH_REQUIRE(rowptr && colidx, "NULL device pointer");
    FH_REQUIRE(n < (1ll << 31), "n too large");
    const unsigned grid = (unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    csr_bandwidth_kernel<<<grid, 256, 0, st>>>(rowptr, colidx, perm, n, klku);
    FH_LAUNCH_CHECK("csr_bandwidth_kernel");
    return FH_OK;
}

extern "C" size_t fh_band_lu_workspace_bytes(int64_t n, int64_t kl, int64_t ku) {
    if (n <= 0 || kl < 0 || ku < 0) return 0;
    return sizeof(fh::cplx) * ((size_t)(2 * kl + ku + 1) * (size_t)n + (size_t)n);  // band + permuted right-hand side
}

extern "C" int fh_csr_band_lu_solve_c128(const int64_t *rowptr, const int32_t *colidx, const fh_c128 *vals,
                                         const fh_c128 *b, const int32_t *perm, int64_t n, int64_t kl, int64_t ku,
                                         fh_c128 *x, void *workspace, size_t workspace_bytes, int32_t *status,
                                         fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && kl >= 0 && ku >= 0, "n, kl, ku must be >= 0");
    if (n == 0) return FH_OK;
    FH_REQUIRE(rowptr && colidx && vals && b && x && workspace && status, "NULL device pointer");
    FH_REQUIRE(n < (1ll << 31) && kl < n && ku < n, "bandwidth must be smaller than n");
    FH_REQUIRE(workspace_bytes >= fh_band_lu_workspace_bytes(n, kl, ku), "workspace too small");
    const size_t smem = sizeof(cplx) * (size_t)(kl > 0 ? kl : 1);
    if (smem > smem_optin_bytes() - 1024) {
        set_error("lower bandwidth %lld needs %zu B of shared memory; use the dense solver", (long long)kl, smem);
        return FH_E_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cplx *AB = static_cast<cplx *>(workspace);
    cplx *rhs = AB + (size_t)(2 * kl + ku + 1) * (size_t)n;
    FH_CUDA(cudaMemsetAsync(AB, 0, sizeof(cplx) * (size_t)(2 * kl + ku + 1) * (size_t)n, st));
    const unsigned grid = (unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    csr_to_band_kernel<<<grid, 256, 0, st>>>(rowptr, colidx, reinterpret_cast<const cplx *>(vals),
                                             reinterpret_cast<const cplx *>(b), perm, n, (int)kl, (int)ku, AB, rhs);
    FH_LAUNCH_CHECK("csr_to_band_kernel");
    FH_CUDA(cudaFuncSetAttribute(band_lu_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    band_lu_solve_kernel<<<1, kBandThreads, smem, st>>>(AB, rhs, n, (int)kl, (int)ku, perm, reinterpret_cast<cplx *>(x),
                                                       status);
    FH_LAUNCH_CHECK("band_lu_solve_kernel");
    return FH_OK;
}
