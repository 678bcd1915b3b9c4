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
#include <cooperative_groups.h>

#include "fh_common.cuh"

namespace cg = cooperative_groups;

namespace fh {

constexpr int kBandThreads = 1024;
constexpr int kBandCluster = 8;  // CTAs that share the columns of the update window (band_lu_cluster_kernel)

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

// The same elimination on a CLUSTER of 8 CTAs (round 2): a column step is bound by the rank-1 update of the
// kl x (kl + ku) window, ~30,000 read-modify-writes that one SM pushes through L2 in ~40 us.  Here every CTA finds the
// pivot and forms the multipliers itself (the same kl + 1 numbers read from L2 by all: no broadcast, no disagreement --
// same data, same code) and then interchanges and updates only the window columns c = rank (mod 8), one warp per column,
// rows across the lanes (contiguous).  One cluster barrier per column step makes the updated columns visible to the
// others (the band is read with ld.global.cg: another SM has written it).  The pivot column is left untouched -- the
// other CTAs may still be reading it -- and the reciprocal of the pivot goes to a side array for the back-substitution,
// which CTA 0 carries out as before.  Same pivot rule, same operations per entry: the solution has the bits of the
// single-CTA kernel.
__device__ __forceinline__ cplx ld_cg(const cplx *p) {
    const double2 v = __ldcg(reinterpret_cast<const double2 *>(p));
    return mk(v.x, v.y);
}

__global__ void __launch_bounds__(kBandThreads)
band_lu_cluster_kernel(cplx *__restrict__ AB, cplx *__restrict__ rhs, cplx *__restrict__ pivinv, int64_t n, int kl, int ku,
                       const int32_t *__restrict__ perm, cplx *__restrict__ x, int32_t *__restrict__ status, int rhs_in_smem) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    cplx *lcol = reinterpret_cast<cplx *>(band_smem);  // multipliers of the current column [kl]
    cplx *ccol = lcol + (kl > 0 ? kl : 1);             // the current column as read from L2 [kl + 1]
    cplx *xs = ccol + kl + 1;                          // back-substitution: the right-hand side / solution [n] (rhs_in_smem)
    __shared__ double wbest[kBandThreads / 32];
    __shared__ int warg[kBandThreads / 32];
    __shared__ int s_jp;
    __shared__ int s_bad;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ld = 2 * kl + ku + 1, kv = kl + ku;
    if (t == 0) s_bad = 0;

    for (int64_t j = 0; j < n; ++j) {
        const cplx *col = AB + j * ld;
        const int km = (int)min((int64_t)kl, n - 1 - j);
        const int cw = (int)min((int64_t)kv, n - 1 - j);
        // ---- pivot search (every CTA, identical); the column is kept in shared memory for the multipliers ----
        double best = -1.0;
        int arg = 0;
        for (int i = t; i <= km; i += kBandThreads) {
            const cplx z = ld_cg(col + kv + i);
            ccol[i] = z;
            const double m = cabs1(z);
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
                if (!(best > 0.0) || !isfinite(best)) s_bad = 1;
            }
        }
        __syncthreads();
        const int jp = s_jp;
        // ---- multipliers, with the interchange of rows j and j + jp applied on the fly ----
        const cplx inv = crecip(ccol[jp]);
        for (int i = 1 + t; i <= km; i += kBandThreads) lcol[i - 1] = ccol[i == jp ? 0 : i] * inv;
        if (rank == 0 && t == 0) pivinv[j] = inv;
        __syncthreads();
        // ---- this CTA's window columns: interchange + rank-1 update, one warp per column ----
        for (int c = 1 + rank + nranks * warp; c <= cw; c += nranks * (kBandThreads / 32)) {
            cplx *p = AB + (j + c) * ld + kv - c;  // A(j, j + c); A(j + i, j + c) is i entries further down
            const cplx a0 = ld_cg(p), ajp = ld_cg(p + jp);
            const cplx top = jp != 0 ? ajp : a0;   // A(j, j + c) after the interchange
            __syncwarp();  // (p[0] and p[jp] have been read by every lane before their owners overwrite them)
            for (int i0 = 1; i0 <= km; i0 += 128) {  // 128 rows at a time: four loads in flight per lane, then four stores
                cplx v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i0 + lane + 32 * q;
                    if (i <= km) v[q] = i == jp ? a0 : ld_cg(p + i);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i0 + lane + 32 * q;
                    if (i <= km) p[i] = nmsub(v[q], lcol[i - 1], top);
                }
            }
            if (lane == 0 && jp != 0) p[0] = top;
        }
        if (rank == nranks - 1) {  // the right-hand side: interchange + update
            const cplx b0 = ld_cg(rhs + j), bjp = ld_cg(rhs + j + jp);
            const cplx top = jp != 0 ? bjp : b0;
            __syncthreads();
            for (int i = 1 + t; i <= km; i += kBandThreads) {
                const cplx vi = i == jp ? b0 : ld_cg(rhs + j + i);
                rhs[j + i] = nmsub(vi, lcol[i - 1], top);
            }
            if (t == 0 && jp != 0) rhs[j] = top;
        }
        cluster.sync();  // the updated columns are visible to every CTA; lcol may be rewritten
    }
    if (rank != 0) return;
    if (rhs_in_smem) {
        // ---- back-substitution (CTA 0) with the right-hand side in shared memory: per column one barrier and no L2
        // round trip on the dependent chain (the column of U for step k - 1 is requested before step k's barrier) ----
        for (int64_t i = t; i < n; i += kBandThreads) xs[i] = ld_cg(rhs + i);
        __syncthreads();
        const int nper = (kv + kBandThreads - 1) / kBandThreads;  // entries of a column per thread: 1 for kv <= 1024
        cplx unext = czero();
        if (n > 0 && nper == 1 && 1 + t <= (int)min((int64_t)kv, n - 1)) unext = ld_cg(AB + (n - 1) * ld + kv - 1 - t);
        for (int64_t k = n - 1; k >= 0; --k) {
            const cplx *col = AB + k * ld;
            const int up = (int)min((int64_t)kv, k);
            const cplx ucur = unext;
            if (nper == 1 && k > 0 && 1 + t <= (int)min((int64_t)kv, k - 1)) unext = ld_cg(AB + (k - 1) * ld + kv - 1 - t);
            const cplx xk = xs[k] * ld_cg(pivinv + k);
            __syncthreads();  // everybody has read xs[k] before it is overwritten
            if (t == 0) {
                xs[k] = xk;
                if (!cfinite(xk)) s_bad = 1;
            }
            if (nper == 1) {
                if (1 + t <= up) xs[k - 1 - t] = nmsub(xs[k - 1 - t], ucur, xk);
            } else {
                for (int i = 1 + t; i <= up; i += kBandThreads) xs[k - i] = nmsub(xs[k - i], ld_cg(col + kv - i), xk);
            }
            __syncthreads();
        }
        for (int64_t i = t; i < n; i += kBandThreads) x[i] = xs[perm ? perm[i] : i];
        if (t == 0 && status != nullptr) *status = s_bad;
        return;
    }
    // ---- back-substitution (CTA 0), column oriented: x_k = b_k / U(k,k), then b_i -= U(i,k) x_k for the kv rows above ----
    for (int64_t k = n - 1; k >= 0; --k) {
        const cplx *col = AB + k * ld;
        const cplx xk = ld_cg(rhs + k) * ld_cg(pivinv + k);
        const int up = (int)min((int64_t)kv, k);
        __syncthreads();  // everybody has read rhs[k] before it is overwritten
        if (t == 0) {
            rhs[k] = xk;
            if (!cfinite(xk)) s_bad = 1;
        }
        for (int i = 1 + t; i <= up; i += kBandThreads) rhs[k - i] = nmsub(ld_cg(rhs + k - i), ld_cg(col + kv - i), xk);
        __syncthreads();
    }
    for (int64_t i = t; i < n; i += kBandThreads) x[i] = ld_cg(rhs + (perm ? perm[i] : i));
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
    // band + permuted right-hand side + reciprocal pivots
    return sizeof(fh::cplx) * ((size_t)(2 * kl + ku + 1) * (size_t)n + 2 * (size_t)n);
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
    static const bool single = [] { const char *e = getenv("FH_BAND"); return e != nullptr && e[0] == 's'; }();
    // A cluster of CTAs shares the update window when it is large enough to pay for a cluster barrier per column
    // (kl x (kl + ku) >= 4096 entries: measured, 225 nodes / (29, 29) is faster on one CTA); 16 CTAs where the device
    // allows the non-portable size, else 8.  Development switch: FH_BAND=single.
    if (!single && kl * (kl + ku) >= 4096) {
        cplx *pivinv = rhs + n;
        const size_t base = sizeof(cplx) * (size_t)((kl > 0 ? kl : 1) + kl + 1);
        const size_t want = base + sizeof(cplx) * (size_t)n;
        const int rhs_in_smem = want <= smem_optin_bytes() - 2048 ? 1 : 0;
        const size_t smem_c = rhs_in_smem ? want : base;
        FH_CUDA(cudaFuncSetAttribute(band_lu_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        static const int max_cluster = [] {
            if (cudaFuncSetAttribute(band_lu_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
                cudaGetLastError();
                return kBandCluster;
            }
            return 16;
        }();
        for (int csize = max_cluster; csize >= kBandCluster; csize /= 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)csize);
            cfg.blockDim = dim3(kBandThreads);
            cfg.dynamicSmemBytes = smem_c;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, band_lu_cluster_kernel, AB, rhs, pivinv, n, (int)kl, (int)ku, perm,
                                               reinterpret_cast<cplx *>(x), status, rhs_in_smem);
            if (e == cudaSuccess) return FH_OK;
            cudaGetLastError();  // (a cluster of 16 that cannot be scheduled with this much shared memory: try 8)
            if (csize == kBandCluster) {
                set_error("band_lu_cluster_kernel: %s", cudaGetErrorString(e));
                return (int)e;
            }
        }
    }
    FH_CUDA(cudaFuncSetAttribute(band_lu_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    band_lu_solve_kernel<<<1, kBandThreads, smem, st>>>(AB, rhs, n, (int)kl, (int)ku, perm, reinterpret_cast<cplx *>(x),
                                                       status);
    FH_LAUNCH_CHECK("band_lu_solve_kernel");
    return FH_OK;
}
