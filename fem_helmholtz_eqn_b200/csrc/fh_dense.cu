// Dense complex128 LU with partial pivoting on the GPU -- the stand-in for np.linalg.solve(K, F) of the
// 2D solver classes (fem2d.py:176 etc.: LAPACK zgesv on a dense n x n matrix) until the sparse 2D solver
//  lands.  Right-looking, one column per step, everything stays on the device:
//   pivot search (one CTA) -> row swap -> rank-1 update of the trailing block (2D grid) -> ... ->
//   back-substitution (one CTA).  O(n^3) flops, O(n^2) bytes; meant for the reference's mesh sizes (n <~ 10^4).
#include "fh_common.cuh"

namespace fh {

__global__ void lu_pivot_kernel(const cplx *__restrict__ A, int64_t n, int64_t k, int32_t *__restrict__ piv,
                                int32_t *__restrict__ status) {
    __shared__ double best[256];
    __shared__ int32_t arg[256];
    double b = -1.0;
    int32_t a = (int32_t)k;
    for (int64_t i = k + threadIdx.x; i < n; i += blockDim.x) {
        const cplx v = A[i * n + k];
        const double m = fabs(v.re) + fabs(v.im);  // LAPACK izamax measure: |re| + |im|
        if (m > b) b = m, a = (int32_t)i;
    }
    best[threadIdx.x] = b, arg[threadIdx.x] = a;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double o = best[threadIdx.x + s];
            const int32_t oa = arg[threadIdx.x + s];
            if (o > best[threadIdx.x] || (o == best[threadIdx.x] && oa < arg[threadIdx.x]))
                best[threadIdx.x] = o, arg[threadIdx.x] = oa;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        piv[k] = arg[0];
        if (!(best[0] > 0.0) || !isfinite(best[0])) atomicOr(status, 1);
    }
}

__global__ void lu_swap_kernel(cplx *__restrict__ A, cplx *__restrict__ b, int64_t n, int64_t k,
                               const int32_t *__restrict__ piv) {
    const int64_t p = piv[k];
    if (p == k) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        const cplx t = A[k * n + j];
        A[k * n + j] = A[p * n + j];
        A[p * n + j] = t;
    }
    if (j == 0) {
        const cplx t = b[k];
        b[k] = b[p];
        b[p] = t;
    }
}

// trailing update: rows i in (k, n), columns j in (k, n): A[i][j] -= (A[i][k] / A[k][k]) * A[k][j]
__global__ void lu_update_kernel(cplx *__restrict__ A, cplx *__restrict__ b, int64_t n, int64_t k) {
    __shared__ cplx lcol[16];
    const int64_t i = k + 1 + (int64_t)blockIdx.y * 16 + threadIdx.y;
    const int64_t j = k + 1 + (int64_t)blockIdx.x * 64 + threadIdx.x;
    const cplx inv = crecip(A[k * n + k]);
    if (threadIdx.x == 0 && i < n) lcol[threadIdx.y] = A[i * n + k] * inv;
    __syncthreads();
    if (i >= n) return;
    const cplx l = lcol[threadIdx.y];
    if (j < n) A[i * n + j] = nmsub(A[i * n + j], l, A[k * n + j]);
    if (blockIdx.x == 0 && threadIdx.x == 0) b[i] = nmsub(b[i], l, b[k]);
}

__global__ void lu_backsub_kernel(const cplx *__restrict__ A, const cplx *__restrict__ b, cplx *__restrict__ x,
                                  int64_t n, int32_t *__restrict__ status) {
    __shared__ double sre[256], sim[256];
    for (int64_t k = n - 1; k >= 0; --k) {
        double re = 0.0, im = 0.0;
        for (int64_t j = k + 1 + threadIdx.x; j < n; j += blockDim.x) {
            const cplx a = A[k * n + j], v = x[j];
            re += a.re * v.re - a.im * v.im;
            im += a.re * v.im + a.im * v.re;
        }
        sre[threadIdx.x] = re, sim[threadIdx.x] = im;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) sre[threadIdx.x] += sre[threadIdx.x + s], sim[threadIdx.x] += sim[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const cplx r = (b[k] - mk(sre[0], sim[0])) * crecip(A[k * n + k]);
            x[k] = r;
            if (!cfinite(r)) atomicOr(status, 1);
        }
        __syncthreads();
    }
}

}  // namespace fh

extern "C" int fh_dense_lu_solve_c128(fh_c128 *A, fh_c128 *b, fh_c128 *x, int64_t n, int32_t *piv, int32_t *status,
                                      fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return FH_OK;
    FH_REQUIRE(A && b && x && piv && status, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    cplx *a = reinterpret_cast<cplx *>(A), *rhs = reinterpret_cast<cplx *>(b);
    for (int64_t k = 0; k < n; ++k) {
        lu_pivot_kernel<<<1, 256, 0, st>>>(a, n, k, piv, status);
        lu_swap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, rhs, n, k, piv);
        const int64_t m = n - k - 1;
        if (m > 0) {
            dim3 grid((unsigned)((m + 63) / 64), (unsigned)((m + 15) / 16)), block(64, 16);
            lu_update_kernel<<<grid, block, 0, st>>>(a, rhs, n, k);
        }
    }
    FH_LAUNCH_CHECK("lu kernels");
    lu_backsub_kernel<<<1, 256, 0, st>>>(a, rhs, reinterpret_cast<cplx *>(x), n, status);
    FH_LAUNCH_CHECK("lu_backsub_kernel");
    return FH_OK;
}
