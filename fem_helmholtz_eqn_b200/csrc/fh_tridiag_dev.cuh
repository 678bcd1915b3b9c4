// Device / host helpers of the cluster solver (fh_tridiag_batched.cu: 2 CTAs x 256 threads per system).
#pragma once
#include <cuda.h>
#include <string.h>

#include "fh_helm1d.cuh"

namespace fh {

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// spinning variant (test_wait does not suspend the warp): lower wake-up latency for waits that are known to be short
__device__ __forceinline__ void mbar_spin(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SPIN_%=:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra SPIN_%=;\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16 bytes into the partner CTA's shared memory, completing 16 tx-bytes on the partner's mbarrier
__device__ __forceinline__ void st_async_remote(uint32_t remote_addr, cplx v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(
                     remote_addr),
                 "d"(v.re), "d"(v.im), "r"(remote_bar)
                 : "memory");
}
// barrier over the first `threads` threads of the CTA (a whole number of warps), hardware barrier `id` (1..15)
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {  // counts towards barrier `id`, does not wait
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}

// A-posteriori accuracy probe.  The elimination is unpivoted; what it can lose shows up as a defect of the
// INTERFACE rows (row 7 of every chunk -- interior rows are satisfied by construction of the back-substitution;
// checked against a numpy model: the largest componentwise row residual always sits on an interface row and
// tracks the normwise backward error with correlation 0.92).  Every thread therefore keeps the original
// coefficients of its row 7 and re-evaluates that one equation with the computed solution; a relative defect above
// kProbeTol sets status bit 1, and the host refines exactly those systems (~2 % of a fine wavenumber sweep).
constexpr double kProbeTol = 5e-14;
// ... and above kSevereTol bit 0 as well (treated like a non-finite value: pivoted re-solve).  One refinement step
// squares the relative error of the unpivoted solve, so 1e-8 is where it stops being enough for 1e-13.
constexpr double kSevereTol = 1e-8;
__device__ __forceinline__ double abs1(cplx z) { return fabs(z.re) + fabs(z.im); }

// ---------------------------------------------------------------------------------------------- band sources
struct TensorMaps {  // built by the host per launch; live in kernel parameter space.  [cluster rank][dl, d, du, b]
    CUtensorMap band[2][4];
};
struct StoreMaps {   // solution tensor as seen by cluster rank 0 and rank 1 (each clipped to its own chunks)
    CUtensorMap r[2];
};

struct BandSource {  // staged pipeline: bands live in global memory
    const cplx *dl, *d, *du, *b;
    static constexpr bool kStaged = true;
    static constexpr bool kUniform = false;
    static constexpr bool kSubset = false;
    static constexpr bool kStoreBands = false;
    static constexpr bool kRefine = false;
};

// One step of iterative refinement of a SUBSET of a batch in one pass (fh_tridiag_refine_flagged_c128): compact system i
// is system list[i] of all five arrays, and the number of systems is read from *count (decided on the device, no host
// round trip).  A separate type, so that the ordinary solve does not carry the indirection.  The right-hand side of the correction solve, r = b - A u, is formed
// while the rows are read (each thread fetches the ten solution values its eight rows touch), the correction leaves
// as a TMA reduce-add onto u (cp.reduce.async.bulk.tensor .add, f64: u += e happens in L2), the small-pivot flag of the
// system is cleared and a breakdown of the correction solve sets the breakdown bit in the batch's own status array:
// 96 bytes of HBM traffic per refined row instead of the 240 of the residual / solve / update kernels it replaces, and
// bit-identical to them (same operation order in the residual, the same solve, the same IEEE addition).
struct BandRefineSource : BandSource {
    const int32_t *list;
    const int32_t *count;
    const cplx *u_old;  // the solutions being refined (the array the store maps address)
    static constexpr bool kSubset = true;
    static constexpr bool kRefine = true;
};

// fused pipeline: bands are generated, never stored.  UNIFORM = scalar-h mesh with at least two elements: every
// interior row of a system is the same, only the two mesh-boundary rows differ (compile-time switch: the generic
// generator, with its per-row geometry and IEEE divisions, stays out of the fast kernel).
// STORE: the generated bands are also written out (dl, d, du, b -- what fh_assemble_1d_c128 would have stored, bit for
// bit), so that "assemble, keep A, solve" costs 64 + 16 bytes per row of HBM traffic instead of 64 + 64 + 16.
template <bool UNIFORM, bool STORE = false>
struct HelmSourceT {
    Problem1d p;
    const double *ks, *k2s;
    cplx *out_dl, *out_d, *out_du, *out_b;  // STORE only
    static constexpr bool kStaged = false;
    static constexpr bool kUniform = UNIFORM;
    static constexpr bool kSubset = false;
    static constexpr bool kStoreBands = STORE;
    static constexpr bool kRefine = false;
};

struct HelmState {  // per-thread state of the generator
    RowGeom g_mid;  // scalar-h mode: geometry of the interior rows
    double k, k2;
    cplx il, id, iu, ib;  // the interior row of the current wavenumber (scalar-h mode), evaluated once per system
    cplx ia, ic;          // il / iu in the CTA's local row order (the reversed half swaps them)
};

// ---------------------------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Tensor view of one band / solution array: [batch][chunks][16 doubles], origin at row `row0` of system 0.
static inline int make_tensor_map(CUtensorMap *out, const cplx *base, int64_t n, int64_t batch, int row0, int chunks_total,
                                  int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return FH_E_UNSUPPORTED;
    }
    const cuuint64_t dims[3] = {16, (cuuint64_t)(chunks_total < 1 ? 1 : chunks_total), (cuuint64_t)batch};
    const cuuint64_t strides[2] = {128, (cuuint64_t)n * sizeof(cplx)};
    const cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    void *addr = const_cast<cplx *>(base) + row0;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, addr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (n=%lld batch=%lld)", (int)r, (long long)n,
                  (long long)batch);
        return FH_E_INVALID;
    }
    return FH_OK;
}


}  // namespace fh
