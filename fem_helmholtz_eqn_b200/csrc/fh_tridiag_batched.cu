// Batched complex128 tridiagonal solve for sm_100a (replaces nb:204, np.linalg.solve(A, b), for
// the tridiagonal A of nb:134-187) and the fused assemble+solve sweep (nb:190-205 for many k).
//
// Algorithm (per system)
//   1. thread-chunk partition: every thread owns kM = 8 consecutive rows in REGISTERS and removes
//      its interior couplings with a downward and an upward elimination sweep (pivots are kept as
//      reciprocals), leaving  a_i x_left + b_i x_i + c_i x_last = d_i  (i < 7) and one interface
//      row per thread;
//   2. the 256 interface rows of a CTA (unit diagonal) are solved in shared memory: one cyclic-reduction
//      step keeps the 128 odd rows; those are solved by <= 7 levels of parallel cyclic reduction with two
//      threads per row (even lane: couplings A and right-hand side D, odd lane: C and the cut column V);
//      the even rows follow from their own equation;
//   3. back-substitution in registers; the solution is staged in shared memory and leaves with one
//      TMA tensor store.
//   A system of up to 2056 rows is handled by one CTA; up to 4112 rows by a 2-CTA cluster: rank 0
//   takes the first half, rank 1 the last half in REVERSED row order, so both halves end at the cut;
//   for odd n the middle row (the "cut row") belongs to neither.  Each half carries one extra
//   right-hand side V (its coupling to the cut unknown / the partner's cut-side unknown); the two
//   (D, V) pairs are exchanged through distributed shared memory with st.async + mbarrier
//   complete_tx (no cluster-wide barrier, no fence) and both CTAs solve the small cut system.
//
// Data movement (staged variant): the CTA is persistent.  While system s is processed, the four band
// tiles of system s + stride are in flight: one elected thread issues four TMA tensor loads
// (cp.async.bulk.tensor.3d, 32 KB each, 128-byte swizzle) that complete on an mbarrier.  The tensor
// view of a band array is [system][chunk of 8 rows][8 complex]; the hardware swizzle makes the
// per-thread chunk reads (stride 128 B) bank-conflict free; they are interleaved with the downward
// sweep.  Rows that do not fill a chunk (at most 8 "head" rows per half, none for n = 4097) travel by
// cp.async and are eliminated serially by one thread.  HBM traffic = algorithmic traffic: 64 B/row
// read + 16 B/row written (ncu: profiles/).
//
// Fused variant: the same kernel with the band loads replaced by the row generator of
// fh_helm1d.cuh (bit-identical to what fh_assemble_1d_c128 would have stored); 16 B/row written.
// Refinement variant (BandRefineSource, pipeline kernel only): systems picked through a device index list, the residual
// formed while the rows are read, the correction added to u -- fh_tridiag_refine_flagged_c128.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "fh_tridiag_dev.cuh"

namespace cg = cooperative_groups;

namespace fh {

constexpr int kM = 8;                          // rows per thread (register chunk)
constexpr int kT = 2048 / kM;                  // threads per CTA: 256
constexpr int kSlots = kT * kM;                // 2048 body rows per CTA
constexpr int kChunks8 = kSlots / 8;           // 8-row memory chunks (TMA tile rows) per CTA: 256
constexpr int kPer8 = 8 / kM;                  // threads sharing one memory chunk: 1
// (a 4-row / 512-thread variant was measured earlier in the round: twice the interface rows, no faster, and its
// interface arrays no longer fit next to the staging tiles)
constexpr int kHeadMax = 8;                    // extra rows at the far end, eliminated serially
constexpr int kStageRows = kSlots + kHeadMax;  // 2056
constexpr int kLevelsMax = 8;                  // log2(kT)
constexpr uint32_t kTileBytes = kSlots * sizeof(cplx);  // 32 KB per band tile

// single-thread duties are spread over warps (whichever warp carries one is late for the next barrier): the band
// prefetch is issued by the first active thread (warp 0 for full systems), the solution store by warp 1, the exchange
// barrier is armed by warp 2
constexpr int kStoreThread = 32, kXchgThread = 64;
constexpr int kBandStoreThread0 = 96;  // kStoreBands: lanes 0 of warps 3..6 store one band each
constexpr int kBandTileChunks = 128;   // ... from a tile of 128 identical chunks (16 KB)
constexpr uint32_t kBandTileBytes = kBandTileChunks * 128u;
constexpr int kH = kT / 2;              // interface rows that survive the cyclic-reduction step
constexpr int kHalfPad = kH / 2;        // zero rows on either side of a half-width array (largest PCR stride)
constexpr int kHalfStride = kH + 2 * kHalfPad + 4;  // (+4 rows = 64 B: the two arrays a lane pair touches together
                                                    // fall on different halves of the 32 banks)

struct __align__(1024) BatchedSmem {
    cplx stage[4][kSlots];      // TMA destination: bands of the NEXT system (a, b, c, d), swizzle-128B
    cplx xs[kSlots];            // TMA source: solution of the CURRENT system, swizzle-128B
    cplx row0[3][kT];           // normalised first rows of the chunks (read by the thread above)
    cplx full[4][kT + 1];       // interface rows A, C, D, V; entry kT is a zero row (and staggers the banks)
    cplx half[2][4][kHalfStride];  // PCR ping-pong over the odd interface rows, zero rows on both sides
    cplx headst[4][kHeadMax];   // head rows of the next system (cp.async by one thread)
    cplx xchg[2][2];            // [parity][{Y, V}] written by the partner CTA (st.async)
    RowGeom bgeom[2];           // fused variant, scalar-h: geometry of the first / last row of the mesh (k-independent)
    unsigned long long full_bar;     // mbarrier: band tiles have landed
    unsigned long long xchg_bar[2];  // mbarrier: partner's interface pair has landed
};

// One PCR / CR node: 1 / (1 - A*Cm - C*Ap) is kept as conj(den) and 1/|den|^2, so that the products with
// conj(den) run while the reciprocal is still in flight (the chain is den -> |den|^2 -> rcp -> one multiply).
struct PcrPivot {
    cplx cden;
    double s;
};
__device__ __forceinline__ PcrPivot pcr_pivot(cplx A, cplx Cm, cplx C, cplx Ap) {
    double t1r = fma(-A.re, Cm.re, 1.0);
    t1r = fma(A.im, Cm.im, t1r);
    double t2r = C.re * Ap.re;
    t2r = fma(-C.im, Ap.im, t2r);
    double t1i = A.re * Cm.im;
    t1i = fma(A.im, Cm.re, t1i);
    double t2i = C.re * Ap.im;
    t2i = fma(C.im, Ap.re, t2i);
    PcrPivot p;
    p.cden = mk(t1r - t2r, t1i + t2i);  // conj(den): den.im = -(t1i + t2i)
    p.s = fast_rcp(fma(p.cden.re, p.cden.re, p.cden.im * p.cden.im));
    return p;
}
__device__ __forceinline__ cplx pcr_scale(cplx num, const PcrPivot &p) {  // num / den
    const cplx m = num * p.cden;
    return mk(m.re * p.s, m.im * p.s);
}

// ---------------------------------------------------------------------------------------------- row -> slot maps
// How the rows of one system map onto the CTAs.  Single CTA: all n rows.  Cluster of two: rank 0 takes the first
// m = n / 2 rows, rank 1 the LAST m rows in reversed order; for odd n the row in the middle (the "cut row") belongs
// to neither: both halves couple to its unknown, which is found together with the exchange (one 4097-row system =
// 2 x 2048 + 1: no leftover rows to eliminate serially -- those cost the reversed CTA 1000 cycles per system).
// Inside a CTA, local rows run from the far end (j = 0) to the cut (j = n_loc - 1).  The first `head` rows are "head
// rows"; the other 8 * chunks rows fill the chunks of threads kT - kPer8 * chunks .. kT - 1 (right-aligned).
struct LocalMap {
    int n_loc, head, chunks;
    int first_active;  // kT - kPer8 * chunks
    int64_t g0;        // global row of local row 0
    int dir;           // +1 forward, -1 reversed
    int box_c1;        // chunk coordinate at which this CTA's TMA LOAD box starts (may be negative)
    __device__ __forceinline__ int64_t global_row(int j) const { return g0 + (int64_t)dir * j; }
    // swizzle-128B slot of element i of thread t's chunk in the STORE tile: TMA stores trap on negative coordinates
    // (measured, tools/tma_probe.cu), so the tile is left-aligned (box starts at chunk 0; the tensor map clips)
    __device__ __forceinline__ int xs_slot(int t, int i) const {
        const int tt = dir > 0 ? t - first_active : kT - 1 - t;
        const int r = tt / kPer8;
        const int e = (dir > 0 ? i : kM - 1 - i) + kM * (tt % kPer8);
        return r * 8 + (e ^ (r & 7));
    }
};

struct HostMap {  // the same numbers, computed on the host for the tensor maps
    int n_loc[2], head[2], chunks[2];
    int row0[2];  // first row of the chunk grid of each rank (origin of its tensor maps)
    int cut;      // 1: row n_loc[0] is the cut row
    // does the first / last chunk of a rank's grid hold a mesh-boundary row (row 0 / row n - 1)?  (They are the only
    // chunks of a uniform mesh whose rows are not all alike: the one-pass assemble+solve writes them separately.)
    int front[2], back[2];
};
__host__ __device__ inline HostMap host_map(int64_t n, int cluster) {
    HostMap m;
    m.cut = cluster == 2 ? (int)(n & 1) : 0;
    m.n_loc[0] = cluster == 1 ? (int)n : (int)(n / 2);
    m.n_loc[1] = cluster == 1 ? 0 : (int)(n / 2);
    for (int r = 0; r < 2; ++r) {
        int c = m.n_loc[r] / 8;          // 8-row memory chunks
        if (c > kChunks8) c = kChunks8;
        m.chunks[r] = c;
        m.head[r] = m.n_loc[r] - c * 8;  // 0..8
    }
    m.row0[0] = m.head[0];               // rank 0: head rows first, then the chunks
    m.row0[1] = m.n_loc[0] + m.cut;      // rank 1 (reversed): chunks first in memory, head rows at the far end
    m.front[0] = m.head[0] == 0 && m.chunks[0] > 0;
    m.back[0] = cluster == 1 && m.chunks[0] > 0;
    m.front[1] = 0;
    m.back[1] = cluster == 2 && m.head[1] == 0 && m.chunks[1] > 0;
    return m;
}

template <int CLUSTER>
__device__ __forceinline__ LocalMap make_map(int64_t n, int rank) {
    const HostMap h = host_map(n, CLUSTER);
    LocalMap m;
    m.n_loc = h.n_loc[rank], m.head = h.head[rank], m.chunks = h.chunks[rank];
    m.first_active = kT - kPer8 * m.chunks;  // in threads
    if (rank == 0) {
        m.g0 = 0, m.dir = 1;
        m.box_c1 = m.chunks - kChunks8;  // tile row r <-> memory chunk r + box_c1 (right-aligned)
    } else {
        m.g0 = n - 1, m.dir = -1;
        m.box_c1 = 0;                    // thread t <-> tile row kT-1-t <-> memory chunk kT-1-t of rank 1's grid
    }
    return m;
}

// The two boundary rows of a system are evaluated out of line: inlining their geometry (two element matrices with
// IEEE divisions) into every unrolled row slot made the fused kernel 6,900 SASS instructions long and
// instruction-cache bound (ncu: 8 % "no instruction" stalls).
struct Row4 {
    cplx lo, b, hi, d;
};
__device__ __noinline__ Row4 helm_boundary_row(const Problem1d &p, int64_t g, double k, double k2) {
    Row4 r;  // returned by value: taking addresses of the caller's register arrays would push them to local memory
    row_values(p, row_geom(p, g), g, k, k2, r.lo, r.b, r.hi, r.d);
    return r;
}

template <class Source>
__device__ __forceinline__ void helm_row(const Source &src, const HelmState &st, const LocalMap &m, int j, cplx &a,
                                         cplx &b, cplx &c, cplx &d) {
    if constexpr (!Source::kStaged) {
        const int64_t g = m.global_row(j);
        cplx lo, hi;
        if (src.p.h_mode == FH_H_SCALAR) {
            lo = st.il, b = st.id, hi = st.iu, d = st.ib;  // interior rows are all alike (same bits as row_values)
            if (g == 0 || g == src.p.n_elem) {
                const Row4 r = helm_boundary_row(src.p, g, st.k, st.k2);
                lo = r.lo, b = r.b, hi = r.hi, d = r.d;
            }
        } else {  // nodal-h: every row has its own geometry
            const Row4 r = helm_boundary_row(src.p, g, st.k, st.k2);
            lo = r.lo, b = r.b, hi = r.hi, d = r.d;
        }
        a = m.dir > 0 ? lo : hi;
        c = m.dir > 0 ? hi : lo;
    }
}

// Development aid (-DFH_TRACE): warp leaders of the first cluster stamp clock64 / globaltimer at the phase
// boundaries of one iteration; tools/trace_phases.py prints the phase lengths.
#ifdef FH_TRACE
__device__ long long *g_trace = nullptr;
__device__ int g_trace_iter = 0;
#define FH_TR(k)                                                                                              \
    do {                                                                                                      \
        if (g_trace != nullptr && it == (uint32_t)g_trace_iter && blockIdx.x < 2 && (t & 31) == 0) {          \
            unsigned long long gt__;                                                                          \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__));                                          \
            g_trace[((blockIdx.x * 8 + (t >> 5)) * 16 + (k)) * 2] = clock64();                                \
            g_trace[((blockIdx.x * 8 + (t >> 5)) * 16 + (k)) * 2 + 1] = (long long)gt__;                      \
        }                                                                                                     \
    } while (0)
#else
#define FH_TR(k) do { } while (0)
#endif

// any single row in GLOBAL terms (lo = sub-diagonal entry, hi = super-diagonal entry): the cut row of the fused variant
template <class Source>
__device__ __forceinline__ Row4 helm_any_row(const Source &src, const HelmState &st, int64_t g) {
    Row4 r;
    r.lo = r.b = r.hi = r.d = czero();
    if constexpr (!Source::kStaged) {
        if (src.p.h_mode == FH_H_SCALAR && g != 0 && g != src.p.n_elem) {
            r.lo = st.il, r.b = st.id, r.hi = st.iu, r.d = st.ib;
        } else {
            r = helm_boundary_row(src.p, g, st.k, st.k2);
        }
    }
    return r;
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int CLUSTER, class Source>
__global__ void __launch_bounds__(kT, 1)
tridiag_batched_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ StoreMaps umaps,
                       Source src, cplx *__restrict__ u, int64_t n, int64_t batch, int levels,
                       int32_t *__restrict__ status) {
    extern __shared__ unsigned char smem_raw[];
    // the dynamic window is only guaranteed 16-byte aligned: round up to the 1024 B the swizzle needs
    BatchedSmem &sm = *reinterpret_cast<BatchedSmem *>(
        smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    const int t = threadIdx.x;
    int rank = 0;
    int64_t first_sys = blockIdx.x, sys_stride = gridDim.x;
    if (CLUSTER > 1) {
        rank = (int)cg::this_cluster().block_rank();
        first_sys = blockIdx.x / CLUSTER;
        sys_stride = gridDim.x / CLUSTER;
    }
    const LocalMap map = make_map<CLUSTER>(n, rank);
    const bool active = t >= map.first_active;

    // reversed half: sub- and super-diagonal swap roles
    const CUtensorMap *tmaps = maps.band[rank];
    const CUtensorMap *map_lo = map.dir > 0 ? &tmaps[0] : &tmaps[2];
    const CUtensorMap *map_hi = map.dir > 0 ? &tmaps[2] : &tmaps[0];
    const bool has_cut = CLUSTER > 1 && (n & 1);
    const int64_t cut_row = n / 2;  // (global row of the cut row when has_cut)

    // zero rows that the half-width PCR reads beyond either end of the interface system (never written again)
    for (int i = t; i < 2 * 4 * kHalfStride; i += kT) {
        const int c = i % kHalfStride;
        if (c < kHalfPad || c >= kHalfPad + kH) (&sm.half[0][0][0])[i] = czero();
    }
    if (t < 4) sm.full[t][kT] = czero();
    if (t == 0) {
        mbar_init(&sm.full_bar, 1);
        mbar_init(&sm.xchg_bar[0], 1);
        mbar_init(&sm.xchg_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CLUSTER > 1) cg::this_cluster().sync();  // partner's barriers exist before anything is sent
    else __syncthreads();

    HelmState hs;
    if constexpr (!Source::kStaged) {
        if (src.p.h_mode == FH_H_SCALAR) hs.g_mid = row_geom(src.p, src.p.n_elem > 1 ? 1 : 0);
        if constexpr (Source::kUniform) {
            // the two boundary rows have their own geometry (IEEE divisions): evaluated once, not once per system
            if (t < 2) sm.bgeom[t] = row_geom(src.p, t == 0 ? 0 : src.p.n_elem);
            __syncthreads();
        }
    }

    // system index of the MATRIX of (compact) system `sys`: they differ when a subset is refined
    auto matrix_of = [&](int64_t sys) -> int64_t {
        if constexpr (Source::kSubset) return src.list[sys];
        else return sys;
    };
    if constexpr (Source::kSubset) batch = *src.count < batch ? *src.count : batch;
    // one elected thread feeds the pipeline: four band tiles by TMA, head rows by cp.async
    auto prefetch = [&](int64_t sys) {
        if constexpr (Source::kStaged) {
            if (t == map.first_active % kT) {  // the thread that also consumes the head rows
                const int64_t sysm = matrix_of(sys);
                mbar_expect_tx(&sm.full_bar, 4 * kTileBytes);
                tma_load_3d(sm.stage[0], map_lo, 0, map.box_c1, (int)sysm, &sm.full_bar);
                tma_load_3d(sm.stage[1], &tmaps[1], 0, map.box_c1, (int)sysm, &sm.full_bar);
                tma_load_3d(sm.stage[2], map_hi, 0, map.box_c1, (int)sysm, &sm.full_bar);
                tma_load_3d(sm.stage[3], &tmaps[3], 0, map.box_c1, (int)sys, &sm.full_bar);
                const cplx *lo = map.dir > 0 ? src.dl : src.du, *hi = map.dir > 0 ? src.du : src.dl;
                for (int j = 0; j < map.head; ++j) {
                    const int64_t g = sysm * n + map.global_row(j);
                    cp_async16(&sm.headst[0][j], lo + g);
                    cp_async16(&sm.headst[1][j], src.d + g);
                    cp_async16(&sm.headst[2][j], hi + g);
                    cp_async16(&sm.headst[3][j], src.b + sys * n + map.global_row(j));
                }
                cp_async_commit();
            }
        }
    };
    const int feeder = map.first_active % kT;  // == first_active, or 0 when the CTA has no chunk at all
    if (first_sys < batch) prefetch(first_sys);

    uint32_t it = 0;
    for (int64_t sys = first_sys; sys < batch; sys += sys_stride, ++it) {
        const uint32_t parity = it & 1;
        if constexpr (!Source::kStaged) {
            hs.k = __ldg(src.ks + sys), hs.k2 = __ldg(src.k2s + sys);
            if (src.p.h_mode == FH_H_SCALAR && src.p.n_elem >= 2)
                row_values(src.p, hs.g_mid, 1, hs.k, hs.k2, hs.il, hs.id, hs.iu, hs.ib);
        }
        FH_TR(0);
        if (CLUSTER > 1 && t == kXchgThread) mbar_expect_tx(&sm.xchg_bar[parity], 2 * sizeof(cplx));

        // ---- bands -> registers, interleaved with phase 1 -------------------------------------------------------
        // Phase 1 is the downward elimination inside the chunk; b[i] (i < kM - 1) becomes 1 / pivot.  Row i + 2 is
        // requested from the staging tiles before step i is computed: issuing all 4 * kM tile reads up front parked
        // every warp behind the shared-memory queue (1024 wavefronts per CTA) before its first multiply, and the two
        // phases ran one after the other (2300 cycles) instead of overlapped.
        cplx a[kM], b[kM], c[kM], d[kM];
        uint32_t tile_addr = 0, tile_x = 0;
        if constexpr (Source::kStaged) {
            mbar_wait(&sm.full_bar, parity);
            if (map.head > 0 && t == feeder) cp_async_wait_all();  // the head rows have landed too
            const int tt = map.dir > 0 ? t : kT - 1 - t;  // thread chunk in memory order
            const int r = tt / kPer8;                     // 8-row memory chunk = TMA tile row
            tile_addr = smem_u32(&sm.stage[0][0]) + (uint32_t)r * 128u;
            tile_x = (uint32_t)(r & 7);
        }
        FH_TR(1);
        // fused, uniform mesh: the interior row in local order, and the CTA's two possible mesh-boundary rows --
        // evaluated by every thread (a dozen operations), so that the one thread that owns such a row does not diverge
        cplx ia, ic, fa, fb, fc, fd, la, lb, lc, ld;
        if constexpr (Source::kUniform) {
            ia = map.dir > 0 ? hs.il : hs.iu, ic = map.dir > 0 ? hs.iu : hs.il;
            cplx lo, hi;
            const int64_t g = map.global_row(0);  // local row 0: mesh row 0 (forward) or n - 1 (reversed)
            row_values(src.p, sm.bgeom[g == 0 ? 0 : 1], g, hs.k, hs.k2, lo, fb, hi, fd);
            fa = map.dir > 0 ? lo : hi, fc = map.dir > 0 ? hi : lo;
            if (CLUSTER == 1) row_values(src.p, sm.bgeom[1], n - 1, hs.k, hs.k2, la, lb, lc, ld);
        }
        // A is kept (kStoreBands): on a uniform mesh every interior row of a system holds the same four numbers, so
        // what is written need not come from the rows a thread eliminates.  Four 16 KB tiles (128 chunks of 8 identical
        // rows: dl, d, du, b) are filled once per system and sent to every block of 128 interior chunks by TMA stores
        // (bigger blocks = fewer store operations: the TMA engine's cost per operation was measured to dominate with
        // 4 KB blocks; smaller tiles = fewer shared-memory writes than staging the rows one by one);
        // the few rows that differ or fall outside the chunk grid (the chunks holding mesh row 0 / n - 1, head rows, the
        // cut row) are written by single lanes of warp 0.  Values = fh_assemble_1d_c128's, bit for bit (same generator).
        if constexpr (Source::kStoreBands) {
            const HostMap hm = host_map(n, CLUSTER);
            const uint32_t mini = smem_u32(&sm.stage[0][0]) + (uint32_t)t * 16u;   // [4 bands][128 chunks][128 B]
#pragma unroll
            for (uint32_t rep = 0; rep < kBandTileBytes; rep += kT * 16u) {
                sts_volatile(mini + rep, hs.il), sts_volatile(mini + kBandTileBytes + rep, hs.id);
                sts_volatile(mini + 2 * kBandTileBytes + rep, hs.iu), sts_volatile(mini + 3 * kBandTileBytes + rep, hs.ib);
            }
            fence_proxy_async();  // (the stores are issued behind the block barrier below)
            if (t < 32) {
                cplx r0l, r0d, r0u, r0b, rNl, rNd, rNu, rNb;
                row_values(src.p, sm.bgeom[0], 0, hs.k, hs.k2, r0l, r0d, r0u, r0b);
                row_values(src.p, sm.bgeom[1], n - 1, hs.k, hs.k2, rNl, rNd, rNu, rNb);
                int64_t g = -1;  // the special row of this lane, if any
                const int64_t grid0 = hm.row0[rank];
                if (t < map.head) g = map.global_row(t);                                        // lanes 0..7: head rows
                if (t == 8 && has_cut && rank == 0) g = cut_row;                                // lane 8: the cut row
                if (t >= 16 && t < 24 && hm.front[rank]) g = grid0 + (t - 16);                  // first chunk
                if (t >= 24 && hm.back[rank]) g = grid0 + 8 * (int64_t)(map.chunks - 1) + (t - 24);  // last chunk
                if (g >= 0) {
                    cplx vl = hs.il, vd = hs.id, vu = hs.iu, vb = hs.ib;
                    if (g == 0) vl = r0l, vd = r0d, vu = r0u, vb = r0b;
                    if (g == n - 1) vl = rNl, vd = rNd, vu = rNu, vb = rNb;
                    const int64_t o = sys * n + g;
                    stg_stream(src.out_dl + o, vl), stg_stream(src.out_d + o, vd);
                    stg_stream(src.out_du + o, vu), stg_stream(src.out_b + o, vb);
                }
            }
        }
        if constexpr (Source::kStoreBands) {
            // The stores are issued NOW, behind one extra block barrier: the TMA engine then reads the tiles while the
            // CTA is busy with the register-only elimination phases.  (Issued after phase 2 the 128 KB of tile reads
            // shared the shared-memory pipe with the interface solve and slowed every one of its levels.)
            __syncthreads();
            if (t >= kBandStoreThread0 && t < kBandStoreThread0 + 128 && (t & 31) == 0) {
                const HostMap hm = host_map(n, CLUSTER);
                const int q = (t - kBandStoreThread0) >> 5;
                const int count = map.chunks - hm.front[rank] - hm.back[rank];
                for (int c = 0; c < count; c += kBandTileChunks)
                    tma_store_3d(&maps.band[rank][q], 0, c, (int)sys, &sm.stage[0][q * (kBandTileBytes / 16)]);
                tma_store_commit();
            }
        }
        auto load_row = [&](int i) {
            if (!active) {
                a[i] = czero(), b[i] = cone(), c[i] = czero(), d[i] = czero();
            } else if constexpr (Source::kStaged) {
                const int tt = map.dir > 0 ? t : kT - 1 - t;
                const uint32_t e = (uint32_t)((map.dir > 0 ? i : kM - 1 - i) + kM * (tt % kPer8));
                const uint32_t ad = tile_addr + ((e ^ tile_x) << 4);
                a[i] = lds_volatile(ad), b[i] = lds_volatile(ad + kTileBytes);
                c[i] = lds_volatile(ad + 2 * kTileBytes), d[i] = lds_volatile(ad + 3 * kTileBytes);
            } else if constexpr (Source::kUniform) {
                // all interior rows are alike; a mesh boundary row can only sit in the first slot of the first active
                // thread (local row 0, when there are no head rows) or, single-CTA, in the last slot of the last thread
                a[i] = ia, b[i] = hs.id, c[i] = ic, d[i] = hs.ib;
                if (i == 0 && map.head == 0 && t == map.first_active) a[i] = fa, b[i] = fb, c[i] = fc, d[i] = fd;
                if (i == kM - 1 && CLUSTER == 1 && t == kT - 1) a[i] = la, b[i] = lb, c[i] = lc, d[i] = ld;
            } else {
                helm_row(src, hs, map, map.head + (t - map.first_active) * kM + i, a[i], b[i], c[i], d[i]);
            }
            if (i == 0 && map.head == 0 && t == map.first_active) a[0] = czero();  // local row 0: nothing above it
            if (i == kM - 1 && CLUSTER == 1 && t == kT - 1) c[kM - 1] = czero();     // last row of the system
        };
        load_row(0);
        load_row(1);

        // ---- head rows (far end, at most 8): serial Thomas by one thread, folded into its row 0 ------
        cplx hcp[kHeadMax], hdp[kHeadMax];
        if (map.head > 0 && t == feeder) {
            cplx cp = czero(), dp = czero();
#pragma unroll
            for (int j = 0; j < kHeadMax; ++j) {
                if (j < map.head) {
                    cplx ha, hb, hc, hd;
                    if constexpr (Source::kStaged) {
                        ha = sm.headst[0][j], hb = sm.headst[1][j], hc = sm.headst[2][j], hd = sm.headst[3][j];
                    } else {
                        helm_row(src, hs, map, j, ha, hb, hc, hd);
                    }
                    if (j == 0) ha = czero();
                    if (CLUSTER == 1 && j == map.n_loc - 1) hc = czero();  // system shorter than one chunk
                    const cplx inv = frecip(nmsub(hb, ha, cp));
                    dp = nmsub(hd, ha, dp) * inv;
                    cp = hc * inv;
                    hcp[j] = cp, hdp[j] = dp;
                }
            }
            if (map.chunks > 0) {
                b[0] = nmsub(b[0], a[0], cp);
                d[0] = nmsub(d[0], a[0], dp);
                a[0] = czero();
            }
        }
        cplx a7o, b7o, c7o, d7o;  // the original interface row, for the probe
        b[0] = frecip(b[0]);
#pragma unroll
        for (int i = 1; i < kM; ++i) {
            if (i + 1 < kM) load_row(i + 1);
            if (i == kM - 1) a7o = a[i], b7o = b[i], c7o = c[i], d7o = d[i];
            const cplx w = a[i] * b[i - 1];
            b[i] = nmsub(b[i], w, c[i - 1]);
            d[i] = nmsub(d[i], w, d[i - 1]);
            a[i] = -(w * a[i - 1]);
            if (i < kM - 1) b[i] = frecip(b[i]);
        }
        FH_TR(2);
        if constexpr (Source::kStaged) {
            __syncthreads();  // every thread has drained the staging tiles ...
            if (sys + sys_stride < batch) prefetch(sys + sys_stride);  // ... refill them
        }
        FH_TR(3);
        // ---- phase 2: upward elimination; afterwards row i < 7 couples to x_left and x_7 only -------
#pragma unroll
        for (int i = kM - 3; i >= 0; --i) {
            const cplx w = c[i] * b[i + 1];
            a[i] = nmsub(a[i], w, a[i + 1]);
            d[i] = nmsub(d[i], w, d[i + 1]);
            c[i] = -(w * c[i + 1]);
        }
        // ---- interface row: eliminate the next chunk's first unknown with its (normalised) row 0 -----
        FH_TR(4);
        sm.row0[0][t] = a[0] * b[0];
        sm.row0[1][t] = c[0] * b[0];
        sm.row0[2][t] = d[0] * b[0];
        __syncthreads();
        FH_TR(5);
        cplx rA, rC, rD, rV;
        {
            cplx rb = b[kM - 1];
            rA = a[kM - 1], rC = czero(), rD = d[kM - 1], rV = czero();
            if (t < kT - 1) {
                const cplx nA = sm.row0[0][t + 1], nC = sm.row0[1][t + 1], nD = sm.row0[2][t + 1];
                rb = nmsub(rb, c[kM - 1], nA);
                rC = -(c[kM - 1] * nC);
                rD = nmsub(rD, c[kM - 1], nD);
            } else if (CLUSTER > 1) {
                rV = c[kM - 1];  // coupling to the partner CTA's interface unknown
            }
            const cplx inv = frecip(rb);
            rA = rA * inv, rC = rC * inv, rD = rD * inv, rV = rV * inv;
        }
        // ---- interface system: kT unit-diagonal rows  x_t + A x_{t-1} + C x_{t+1} = D - V x_partner ----------------
        // The full-width 8-level PCR was bound by the fp64 pipe (70 DP instructions per row and level) and by
        // shared-memory bandwidth.  Now: (1) ONE cyclic-reduction step keeps the odd rows only (the last row, which
        // holds the cut unknown, is odd); (2) the kT/2 survivors are solved by PCR with TWO threads per row: role 0
        // updates {D, A}, role 1 updates {V, C} -- the same instruction stream, operands picked by array index, so
        // there is no divergence; (3) the even rows follow from their own equation, which never left the registers.
        sm.full[0][t] = rA, sm.full[1][t] = rC, sm.full[2][t] = rD, sm.full[3][t] = rV;
        FH_TR(13);
        if (t == kStoreThread) tma_store_wait_read();  // the previous system's TMA store has finished reading xs
        FH_TR(14);
        __syncthreads();
        FH_TR(6);
        // The cut row (lo x(rank 0 side) + diag x_cut + hi x(rank 1 side) = rhs) is needed right after the exchange:
        // requested here (its sectors were just fetched by the tile loads of one of the two CTAs: an L2 hit), so the
        // latency hides behind the PCR.
        cplx cl = czero(), cb = cone(), ch = czero(), cd = czero();
        if (has_cut) {
            if constexpr (Source::kStaged) {
                const int64_t g = matrix_of(sys) * n + cut_row;
                cl = ldg_stream(src.dl + g), cb = ldg_stream(src.d + g), ch = ldg_stream(src.du + g);
                cd = ldg_stream(src.b + sys * n + cut_row);
            } else {
                const Row4 r = helm_any_row(src, hs, cut_row);
                cl = r.lo, cb = r.b, ch = r.hi, cd = r.d;
            }
        }
        // Thread pair (t, t ^ 1) owns the odd row r = t | 1, compacted index j = t >> 1: role 0 (even lane) updates
        // {A, D}, role 1 (odd lane) updates {C, V}.  "own" = the coupling this thread updates, "oth" = its partner's
        // (one shuffle away).  Both threads evaluate den = 1 - A * C(j-s) - C * A(j+s) in the SAME operand order: their
        // pivots must agree to the last bit, or the row ends up scaled inconsistently (that cost two digits of
        // backward error near small pivots).  The pair reads the same neighbour couplings (one wavefront serves both).
        int cur = 0;
        const int role = t & 1, j = t >> 1, va = 2 + role;
        cplx own, xv;
        {   // level 0 (cyclic reduction), from the full-width arrays; row kT is a zero row
            const int r = t | 1;
            const cplx oA = sm.full[0][r], oC = sm.full[1][r], oX = sm.full[va][r];
            const cplx Cm = sm.full[1][r - 1], Ap = sm.full[0][r + 1];
            const cplx Xm = sm.full[va][r - 1], Xp = sm.full[va][r + 1];
            const cplx Qn = role ? sm.full[1][r + 1] : sm.full[0][r - 1];
            const PcrPivot pv = pcr_pivot(oA, Cm, oC, Ap);
            xv = pcr_scale(nmsub(nmsub(oX, oA, Xm), oC, Xp), pv);
            own = -pcr_scale((role ? oC : oA) * Qn, pv);
        }
        constexpr uint32_t kArr = kHalfStride * sizeof(cplx);
        const uint32_t hb0 = smem_u32(&sm.half[0][0][kHalfPad + j]), hb1 = smem_u32(&sm.half[1][0][kHalfPad + j]);
        const uint32_t o_val = va * kArr, o_own = role * kArr;
        // the partner lane's coupling travels by shuffle, issued BEFORE the barrier so that its latency hides behind it
        cplx oth = mk(__shfl_xor_sync(0xffffffffu, own.re, 1), __shfl_xor_sync(0xffffffffu, own.im, 1));
        sts_volatile(hb0 + o_val, xv);
        sts_volatile(hb0 + o_own, own);
        __syncthreads();
        FH_TR(7);
#pragma unroll 1
        for (int lv = 0, sb = (int)sizeof(cplx); lv < levels; ++lv, sb <<= 1) {  // rows beyond either end: zero rows
            const uint32_t rd = cur ? hb1 : hb0, wr = cur ? hb0 : hb1;
            const cplx Cm = lds_volatile(rd + kArr - sb), Ap = lds_volatile(rd + sb);
            const cplx Xm = lds_volatile(rd + o_val - sb), Xp = lds_volatile(rd + o_val + sb);
            const cplx Q = lds_volatile(rd + o_own + (role ? sb : -sb));  // own array, own side
            const cplx Aj = role ? oth : own, Cj = role ? own : oth;
            const PcrPivot pv = pcr_pivot(Aj, Cm, Cj, Ap);
            own = -pcr_scale(own * Q, pv);
            oth = mk(__shfl_xor_sync(0xffffffffu, own.re, 1), __shfl_xor_sync(0xffffffffu, own.im, 1));
            xv = pcr_scale(nmsub(nmsub(xv, Aj, Xm), Cj, Xp), pv);
            sts_volatile(wr + o_val, xv);
            sts_volatile(wr + o_own, own);  // (not needed after the last level; kept for uniformity)
            cur ^= 1;
            __syncthreads();
        }
        FH_TR(8);
        // now x(row 2j+1) = D_j - V_j * x_partner, with D, V in half[cur]
        const cplx *HD = sm.half[cur][2] + kHalfPad, *HV = sm.half[cur][3] + kHalfPad;
        // x_partner: the unknown the V column couples to -- the partner's cut-side unknown, or the cut row's
        cplx x_partner = czero();
        if (CLUSTER > 1) {
            if (j == kH - 1) {  // two threads: role 0 sends D, role 1 sends V
                const uint32_t rbar = map_to_rank(&sm.xchg_bar[parity], rank ^ 1);
                st_async_remote(map_to_rank(&sm.xchg[parity][role], rank ^ 1), xv, rbar);
            }
            mbar_wait(&sm.xchg_bar[parity], (it >> 1) & 1);
            const cplx Yp = sm.xchg[parity][0], Vp = sm.xchg[parity][1];
            const cplx Ym = HD[kH - 1], Vm = HV[kH - 1];
            if (has_cut) {
                // x_0 = Y0 - V0 x_cut, x_1 = Y1 - V1 x_cut (subscript = cluster rank); evaluated in the same order
                // on both CTAs, so both obtain the same bits
                const cplx Y0 = rank == 0 ? Ym : Yp, V0 = rank == 0 ? Vm : Vp;
                const cplx Y1 = rank == 0 ? Yp : Ym, V1 = rank == 0 ? Vp : Vm;
                const cplx den = nmsub(nmsub(cb, cl, V0), ch, V1);
                x_partner = nmsub(nmsub(cd, cl, Y0), ch, Y1) * frecip(den);
                if (rank == 0 && t == kT - 1) stg_stream(u + sys * n + cut_row, x_partner);
            } else {
                // x_me = Ym - Vm x_p ; x_p = Yp - Vp x_me
                const cplx x_me = nmsub(Ym, Vm, Yp) * frecip(nmsub(cone(), Vm, Vp));
                x_partner = nmsub(Yp, Vp, x_me);
            }
        }
        FH_TR(9);
        cplx X, xl;
        {
            const int jo = t >> 1;  // odd t: own row; even t: the odd row above (t + 1)
            const cplx Xhi = nmsub(HD[jo], HV[jo], x_partner);
            if (t & 1) {
                X = Xhi;
            } else {
                xl = nmsub(HD[jo - 1], HV[jo - 1], x_partner);  // (t == 0 reads the zero row: xl = 0)
                X = nmsub(nmsub(rD, rA, xl), rC, Xhi);
            }
            const double ure = __shfl_up_sync(0xffffffffu, X.re, 1), uim = __shfl_up_sync(0xffffffffu, X.im, 1);
            if (t & 1) xl = mk(ure, uim);
        }

        // ---- back-substitution + staging -----------------------------------------------------------
        // Non-finite values: a vanishing pivot anywhere in this thread's chunk turns into a NaN that travels down the
        // elimination chain into the thread's own interface row (0 * Inf and Inf * 0 are NaN, NaN never cancels), so
        // testing X -- not all eight unknowns -- is enough; the interior rows only multiply finite numbers.
        bool bad = !cfinite(X) || !cfinite(x_partner);
        cplx xr[kM];
#pragma unroll
        for (int i = 0; i < kM - 1; ++i) xr[i] = nmsub(nmsub(d[i], a[i], xl), c[i], X) * b[i];
        xr[kM - 1] = X;
        const cplx x0 = xr[0], x6 = xr[kM - 2];
        if (active) {
            const int tt = map.dir > 0 ? t - map.first_active : kT - 1 - t;  // chunk in memory order, left-aligned
            const int r = tt / kPer8;
            const uint32_t base = smem_u32(&sm.xs[0]) + (uint32_t)r * 128u, x = (uint32_t)(r & 7);
#pragma unroll
            for (int i = 0; i < kM; ++i) {
                const uint32_t e = (uint32_t)((map.dir > 0 ? i : kM - 1 - i) + kM * (tt % kPer8));
                sts_volatile(base + ((e ^ x) << 4), xr[i]);
            }
        }
        if (map.head > 0 && t == feeder) {
            cplx xn = x0;  // first unknown after the head rows (unused when the system has no chunk)
            cplx *out = u + sys * n;
#pragma unroll
            for (int j = kHeadMax - 1; j >= 0; --j) {
                if (j < map.head) {
                    xn = (map.chunks == 0 && j == map.head - 1) ? hdp[j] : nmsub(hdp[j], hcp[j], xn);
                    bad |= !cfinite(xn);
                    stg_stream(out + map.global_row(j), xn);
                }
            }
        }
        FH_TR(10);
        fence_proxy_async();  // make this thread's xs writes visible to the TMA engine
        if constexpr (Source::kStoreBands) {  // the band tiles are refilled right after this barrier
            if (t >= kBandStoreThread0 && t < kBandStoreThread0 + 128 && (t & 31) == 0) tma_store_wait_read();
        }
        const int any_nonfinite = __syncthreads_or(bad ? 1 : 0);  // (returns a truth value, not a bit-wise OR)
        FH_TR(11);
        if (t == kStoreThread) {
            if (map.chunks > 0) {  // (a system shorter than one chunk consists of head rows only)
                tma_store_3d(&umaps.r[rank], 0, 0, (int)sys, sm.xs);
                tma_store_commit();
            }
            // status[] is zeroed by the host wrapper before the launch; both cluster ranks may flag it
            if (status != nullptr && any_nonfinite) atomicOr(&status[sys], 1);
        }
        {   // probe: defect of the original interface row with the computed solution
            cplx xn = x_partner;  // thread 255: the partner's interface unknown (or nothing: c7o == 0)
            if (active && t < kT - 1) xn = sm.xs[map.xs_slot(t + 1, 0)];
            const cplx r = nmsub(nmsub(nmsub(d7o, a7o, x6), b7o, X), c7o, xn);
            const double scale = abs1(a7o) * abs1(x6) + abs1(b7o) * abs1(X) + abs1(c7o) * abs1(xn) + abs1(d7o);
            const bool off = active && abs1(r) > kProbeTol * scale;
            // a defect that one refinement step cannot be trusted to repair counts as a breakdown: bit 0 sends the
            // system to the pivoted re-solve, exactly like a non-finite value
            const bool severe = active && abs1(r) > kSevereTol * scale;
            // rare: one global atomic per affected warp (status[] was zeroed by the host wrapper)
            const bool any_off = __any_sync(0xffffffffu, off), any_severe = __any_sync(0xffffffffu, severe);
            if (any_off && (t & 31) == 0 && status != nullptr) atomicOr(&status[sys], any_severe ? 3 : 2);
        }
        FH_TR(12);
    }
    if (t == kStoreThread) tma_store_wait_read();
    if (CLUSTER > 1) cg::this_cluster().sync();  // no CTA leaves while its partner may still address it
}

}  // namespace fh

#include "fh_tridiag_pipe.cuh"

namespace fh {

// ----------------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------------
constexpr int64_t kBatchedMaxN = 2 * (int64_t)kStageRows;  // 4112

static int ceil_log2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

static int pcr_levels(int64_t n, int cluster) {
    const HostMap h = host_map(n, cluster);
    const int chunks = h.chunks[0] > h.chunks[1] ? h.chunks[0] : h.chunks[1];
    const int rows = chunks < 1 ? 1 : kPer8 * chunks;  // active interface rows = thread chunks (the last `rows` threads)
    // level 0 is the cyclic-reduction step; PCR then runs over the odd rows among the active ones
    const int lv = ceil_log2((rows + 1) / 2);
    return lv > kLevelsMax - 1 ? kLevelsMax - 1 : lv;
}

template <int CLUSTER, class Source>
static int launch_batched(const Source &src, const cplx *const bands[4], cplx *u, int64_t n, int64_t batch,
                          int32_t *status, cudaStream_t stream, int64_t batch_matrices = -1) {
    // batch = number of right-hand sides / solutions (an upper bound when src.count is set); batch_matrices = number
    // of systems in the dl / d / du arrays when it differs (subset refinement)
    if (batch_matrices < 0) batch_matrices = batch;
    // The pipeline of warp roles (fh_tridiag_pipe.cuh) serves stored bands and rows generated on a uniform mesh; rows
    // generated on a non-uniform mesh (per-row geometry with IEEE divisions, far more registers) stay with the
    // one-role-per-CTA kernel.  Development switch: FH_BATCHED=serial launches that kernel for every source.
    constexpr bool kPipe = Source::kStaged || Source::kUniform;
    static const bool serial = [] {
        const char *e = getenv("FH_BATCHED");
        return !kPipe || (!Source::kRefine && e != nullptr && e[0] == 's');
    }();
    auto pipe_or_serial = [] {
        if constexpr (kPipe) return tridiag_pipe_kernel<CLUSTER, Source>;
        else return tridiag_batched_kernel<CLUSTER, Source>;
    };
    auto serial_or_pipe = [] {  // (the one-pass refinement exists as a pipeline instance only)
        if constexpr (Source::kRefine) return tridiag_pipe_kernel<CLUSTER, Source>;
        else return tridiag_batched_kernel<CLUSTER, Source>;
    };
    auto kernel = serial ? serial_or_pipe() : pipe_or_serial();
    const size_t smem = (serial ? sizeof(BatchedSmem) : sizeof(PipeSmem)) + 1024;  // slack for the 1024-byte round-up
    if (smem > smem_optin_bytes()) {
        set_error("kernel needs %zu B of shared memory, device offers %zu", smem, smem_optin_bytes());
        return FH_E_UNSUPPORTED;
    }
    if (batch >= (1ll << 31) || batch_matrices >= (1ll << 31)) {
        set_error("batch too large for one launch");
        return FH_E_UNSUPPORTED;
    }
    const HostMap h = host_map(n, CLUSTER);
    TensorMaps maps;
    StoreMaps umaps;
    memset(&maps, 0, sizeof(maps));
    memset(&umaps, 0, sizeof(umaps));
    for (int r = 0; r < CLUSTER; ++r) {  // every rank sees only its own chunks: loads zero-fill, stores clip
        // (refinement in place: the solutions are those of the whole batch, addressed through the list)
        int rc = make_tensor_map(&umaps.r[r], u, n, Source::kRefine ? batch_matrices : batch, h.row0[r], h.chunks[r],
                                 kChunks8);
        if (rc != FH_OK) return rc;
        if (Source::kStoreBands) {  // store views over the interior chunks of the rank, in boxes of kBandTileChunks chunks
            for (int i = 0; i < 4; ++i) {
                rc = make_tensor_map(&maps.band[r][i], bands[i], n, batch, h.row0[r] + 8 * h.front[r],
                                     h.chunks[r] - h.front[r] - h.back[r], kBandTileChunks);
                if (rc != FH_OK) return rc;
            }
        }
        if (Source::kStaged) {
            for (int i = 0; i < 4; ++i) {
                rc = make_tensor_map(&maps.band[r][i], bands[i], n, i < 3 || Source::kRefine ? batch_matrices : batch,
                                     h.row0[r], h.chunks[r], kChunks8);
                if (rc != FH_OK) return rc;
            }
        }
    }
    FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (status != nullptr && !Source::kRefine) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * batch, stream));
    int64_t groups = sm_count() / CLUSTER;  // one CTA per SM (~194 KB of shared memory each)
    if (groups < 1) groups = 1;
    if (groups > batch) groups = batch;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * CLUSTER));
    cfg.blockDim = dim3(serial ? kT : pipe_threads<Source>());
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int levels = pcr_levels(n, CLUSTER);
    FH_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, umaps, src, u, n, batch, levels, status));
    return FH_OK;
}

// from fh_tridiag_large.cu
int solve_large(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n,
                int32_t *status, void *workspace, size_t workspace_bytes, cudaStream_t stream);
size_t solve_large_workspace_bytes(int64_t n);

int batched_solve_bands(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n,
                        int64_t batch, int32_t *status, cudaStream_t stream) {
    if (n > kBatchedMaxN) {
        set_error("batched kernel handles n <= %lld", (long long)kBatchedMaxN);
        return FH_E_UNSUPPORTED;
    }
    BandSource src;
    src.dl = dl, src.d = d, src.du = du, src.b = b;
    const cplx *const bands[4] = {dl, d, du, b};
    if (n <= kStageRows) return launch_batched<1>(src, bands, u, n, batch, status, stream);
    return launch_batched<2>(src, bands, u, n, batch, status, stream);
}

// One refinement step  u(list[i]) += A(list[i])^-1 (b - A u)(list[i]),  i < min(*count, cap), in one pass over the
// flagged systems (BandRefineSource): status[list[i]] loses the small-pivot bit and gains the breakdown bit if the
// correction solve breaks down.
int batched_refine_subset(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n, int64_t batch,
                          const int32_t *list, const int32_t *count, int64_t cap, int32_t *status, cudaStream_t stream) {
    BandRefineSource src;
    src.dl = dl, src.d = d, src.du = du, src.b = b;
    src.list = list, src.count = count, src.u_old = u;
    const cplx *const bands[4] = {dl, d, du, b};
    if (n <= kStageRows) return launch_batched<1>(src, bands, u, n, cap, status, stream, batch);
    return launch_batched<2>(src, bands, u, n, cap, status, stream, batch);
}

}  // namespace fh

extern "C" {

#ifdef FH_TRACE
__attribute__((visibility("default"))) int fh_debug_set_trace(long long *buf, int iter) {
    cudaMemcpyToSymbol(fh::g_trace, &buf, sizeof(buf));
    cudaMemcpyToSymbol(fh::g_trace_iter, &iter, sizeof(iter));
    return (int)cudaGetLastError();
}
#endif

int64_t fh_tridiag_batched_max_n(void) { return fh::kBatchedMaxN; }

size_t fh_tridiag_solve_batched_workspace_bytes(int64_t n, int64_t batch) {
    (void)batch;
    return n > fh::kBatchedMaxN ? fh::solve_large_workspace_bytes(n) : 0;
}

int fh_tridiag_solve_batched_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                  fh_c128 *u, int64_t n, int64_t batch, int32_t *status, void *workspace,
                                  size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0 && batch >= 0, "n and batch must be >= 0");
    if (n == 0 || batch == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u, "NULL device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n > kBatchedMaxN) {  // systems too long for a cluster: partitioned solver, one system at a time
        for (int64_t s = 0; s < batch; ++s) {
            int rc = solve_large(reinterpret_cast<const cplx *>(dl) + s * n, reinterpret_cast<const cplx *>(d) + s * n,
                                 reinterpret_cast<const cplx *>(du) + s * n, reinterpret_cast<const cplx *>(b) + s * n,
                                 reinterpret_cast<cplx *>(u) + s * n, n, status ? status + s : nullptr, workspace,
                                 workspace_bytes, st);
            if (rc != FH_OK) return rc;
        }
        return FH_OK;
    }
    return batched_solve_bands(reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d),
                               reinterpret_cast<const cplx *>(du), reinterpret_cast<const cplx *>(b),
                               reinterpret_cast<cplx *>(u), n, batch, status, st);
}

int fh_helmholtz1d_sweep_fused_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s,
                                    int64_t n_k, fh_c128 *u, int32_t *status, void *workspace,
                                    size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    (void)workspace, (void)workspace_bytes;
    Problem1d p;
    int rc = problem_from_host(prob_host, p);
    if (rc != FH_OK) return rc;
    FH_REQUIRE(n_k >= 0, "n_k must be >= 0");
    if (n_k == 0) return FH_OK;
    FH_REQUIRE(ks && k2s && u, "NULL device pointer");
    const int64_t n = p.n_elem + 1;
    if (n > kBatchedMaxN) {
        set_error("fused sweep supports meshes up to %lld nodes; use fh_helmholtz1d_solve_large_fused_c128 beyond",
                  (long long)kBatchedMaxN);
        return FH_E_UNSUPPORTED;
    }
    const cplx *const none[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cplx *out = reinterpret_cast<cplx *>(u);
    if (p.h_mode == FH_H_SCALAR && p.n_elem >= 2) {  // uniform mesh: the fast generator
        HelmSourceT<true> src;
        src.p = p, src.ks = ks, src.k2s = k2s;
        src.out_dl = src.out_d = src.out_du = src.out_b = nullptr;
        if (n <= kStageRows) return launch_batched<1>(src, none, out, n, n_k, status, st);
        return launch_batched<2>(src, none, out, n, n_k, status, st);
    }
    HelmSourceT<false> src;
    src.p = p, src.ks = ks, src.k2s = k2s;
    src.out_dl = src.out_d = src.out_du = src.out_b = nullptr;
    if (n <= kStageRows) return launch_batched<1>(src, none, out, n, n_k, status, st);
    return launch_batched<2>(src, none, out, n, n_k, status, st);
}

int fh_helmholtz1d_sweep_assemble_solve_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s,
                                             int64_t n_k, fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b, fh_c128 *u,
                                             int32_t *status, void *workspace, size_t workspace_bytes,
                                             fh_stream_t stream) {
    using namespace fh;
    (void)workspace, (void)workspace_bytes;
    Problem1d p;
    int rc = problem_from_host(prob_host, p);
    if (rc != FH_OK) return rc;
    FH_REQUIRE(n_k >= 0, "n_k must be >= 0");
    if (n_k == 0) return FH_OK;
    FH_REQUIRE(ks && k2s && dl && d && du && b && u, "NULL device pointer");
    const int64_t n = p.n_elem + 1;
    if (n > kBatchedMaxN || p.h_mode != FH_H_SCALAR || p.n_elem < 2) {
        set_error("the one-pass assemble+solve covers uniform meshes of 3 .. %lld nodes; call fh_assemble_1d_c128 and "
                  "fh_tridiag_solve_batched_c128 otherwise", (long long)kBatchedMaxN);
        return FH_E_UNSUPPORTED;
    }
    HelmSourceT<true, true> src;
    src.p = p, src.ks = ks, src.k2s = k2s;
    src.out_dl = reinterpret_cast<cplx *>(dl), src.out_d = reinterpret_cast<cplx *>(d);
    src.out_du = reinterpret_cast<cplx *>(du), src.out_b = reinterpret_cast<cplx *>(b);
    const cplx *const bands[4] = {src.out_dl, src.out_d, src.out_du, src.out_b};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cplx *out = reinterpret_cast<cplx *>(u);
    if (n <= kStageRows) return launch_batched<1>(src, bands, out, n, n_k, status, st);
    return launch_batched<2>(src, bands, out, n, n_k, status, st);
}

}  // extern "C"
