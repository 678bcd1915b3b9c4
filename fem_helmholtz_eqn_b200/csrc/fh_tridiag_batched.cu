// Batched complex128 tridiagonal solve for sm_100a (replaces nb:204, np.linalg.solve(A, b), for
// the tridiagonal A of nb:134-187) and the fused assemble+solve sweep (nb:190-205 for many k).
//
// Algorithm (per system)
//   1. thread-chunk partition: every thread owns kM = 8 consecutive rows in REGISTERS and removes
//      its interior couplings with a downward and an upward elimination sweep, leaving
//        a_i x_left + b_i x_i + c_i x_last = d_i   (i < 7)   and one interface row per thread;
//   2. the 256 interface rows of a CTA are solved by parallel cyclic reduction (PCR) in shared
//      memory, 8 levels, rows kept normalised (unit diagonal);
//   3. back-substitution in registers, results staged through shared memory so that the global
//      store is fully coalesced.
//   A system of up to 2056 rows is handled by one CTA; up to 4112 rows by a 2-CTA cluster: rank 0
//   takes the first half, rank 1 the second half in REVERSED row order, so both halves end at the
//   cut.  Each half carries one extra PCR right-hand side (its coupling to the partner's interface
//   unknown); the two 2x2-coupled interface values are exchanged once through distributed shared
//   memory (one cluster barrier per system).
//
// Data movement (staged variant): the CTA is persistent.  While system s is being processed from
// registers, the bands of system s + gridDim are already in flight into shared memory via 16-byte
// cp.async (LDGSTS) -- consecutive threads fetch consecutive 16-byte entries (coalesced), the
// shared-memory slot index is XOR-swizzled so that the later per-thread chunk reads (stride 128 B)
// are bank-conflict free.  HBM traffic = algorithmic traffic: 64 B/row read + 16 B/row written.
//
// Fused variant: the same kernel with the band loads replaced by the row generator of
// fh_helm1d.cuh (bit-identical to what fh_assemble_1d_c128 would have stored); 16 B/row written.
#include <cooperative_groups.h>

#include "fh_helm1d.cuh"

namespace cg = cooperative_groups;

namespace fh {

constexpr int kT = 256;                        // threads per CTA
constexpr int kM = 8;                          // rows per thread (register chunk)
constexpr int kSlots = kT * kM;                // 2048 body rows per CTA
constexpr int kHeadMax = 8;                    // extra rows at the far end, eliminated serially
constexpr int kStageRows = kSlots + kHeadMax;  // 2056
constexpr int kLevelsMax = 8;                  // log2(kT)

// chunk p = q >> 3 keeps its 8 entries inside one 128-byte line, permuted by p & 7
__device__ __forceinline__ int swz(int q) { return q ^ ((q >> 3) & 7); }

struct BatchedSmem {
    cplx stage[4][kStageRows];  // prefetched bands of the NEXT system (a, b, c, d)
    cplx xs[kStageRows];        // solution staging of the CURRENT system
    cplx pcr[2][4][kT];         // PCR ping-pong: A, C, D, V  (pcr[1] doubles as the row-0 exchange)
    cplx head[2][kHeadMax];     // normalised c', d' of the head rows
    cplx xchg[2][2];            // [parity][{Y, V}] written by the partner CTA (DSMEM)
};

// How the rows of one system map onto the slots of one CTA.
struct LocalMap {
    int n_loc;    // rows handled by this CTA
    int head;     // rows beyond the 2048 body slots (<= kHeadMax), at the far end
    int pad;      // identity slots in front of the first body row
    int64_t g0;   // global row of local row 0
    int dir;      // +1 forward, -1 reversed
    __device__ __forceinline__ int64_t global_row(int j) const { return g0 + (int64_t)dir * j; }
    __device__ __forceinline__ int slot_of(int j) const {  // stage / xs index of local row j
        return j < head ? kSlots + j : swz(pad + j - head);
    }
};

template <int CLUSTER>
__device__ __forceinline__ LocalMap make_map(int64_t n, int rank) {
    LocalMap m;
    if (CLUSTER == 1) {
        m.n_loc = (int)n, m.g0 = 0, m.dir = 1;
    } else {
        const int n0 = (int)(n / 2);
        if (rank == 0) m.n_loc = n0, m.g0 = 0, m.dir = 1;
        else m.n_loc = (int)n - n0, m.g0 = n - 1, m.dir = -1;
    }
    m.head = max(0, m.n_loc - kSlots);
    m.pad = kSlots - (m.n_loc - m.head);
    return m;
}

// ----------------------------------------------------------------------------------------------
// Band sources
// ----------------------------------------------------------------------------------------------
struct BandSource {  // staged pipeline: bands live in global memory
    const cplx *dl, *d, *du, *b;
    static constexpr bool kStaged = true;

    __device__ __forceinline__ void prefetch(BatchedSmem &sm, const LocalMap &m, int64_t sys, int64_t n) const {
        const cplx *lo = m.dir > 0 ? dl : du;  // reversed order swaps sub- and super-diagonal
        const cplx *hi = m.dir > 0 ? du : dl;
        const int64_t base = sys * n;
        for (int j = threadIdx.x; j < m.n_loc; j += kT) {
            const int64_t g = base + m.global_row(j);
            const int s = m.slot_of(j);
            cp_async16(&sm.stage[0][s], lo + g);
            cp_async16(&sm.stage[1][s], d + g);
            cp_async16(&sm.stage[2][s], hi + g);
            cp_async16(&sm.stage[3][s], b + g);
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void begin_system(int64_t) {}
    __device__ __forceinline__ void row(const BatchedSmem &sm, const LocalMap &m, int j, cplx &a, cplx &b_,
                                        cplx &c, cplx &d_) const {
        const int s = m.slot_of(j);
        a = sm.stage[0][s], b_ = sm.stage[1][s], c = sm.stage[2][s], d_ = sm.stage[3][s];
    }
};

struct HelmSource {  // fused pipeline: bands are generated, never stored
    Problem1d p;
    const double *ks, *k2s;
    double k, k2;
    RowGeom g_first, g_mid, g_last;  // scalar-h mode: the only three distinct geometries
    static constexpr bool kStaged = false;

    __device__ __forceinline__ void init() {
        if (p.h_mode == FH_H_SCALAR) {
            g_first = row_geom(p, 0);
            g_mid = row_geom(p, p.n_elem > 1 ? 1 : 0);
            g_last = row_geom(p, p.n_elem);
        }
    }
    __device__ __forceinline__ void prefetch(BatchedSmem &, const LocalMap &, int64_t, int64_t) const {}
    __device__ __forceinline__ void begin_system(int64_t sys) { k = __ldg(ks + sys), k2 = __ldg(k2s + sys); }
    __device__ __forceinline__ void row(const BatchedSmem &, const LocalMap &m, int j, cplx &a, cplx &b_,
                                        cplx &c, cplx &d_) const {
        const int64_t g = m.global_row(j);
        cplx lo, hi;
        if (p.h_mode == FH_H_SCALAR) {
            const RowGeom &rg = g == 0 ? g_first : (g == p.n_elem ? g_last : g_mid);
            row_values(p, rg, g, k, k2, lo, b_, hi, d_);
        } else {
            row_values(p, row_geom(p, g), g, k, k2, lo, b_, hi, d_);
        }
        a = m.dir > 0 ? lo : hi;
        c = m.dir > 0 ? hi : lo;
    }
};

// ----------------------------------------------------------------------------------------------
// The kernel
// ----------------------------------------------------------------------------------------------
template <int CLUSTER, class Source>
__global__ void __launch_bounds__(kT, 1)
tridiag_batched_kernel(Source src, cplx *__restrict__ u, int64_t n, int64_t batch, int levels,
                       int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BatchedSmem &sm = *reinterpret_cast<BatchedSmem *>(smem_raw);

    const int t = threadIdx.x;
    int rank = 0;
    int64_t first_sys = blockIdx.x, sys_stride = gridDim.x;
    if (CLUSTER > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        rank = (int)cluster.block_rank();
        first_sys = blockIdx.x / CLUSTER;
        sys_stride = gridDim.x / CLUSTER;
    }
    const LocalMap map = make_map<CLUSTER>(n, rank);
    const int first_active = map.pad / kM;  // threads below this own identity rows only

    if constexpr (!Source::kStaged) src.init();
    if (first_sys < batch) src.prefetch(sm, map, first_sys, n);

    int parity = 0;
    for (int64_t sys = first_sys; sys < batch; sys += sys_stride, parity ^= 1) {
        src.begin_system(sys);
        if constexpr (Source::kStaged) {
            cp_async_wait_all();
            __syncthreads();
        }

        // ---- bands -> registers ------------------------------------------------------------
        cplx a[kM], b[kM], c[kM], d[kM];
#pragma unroll
        for (int i = 0; i < kM; ++i) {
            const int q = t * kM + i;
            if (q >= map.pad) {
                src.row(sm, map, map.head + q - map.pad, a[i], b[i], c[i], d[i]);
            } else {
                a[i] = czero(), b[i] = cone(), c[i] = czero(), d[i] = czero();
            }
        }
        if (map.head == 0 && t == first_active) {  // local row 0: nothing above it
            const int i0 = map.pad - first_active * kM;
#pragma unroll
            for (int i = 0; i < kM; ++i)
                if (i == i0) a[i] = czero();
        }
        if (CLUSTER == 1 && t == kT - 1) c[kM - 1] = czero();  // last row of the system

        // ---- head rows (far end, at most 8): serial Thomas by thread 0, folded into its row 0 --
        if (map.head > 0 && t == 0) {
            cplx cp = czero(), dp = czero();
            for (int j = 0; j < map.head; ++j) {
                cplx ha, hb, hc, hd;
                src.row(sm, map, j, ha, hb, hc, hd);
                if (j == 0) ha = czero();
                const cplx inv = crecip(nmsub(hb, ha, cp));
                dp = nmsub(hd, ha, dp) * inv;
                cp = hc * inv;
                sm.head[0][j] = cp, sm.head[1][j] = dp;
            }
            b[0] = nmsub(b[0], a[0], cp);
            d[0] = nmsub(d[0], a[0], dp);
            a[0] = czero();
        }
        if constexpr (Source::kStaged) {
            __syncthreads();  // everyone has drained the staging buffers ...
            if (sys + sys_stride < batch) src.prefetch(sm, map, sys + sys_stride, n);  // ... refill them
        }

        // ---- phase 1: downward elimination inside the chunk ------------------------------------
#pragma unroll
        for (int i = 1; i < kM; ++i) {
            const cplx w = a[i] * crecip(b[i - 1]);
            b[i] = nmsub(b[i], w, c[i - 1]);
            d[i] = nmsub(d[i], w, d[i - 1]);
            a[i] = -(w * a[i - 1]);
        }
        // ---- phase 2: upward elimination; afterwards row i < 7 couples to x_left and x_7 only ---
#pragma unroll
        for (int i = kM - 3; i >= 0; --i) {
            const cplx w = c[i] * crecip(b[i + 1]);
            a[i] = nmsub(a[i], w, a[i + 1]);
            d[i] = nmsub(d[i], w, d[i + 1]);
            c[i] = -(w * c[i + 1]);
        }
        // ---- interface row: eliminate the next chunk's first unknown with its (normalised) row 0 -
        {
            const cplx inv0 = crecip(b[0]);
            sm.pcr[1][0][t] = a[0] * inv0;
            sm.pcr[1][1][t] = c[0] * inv0;
            sm.pcr[1][2][t] = d[0] * inv0;
        }
        __syncthreads();
        cplx rA, rC, rD, rV;
        {
            cplx rb = b[kM - 1];
            rA = a[kM - 1], rC = czero(), rD = d[kM - 1], rV = czero();
            if (t < kT - 1) {
                const cplx nA = sm.pcr[1][0][t + 1], nC = sm.pcr[1][1][t + 1], nD = sm.pcr[1][2][t + 1];
                rb = nmsub(rb, c[kM - 1], nA);
                rC = -(c[kM - 1] * nC);
                rD = nmsub(rD, c[kM - 1], nD);
            } else if (CLUSTER > 1) {
                rV = c[kM - 1];  // coupling to the partner CTA's interface unknown
            }
            const cplx inv = crecip(rb);
            rA = rA * inv, rC = rC * inv, rD = rD * inv, rV = rV * inv;
        }
        // ---- PCR over the CTA's 256 interface rows (unit diagonal) ---------------------------------
        int cur = 0;
        sm.pcr[0][0][t] = rA, sm.pcr[0][1][t] = rC, sm.pcr[0][2][t] = rD;
        if (CLUSTER > 1) sm.pcr[0][3][t] = rV;
        __syncthreads();
        for (int lv = 0, s = 1; lv < levels; ++lv, s <<= 1) {
            cplx Am = czero(), Cm = czero(), Dm = czero(), Vm = czero();
            cplx Ap = czero(), Cp = czero(), Dp = czero(), Vp = czero();
            if (t - s >= 0) {
                Am = sm.pcr[cur][0][t - s], Cm = sm.pcr[cur][1][t - s], Dm = sm.pcr[cur][2][t - s];
                if (CLUSTER > 1) Vm = sm.pcr[cur][3][t - s];
            }
            if (t + s < kT) {
                Ap = sm.pcr[cur][0][t + s], Cp = sm.pcr[cur][1][t + s], Dp = sm.pcr[cur][2][t + s];
                if (CLUSTER > 1) Vp = sm.pcr[cur][3][t + s];
            }
            const cplx inv = crecip(nmsub(nmsub(cone(), rA, Cm), rC, Ap));
            const cplx nD = nmsub(nmsub(rD, rA, Dm), rC, Dp) * inv;
            cplx nV = czero();
            if (CLUSTER > 1) nV = nmsub(nmsub(rV, rA, Vm), rC, Vp) * inv;
            const cplx nA = -(rA * Am) * inv;
            const cplx nC = -(rC * Cp) * inv;
            rA = nA, rC = nC, rD = nD, rV = nV;
            cur ^= 1;
            sm.pcr[cur][0][t] = rA, sm.pcr[cur][1][t] = rC, sm.pcr[cur][2][t] = rD;
            if (CLUSTER > 1) sm.pcr[cur][3][t] = rV;
            __syncthreads();
        }
        // now x_last(t) = rD - rV * x_partner
        cplx X = rD;
        if (CLUSTER > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (t == kT - 1) {
                cplx *remote = cluster.map_shared_rank(&sm.xchg[parity][0], rank ^ 1);
                remote[0] = rD, remote[1] = rV;
            }
            cluster.sync();
            const cplx Yp = sm.xchg[parity][0], Vp = sm.xchg[parity][1];
            const cplx Ym = sm.pcr[cur][2][kT - 1], Vm = sm.pcr[cur][3][kT - 1];
            // x_me = Ym - Vm x_p ; x_p = Yp - Vp x_me
            const cplx x_me = nmsub(Ym, Vm, Yp) * crecip(nmsub(cone(), Vm, Vp));
            const cplx x_p = nmsub(Yp, Vp, x_me);
            X = nmsub(rD, rV, x_p);
        }
        sm.pcr[cur ^ 1][0][t] = X;
        __syncthreads();
        const cplx xl = t > 0 ? sm.pcr[cur ^ 1][0][t - 1] : czero();

        // ---- back-substitution + staging --------------------------------------------------------
        bool bad = !cfinite(X);
        cplx x0 = X;
#pragma unroll
        for (int i = 0; i < kM; ++i) {
            cplx xi = X;
            if (i < kM - 1) xi = nmsub(nmsub(d[i], a[i], xl), c[i], X) * crecip(b[i]);
            if (i == 0) x0 = xi;
            bad |= !cfinite(xi);
            sm.xs[swz(t * kM + i)] = xi;
        }
        if (map.head > 0 && t == 0) {
            cplx xn = x0;
            for (int j = map.head - 1; j >= 0; --j) {
                xn = nmsub(sm.head[1][j], sm.head[0][j], xn);
                bad |= !cfinite(xn);
                sm.xs[kSlots + j] = xn;
            }
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        {
            cplx *out = u + sys * n;
            for (int j = t; j < map.n_loc; j += kT) stg_stream(out + map.global_row(j), sm.xs[map.slot_of(j)]);
        }
        // status[] is zeroed by the host wrapper before the launch; both cluster ranks may flag it
        if (status != nullptr && t == 0 && any_bad) atomicOr(&status[sys], 1);
        // xs and pcr buffers are rewritten only after the next iteration's barriers
    }
    if constexpr (Source::kStaged) cp_async_wait_all();
}


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
    int64_t n_loc = cluster == 1 ? n : n - n / 2;  // the larger half
    if (n_loc > kSlots) n_loc = kSlots;
    const int chunks = (int)((n_loc + kM - 1) / kM);
    const int lv = ceil_log2(chunks);
    return lv > kLevelsMax ? kLevelsMax : lv;
}

template <int CLUSTER, class Source>
static int launch_batched(const Source &src, cplx *u, int64_t n, int64_t batch, int32_t *status,
                          cudaStream_t stream) {
    auto kernel = tridiag_batched_kernel<CLUSTER, Source>;
    const size_t smem = sizeof(BatchedSmem);
    if (smem > smem_optin_bytes()) {
        set_error("kernel needs %zu B of shared memory, device offers %zu", smem, smem_optin_bytes());
        return FH_E_UNSUPPORTED;
    }
    FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (status != nullptr) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * batch, stream));
    int64_t groups = sm_count() / CLUSTER;  // one CTA per SM (197 KB of shared memory each)
    if (groups < 1) groups = 1;
    if (groups > batch) groups = batch;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * CLUSTER));
    cfg.blockDim = dim3(kT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int levels = pcr_levels(n, CLUSTER);
    FH_CUDA(cudaLaunchKernelEx(&cfg, kernel, src, u, n, batch, levels, status));
    return FH_OK;
}

}  // namespace fh

namespace fh {
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
    if (n <= kStageRows) return launch_batched<1>(src, u, n, batch, status, stream);
    return launch_batched<2>(src, u, n, batch, status, stream);
}
}  // namespace fh

extern "C" {

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
    HelmSource src;
    int rc = problem_from_host(prob_host, src.p);
    if (rc != FH_OK) return rc;
    FH_REQUIRE(n_k >= 0, "n_k must be >= 0");
    if (n_k == 0) return FH_OK;
    FH_REQUIRE(ks && k2s && u, "NULL device pointer");
    const int64_t n = src.p.n_elem + 1;
    if (n > kBatchedMaxN) {
        set_error("fused sweep supports meshes up to %lld nodes; assemble + fh_tridiag_solve_large_c128 beyond",
                  (long long)kBatchedMaxN);
        return FH_E_UNSUPPORTED;
    }
    src.ks = ks, src.k2s = k2s, src.k = 0.0, src.k2 = 0.0;
    src.g_first = src.g_mid = src.g_last = RowGeom{0, 0, 0, 0, 0, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= kStageRows) return launch_batched<1>(src, reinterpret_cast<cplx *>(u), n, n_k, status, st);
    return launch_batched<2>(src, reinterpret_cast<cplx *>(u), n, n_k, status, st);
}

}  // extern "C"
