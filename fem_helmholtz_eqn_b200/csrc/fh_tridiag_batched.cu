// Batched complex128 tridiagonal solve for sm_100a (replaces nb:204, np.linalg.solve(A, b), for
// the tridiagonal A of nb:134-187) and the fused assemble+solve sweep (nb:190-205 for many k).
//
// Algorithm (per system)
//   1. thread-chunk partition: every thread owns kM = 8 consecutive rows in REGISTERS and removes
//      its interior couplings with a downward and an upward elimination sweep (pivots are kept as
//      reciprocals), leaving  a_i x_left + b_i x_i + c_i x_last = d_i  (i < 7) and one interface
//      row per thread;
//   2. the 256 interface rows of a CTA are solved by parallel cyclic reduction (PCR) in shared
//      memory, <= 8 levels, rows kept normalised (unit diagonal);
//   3. back-substitution in registers; the solution is staged in shared memory and leaves with one
//      TMA tensor store.
//   A system of up to 2056 rows is handled by one CTA; up to 4112 rows by a 2-CTA cluster: rank 0
//   takes the first half, rank 1 the second half in REVERSED row order, so both halves end at the
//   cut.  Each half carries one extra PCR right-hand side (its coupling to the partner's interface
//   unknown); the two coupled interface values are exchanged through distributed shared memory
//   with st.async + mbarrier complete_tx (no cluster-wide barrier, no fence).
//
// Data movement (staged variant): the CTA is persistent.  While system s is processed from
// registers, the four band tiles of system s + stride are in flight: one elected thread issues four
// TMA tensor loads (cp.async.bulk.tensor.3d, 32 KB each, 128-byte swizzle) that complete on an
// mbarrier.  The tensor view of a band array is [system][chunk of 8 rows][8 complex]; the
// hardware swizzle makes the later per-thread chunk reads (stride 128 B) bank-conflict free.  Rows
// that do not fill a chunk (n mod 8, at most 8 "head" rows at the far end) travel by cp.async and
// are eliminated serially by one thread.  HBM traffic = algorithmic traffic: 64 B/row read +
// 16 B/row written (ncu: profiles/).
//
// Fused variant: the same kernel with the band loads replaced by the row generator of
// fh_helm1d.cuh (bit-identical to what fh_assemble_1d_c128 would have stored); 16 B/row written.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "fh_tridiag_dev.cuh"

namespace cg = cooperative_groups;

namespace fh {

#ifndef FH_PAIR_ROWS
#define FH_PAIR_ROWS 8
#endif
constexpr int kM = FH_PAIR_ROWS;               // rows per thread (register chunk): 8 or 4
constexpr int kT = 2048 / kM;                  // threads per CTA: 256 or 512
constexpr int kSlots = kT * kM;                // 2048 body rows per CTA
constexpr int kChunks8 = kSlots / 8;           // 8-row memory chunks (TMA tile rows) per CTA: 256
constexpr int kPer8 = 8 / kM;                  // threads sharing one memory chunk: 1 or 2
static_assert(kM == 8 || kM == 4, "register chunk must be 8 or 4 rows");
constexpr int kHeadMax = 8;                    // extra rows at the far end, eliminated serially
constexpr int kStageRows = kSlots + kHeadMax;  // 2056
constexpr int kLevelsMax = kM == 8 ? 8 : 9;    // log2(kT)
constexpr uint32_t kTileBytes = kSlots * sizeof(cplx);  // 32 KB per band tile

struct __align__(1024) BatchedSmem {
    cplx stage[4][kSlots];      // TMA destination: bands of the NEXT system (a, b, c, d), swizzle-128B
    cplx xs[kSlots];            // TMA source: solution of the CURRENT system, swizzle-128B
    cplx pcr[2][4][kT];         // PCR ping-pong: A, C, D, V  (pcr[1] doubles as the row-0 exchange)
    cplx headst[4][kHeadMax];   // head rows of the next system (cp.async by one thread)
    cplx xchg[2][2];            // [parity][{Y, V}] written by the partner CTA (st.async)
    unsigned long long full_bar;     // mbarrier: band tiles have landed
    unsigned long long xchg_bar[2];  // mbarrier: partner's interface pair has landed
};

// ---------------------------------------------------------------------------------------------- row -> slot maps
// How the rows of one system map onto one CTA.  Local rows run from the far end (j = 0) to the cut / the
// system's last row (j = n_loc - 1).  The first `head` rows are "head rows"; the other 8 * chunks rows
// fill the chunks of threads kT - chunks .. kT - 1 (right-aligned).
struct LocalMap {
    int n_loc, head, chunks;
    int first_active;  // kT - chunks
    int64_t g0;        // global row of local row 0
    int dir;           // +1 forward, -1 reversed
    int box_c1;        // chunk coordinate at which this CTA's TMA box starts (may be negative)
    __device__ __forceinline__ int64_t global_row(int j) const { return g0 + (int64_t)dir * j; }
    // swizzle-128B slot of element i of thread t's chunk in the LOAD tiles (box right-aligned: TMA loads
    // accept negative start coordinates and zero-fill)
    __device__ __forceinline__ int slot(int t, int i) const {
        const int tt = dir > 0 ? t : kT - 1 - t;           // thread chunk in memory order
        const int r = tt / kPer8;                          // 8-row memory chunk = TMA tile row
        const int e = (dir > 0 ? i : kM - 1 - i) + kM * (tt % kPer8);
        return r * 8 + (e ^ (r & 7));
    }
    // ... and in the STORE tile: TMA stores trap on negative coordinates (measured, tools/tma_probe.cu), so
    // the forward half is left-aligned (box starts at chunk 0, the tensor map clips at the half's last chunk)
    __device__ __forceinline__ int xs_slot(int t, int i) const {
        const int tt = dir > 0 ? t - first_active : kT - 1 - t;
        const int r = tt / kPer8;
        const int e = (dir > 0 ? i : kM - 1 - i) + kM * (tt % kPer8);
        return r * 8 + (e ^ (r & 7));
    }
};

struct HostMap {  // the same numbers, computed on the host for the tensor maps
    int n_loc[2], head[2], chunks[2];
};
__host__ __device__ inline HostMap host_map(int64_t n, int cluster) {
    HostMap m;
    m.n_loc[0] = cluster == 1 ? (int)n : (int)(n / 2);
    m.n_loc[1] = cluster == 1 ? 0 : (int)(n - n / 2);
    for (int r = 0; r < 2; ++r) {
        int c = m.n_loc[r] / 8;          // 8-row memory chunks
        if (c > kChunks8) c = kChunks8;
        m.chunks[r] = c;
        m.head[r] = m.n_loc[r] - c * 8;  // 0..8
    }
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
        m.box_c1 = h.chunks[0];    // thread t <-> box row kT-1-t <-> memory chunk chunks0 + (kT-1-t)
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
    const CUtensorMap *map_lo = map.dir > 0 ? &maps.dl : &maps.du;
    const CUtensorMap *map_hi = map.dir > 0 ? &maps.du : &maps.dl;

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
    }

    // one elected thread feeds the pipeline: four band tiles by TMA, head rows by cp.async
    auto prefetch = [&](int64_t sys) {
        if constexpr (Source::kStaged) {
            if (t == map.first_active % kT) {  // the thread that also consumes the head rows
                mbar_expect_tx(&sm.full_bar, 4 * kTileBytes);
                tma_load_3d(sm.stage[0], map_lo, 0, map.box_c1, (int)sys, &sm.full_bar);
                tma_load_3d(sm.stage[1], &maps.d, 0, map.box_c1, (int)sys, &sm.full_bar);
                tma_load_3d(sm.stage[2], map_hi, 0, map.box_c1, (int)sys, &sm.full_bar);
                tma_load_3d(sm.stage[3], &maps.b, 0, map.box_c1, (int)sys, &sm.full_bar);
                const cplx *lo = map.dir > 0 ? src.dl : src.du, *hi = map.dir > 0 ? src.du : src.dl;
                for (int j = 0; j < map.head; ++j) {
                    const int64_t g = sys * n + map.global_row(j);
                    cp_async16(&sm.headst[0][j], lo + g);
                    cp_async16(&sm.headst[1][j], src.d + g);
                    cp_async16(&sm.headst[2][j], hi + g);
                    cp_async16(&sm.headst[3][j], src.b + g);
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
        if (CLUSTER > 1 && t == 0) mbar_expect_tx(&sm.xchg_bar[parity], 2 * sizeof(cplx));

        // ---- bands -> registers ----------------------------------------------------------------
        cplx a[kM], b[kM], c[kM], d[kM];
        if constexpr (Source::kStaged) {
            mbar_wait(&sm.full_bar, parity);
            if (active) {
#pragma unroll
                for (int i = 0; i < kM; ++i) {
                    const int s = map.slot(t, i);
                    a[i] = sm.stage[0][s], b[i] = sm.stage[1][s], c[i] = sm.stage[2][s], d[i] = sm.stage[3][s];
                }
            }
        } else {
            if (active) {
#pragma unroll
                for (int i = 0; i < kM; ++i)
                    helm_row(src, hs, map, map.head + (t - map.first_active) * kM + i, a[i], b[i], c[i], d[i]);
            }
        }
        if (!active) {
#pragma unroll
            for (int i = 0; i < kM; ++i) a[i] = czero(), b[i] = cone(), c[i] = czero(), d[i] = czero();
        }
        if (map.head == 0 && t == map.first_active) a[0] = czero();  // local row 0: nothing above it
        if (CLUSTER == 1 && t == kT - 1) c[kM - 1] = czero();        // last row of the system

        // ---- head rows (far end, at most 8): serial Thomas by one thread, folded into its row 0 ------
        cplx hcp[kHeadMax], hdp[kHeadMax];
        if (map.head > 0 && t == feeder) {
            if constexpr (Source::kStaged) cp_async_wait_all();
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
        if constexpr (Source::kStaged) {
            __syncthreads();  // every thread has drained the staging tiles ...
            if (sys + sys_stride < batch) prefetch(sys + sys_stride);  // ... refill them
        }

        // ---- phase 1: downward elimination inside the chunk; b[i] (i < 7) becomes 1 / pivot --------
        const cplx a7o = a[kM - 1], b7o = b[kM - 1], c7o = c[kM - 1], d7o = d[kM - 1];  // for the probe
        b[0] = frecip(b[0]);
#pragma unroll
        for (int i = 1; i < kM; ++i) {
            const cplx w = a[i] * b[i - 1];
            b[i] = nmsub(b[i], w, c[i - 1]);
            d[i] = nmsub(d[i], w, d[i - 1]);
            a[i] = -(w * a[i - 1]);
            if (i < kM - 1) b[i] = frecip(b[i]);
        }
        // ---- phase 2: upward elimination; afterwards row i < 7 couples to x_left and x_7 only -------
#pragma unroll
        for (int i = kM - 3; i >= 0; --i) {
            const cplx w = c[i] * b[i + 1];
            a[i] = nmsub(a[i], w, a[i + 1]);
            d[i] = nmsub(d[i], w, d[i + 1]);
            c[i] = -(w * c[i + 1]);
        }
        // ---- interface row: eliminate the next chunk's first unknown with its (normalised) row 0 -----
        sm.pcr[1][0][t] = a[0] * b[0];
        sm.pcr[1][1][t] = c[0] * b[0];
        sm.pcr[1][2][t] = d[0] * b[0];
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
            const cplx inv = frecip(rb);
            rA = rA * inv, rC = rC * inv, rD = rD * inv, rV = rV * inv;
        }
        // ---- PCR over the CTA's 256 interface rows (unit diagonal) ------------------------------------
        int cur = 0;
        sm.pcr[0][0][t] = rA, sm.pcr[0][1][t] = rC, sm.pcr[0][2][t] = rD;
        if (CLUSTER > 1) sm.pcr[0][3][t] = rV;
        __syncthreads();
        for (int lv = 0, s = 1; lv < levels; ++lv, s <<= 1) {
            const bool last = lv == levels - 1;
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
            const cplx inv = frecip(nmsub(nmsub(cone(), rA, Cm), rC, Ap));
            const cplx nD = nmsub(nmsub(rD, rA, Dm), rC, Dp) * inv;
            cplx nV = czero();
            if (CLUSTER > 1) nV = nmsub(nmsub(rV, rA, Vm), rC, Vp) * inv;
            cur ^= 1;
            if (!last) {  // the last level leaves no couplings: A and C are not needed any more
                const cplx nA = -((rA * Am) * inv);
                const cplx nC = -((rC * Cp) * inv);
                rA = nA, rC = nC;
                sm.pcr[cur][0][t] = rA, sm.pcr[cur][1][t] = rC;
            }
            rD = nD, rV = nV;
            sm.pcr[cur][2][t] = rD;
            if (CLUSTER > 1) sm.pcr[cur][3][t] = rV;
            __syncthreads();
        }
        // now x_last(t) = rD - rV * x_partner
        cplx X = rD, x_partner = czero();
        if (CLUSTER > 1) {
            if (t == kT - 1) {
                const uint32_t rbar = map_to_rank(&sm.xchg_bar[parity], rank ^ 1);
                st_async_remote(map_to_rank(&sm.xchg[parity][0], rank ^ 1), rD, rbar);
                st_async_remote(map_to_rank(&sm.xchg[parity][1], rank ^ 1), rV, rbar);
            }
            mbar_wait(&sm.xchg_bar[parity], (it >> 1) & 1);
            const cplx Yp = sm.xchg[parity][0], Vp = sm.xchg[parity][1];
            const cplx Ym = sm.pcr[cur][2][kT - 1], Vm = sm.pcr[cur][3][kT - 1];
            // x_me = Ym - Vm x_p ; x_p = Yp - Vp x_me
            const cplx x_me = nmsub(Ym, Vm, Yp) * frecip(nmsub(cone(), Vm, Vp));
            const cplx x_p = nmsub(Yp, Vp, x_me);
            X = nmsub(rD, rV, x_p);
            x_partner = x_p;
        }
        sm.pcr[cur ^ 1][0][t] = X;
        if (t == 0) tma_store_wait_read();  // the previous system's TMA store has finished reading xs
        __syncthreads();
        const cplx xl = t > 0 ? sm.pcr[cur ^ 1][0][t - 1] : czero();

        // ---- back-substitution + staging -----------------------------------------------------------
        bool bad = !cfinite(X);
        cplx x0 = X, x6 = X;
#pragma unroll
        for (int i = 0; i < kM; ++i) {
            cplx xi = X;
            if (i < kM - 1) xi = nmsub(nmsub(d[i], a[i], xl), c[i], X) * b[i];
            if (i == 0) x0 = xi;
            if (i == kM - 2) x6 = xi;
            bad |= !cfinite(xi);
            if (active) sm.xs[map.xs_slot(t, i)] = xi;
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
        fence_proxy_async();  // make this thread's xs writes visible to the TMA engine
        const int any_nonfinite = __syncthreads_or(bad ? 1 : 0);  // (returns a truth value, not a bit-wise OR)
        if (t == 0) {
            if (map.chunks > 0) {  // (a system shorter than one chunk consists of head rows only)
                tma_store_3d(rank == 0 ? &umaps.r0 : &umaps.r1, 0, rank == 0 ? 0 : map.box_c1, (int)sys, sm.xs);
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
            // rare: one global atomic per affected warp (status[] was zeroed by the host wrapper)
            if (__any_sync(0xffffffffu, off) && (t & 31) == 0 && status != nullptr) atomicOr(&status[sys], 2);
        }
    }
    if (t == 0) tma_store_wait_read();
    if (CLUSTER > 1) cg::this_cluster().sync();  // no CTA leaves while its partner may still address it
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
    const HostMap h = host_map(n, cluster);
    const int chunks = h.chunks[0] > h.chunks[1] ? h.chunks[0] : h.chunks[1];
    const int lv = ceil_log2(chunks < 1 ? 1 : kPer8 * chunks);  // interface rows = thread chunks
    return lv > kLevelsMax ? kLevelsMax : lv;
}

template <int CLUSTER, class Source>
static int launch_batched(const Source &src, const cplx *const bands[4], cplx *u, int64_t n, int64_t batch,
                          int32_t *status, cudaStream_t stream) {
    auto kernel = tridiag_batched_kernel<CLUSTER, Source>;
    const size_t smem = sizeof(BatchedSmem) + 1024;  // slack for the in-kernel 1024-byte round-up
    if (smem > smem_optin_bytes()) {
        set_error("kernel needs %zu B of shared memory, device offers %zu", smem, smem_optin_bytes());
        return FH_E_UNSUPPORTED;
    }
    if (batch >= (1ll << 31)) {
        set_error("batch too large for one launch");
        return FH_E_UNSUPPORTED;
    }
    const HostMap h = host_map(n, CLUSTER);
    const int chunks_total = h.chunks[0] + h.chunks[1];
    TensorMaps maps;
    StoreMaps umaps;
    memset(&maps, 0, sizeof(maps));
    memset(&umaps, 0, sizeof(umaps));
    int rc = make_tensor_map(&umaps.r0, u, n, batch, h.head[0], h.chunks[0], kChunks8);
    if (rc != FH_OK) return rc;
    if (CLUSTER > 1) {
        rc = make_tensor_map(&umaps.r1, u, n, batch, h.head[0], chunks_total, kChunks8);
        if (rc != FH_OK) return rc;
    }
    if (Source::kStaged) {
        CUtensorMap *dst[4] = {&maps.dl, &maps.d, &maps.du, &maps.b};
        for (int i = 0; i < 4; ++i) {
            rc = make_tensor_map(dst[i], bands[i], n, batch, h.head[0], chunks_total, kChunks8);
            if (rc != FH_OK) return rc;
        }
    }
    FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (status != nullptr) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * batch, stream));
    int64_t groups = sm_count() / CLUSTER;  // one CTA per SM (~194 KB of shared memory each)
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
    FH_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, umaps, src, u, n, batch, levels, status));
    return FH_OK;
}

// from fh_tridiag_chain.cu: 4-CTA x 128-thread variant, two CTAs per SM
bool chain_covers(int64_t n);
int chain_solve_bands(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n, int64_t batch,
                      int32_t *status, cudaStream_t stream);
int chain_solve_helm(const HelmSource &src, cplx *u, int64_t n, int64_t batch, int32_t *status, cudaStream_t stream);
// FH_TRIDIAG_KERNEL=pair|chain picks the cluster kernel for systems both cover (development switch)
static bool prefer_chain() {
    const char *e = getenv("FH_TRIDIAG_KERNEL");
    return e != nullptr && strcmp(e, "chain") == 0;
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
    if (prefer_chain() && chain_covers(n)) return chain_solve_bands(dl, d, du, b, u, n, batch, status, stream);
    BandSource src;
    src.dl = dl, src.d = d, src.du = du, src.b = b;
    const cplx *const bands[4] = {dl, d, du, b};
    if (n <= kStageRows) return launch_batched<1>(src, bands, u, n, batch, status, stream);
    return launch_batched<2>(src, bands, u, n, batch, status, stream);
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
        set_error("fused sweep supports meshes up to %lld nodes; use fh_helmholtz1d_solve_large_fused_c128 beyond",
                  (long long)kBatchedMaxN);
        return FH_E_UNSUPPORTED;
    }
    src.ks = ks, src.k2s = k2s;
    const cplx *const none[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (prefer_chain() && chain_covers(n)) return chain_solve_helm(src, reinterpret_cast<cplx *>(u), n, n_k, status, st);
    if (n <= kStageRows) return launch_batched<1>(src, none, reinterpret_cast<cplx *>(u), n, n_k, status, st);
    return launch_batched<2>(src, none, reinterpret_cast<cplx *>(u), n, n_k, status, st);
}

}  // extern "C"
