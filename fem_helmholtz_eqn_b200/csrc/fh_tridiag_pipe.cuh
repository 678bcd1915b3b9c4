// The batched cluster solver as a PIPELINE of warp roles (included by fh_tridiag_batched.cu; same algorithm, same
// row -> CTA -> thread maps, same tensor maps as tridiag_batched_kernel).
//
// Why: one system kept a CTA busy for ~8,800 cycles of which every phase is a dependent chain (8 warps, all
// registers taken), while the memory side needed ~7,300: latency-bound at 0.83 of the HBM roofline, and nothing could
// overlap because registers + shared memory hold exactly one system in flight and one in prefetch.  Here the phases of
// CONSECUTIVE systems overlap on one SM:
//
//   warps 0..7   "E/B"  elimination of system s+1 (tile reads, downward / upward sweep, interface row), then
//                       back-substitution + probe + TMA store of system s;
//   warps 8..11  "I"    interface system of system s+1 in shared memory (cyclic-reduction step, PCR with one thread
//                       per surviving row, cluster exchange through DSMEM, cut solve), concurrently with B(s), E(s+2);
//   warps 12..15 "S"    (one-pass assemble+solve only) the band stores of A: tiles of identical rows filled and sent
//                       by TMA, decoupled from everything else.
//
// What makes it fit: the state a thread must keep between its elimination and its back-substitution (7 rows x
// (a, c, d) pre-scaled by the pivot reciprocal + the original interface row for the probe = 50 doubles) is PARKED IN
// TENSOR MEMORY (tcgen05.st / tcgen05.ld, 32x32b: every thread owns one lane x 256 columns; two parking buffers of 100
// columns).  The 256 KB of TMEM are otherwise idle in this kernel; measured (tools/tmem_probe.cu) 466 cycles to park
// and 295 to un-park the 100 KB of a CTA.  Register budgets are re-balanced with setmaxnreg (E/B 200, I 96 of a
// 384-thread CTA; E/B 192, I 88, S 40 of the 512-thread one-pass CTA).  Roles hand over through named barriers
// (bar.arrive by the producer, bar.sync by the consumer):
//   FullFree  I -> E   `full` (raw interface rows) consumed        FullReady  E -> I   `full` written
//   X0 / X1   I -> B   interface unknowns of an even / odd system are in `xi`
#pragma once

#include <type_traits>

#include "fh_tmem.cuh"

namespace fh {

constexpr int kEBThreads = kT;   // 256
constexpr int kIThreads = kH;    // 128: one per interface row that survives the cyclic-reduction step
constexpr int kSThreads = 128;
constexpr int kParkDoubles = 6 * (kM - 1) + 8;   // 50
constexpr int kParkCols = 2 * kParkDoubles;      // 100 columns of 32 bits per thread and system
constexpr uint32_t kParkStride = 128;            // columns between a thread's two parking buffers
static_assert(kParkCols == 100 && kParkCols <= (int)kParkStride, "parking layout");

enum : int { kBarEB = 1, kBarI = 2, kBarFullReady = 3, kBarFullFree = 4, kBarX0 = 5, kBarX1 = 6, kBarS = 7 };
constexpr int kRegsEB = 200, kRegsI = 96;                      // 256 * 200 + 128 * 96 <= 384 * 168 = what the CTA was
                                                               // launched with: the pool setmaxnreg re-distributes (a
                                                               // budget that sums to more, e.g. 208 / 96, makes
                                                               // setmaxnreg.inc wait for ever -- measured)
constexpr int kRegsEB1 = 184, kRegsI1 = 88, kRegsS1 = 56;      // 256 * 184 + 128 * 88 + 128 * 56 = 512 * 128

// PCR arrays of the I group: kH rows each, kHalfPad zero rows before and after (the largest PCR stride); the pad behind
// one array is the pad in front of the next
constexpr int kPcrArr = kH + kHalfPad;                 // entries from one array to the next: 192
constexpr int kPcrEntries = kHalfPad + 4 * kPcrArr;    // 832

struct __align__(1024) PipeSmem {
    cplx stage[4][kSlots];      // TMA destination: bands of the NEXT system (one-pass: 2 x 4 band tiles of 16 KB)
    cplx xs[kSlots];            // TMA source: solution of the system being back-substituted
    cplx row0[3][kT];           // normalised first rows of the chunks (read by the thread above)
    cplx full[4][kT + 1];       // raw interface rows A, C, D, V; entry kT is a zero row
    cplx half[2][kPcrEntries];  // PCR ping-pong over the odd interface rows: [pad][A][pad][C][pad][D][pad][V][pad]
    cplx xi[2][kT + 2];         // [parity]: 0, the kT interface unknowns, x_partner
    cplx headst[4][kHeadMax];   // head rows of the next system (cp.async by one thread)
    cplx headf[2][2][kHeadMax]; // [parity]: elimination factors of the head rows (kept for the back-substitution)
    cplx xchg[2][2];            // [parity][{Y, V}] written by the partner CTA (st.async)
    RowGeom bgeom[3];           // generated rows, uniform mesh: geometry of mesh row 0, row N and the interior rows
    unsigned long long full_bar;
    unsigned long long xchg_bar[2];
    uint32_t tmem_base;
    uint32_t i_done;            // systems the I group has finished (paces the S group)
};

__device__ __forceinline__ int named_bar_or(int id, int threads, bool pred) {
    int r;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "bar.red.or.pred q, %2, %3, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t"
        "}"
        : "=r"(r)
        : "r"((int)pred), "r"(id), "r"(threads)
        : "memory");
    return r;
}
__device__ __forceinline__ void tma_store_3d_from(const CUtensorMap *map, int c0, int c1, int c2, uint32_t smem_addr) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr)
                 : "memory");
}
// u += tile: the TMA engine hands the tile to L2, which carries out the IEEE additions (f64 tensor map)
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void red_add(cplx *p, cplx v) {  // *p += v (two fp64 reductions in L2)
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&p->re), "d"(v.re) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&p->im), "d"(v.im) : "memory");
}
__device__ __forceinline__ cplx ldg_cg(const cplx *p) {  // coherent (L2) read of data this kernel also updates elsewhere
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return mk(v.x, v.y);
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__device__ __forceinline__ void park100(uint32_t ta, const uint32_t *v) {
    tmem_st32(ta, v), tmem_st32(ta + 32, v + 32), tmem_st32(ta + 64, v + 64), tmem_st4(ta + 96, v + 96);
}
__device__ __forceinline__ void unpark100(uint32_t ta, uint32_t *v) {
    tmem_ld32(ta, v), tmem_ld32(ta + 32, v + 32), tmem_ld32(ta + 64, v + 64), tmem_ld4(ta + 96, v + 96);
}
__device__ __forceinline__ void put2(uint32_t *v, int i, cplx z) {
    v[4 * i] = (uint32_t)__double2loint(z.re), v[4 * i + 1] = (uint32_t)__double2hiint(z.re);
    v[4 * i + 2] = (uint32_t)__double2loint(z.im), v[4 * i + 3] = (uint32_t)__double2hiint(z.im);
}
__device__ __forceinline__ cplx get2(const uint32_t *v, int i) {
    return mk(__hiloint2double((int)v[4 * i + 1], (int)v[4 * i]), __hiloint2double((int)v[4 * i + 3], (int)v[4 * i + 2]));
}

// v[j] <- v[j ^ c], c in 0..7 (three conditional exchange stages; register arrays stay statically indexed)
__device__ __forceinline__ void xor_permute8(cplx (&v)[8], int c) {
#pragma unroll
    for (int bit = 1; bit < 8; bit <<= 1) {
        const bool sw = (c & bit) != 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if ((m & bit) == 0) {
                const cplx a = v[m], b = v[m | bit];
                v[m] = mk(sw ? b.re : a.re, sw ? b.im : a.im);
                v[m | bit] = mk(sw ? a.re : b.re, sw ? a.im : b.im);
            }
        }
    }
}
__device__ __forceinline__ cplx shfl_from(cplx z, int lane) {
    return mk(__shfl_sync(0xffffffffu, z.re, lane), __shfl_sync(0xffffffffu, z.im, lane));
}
__device__ __forceinline__ cplx shfl_up1(cplx z) {
    return mk(__shfl_up_sync(0xffffffffu, z.re, 1), __shfl_up_sync(0xffffffffu, z.im, 1));
}
__device__ __forceinline__ cplx shfl_down1(cplx z) {
    return mk(__shfl_down_sync(0xffffffffu, z.re, 1), __shfl_down_sync(0xffffffffu, z.im, 1));
}
// In: v[k] = row 32 k + lane of a warp's 256 rows (how coalesced loads deliver them).  Out: v[i] = row 8 lane + i (how
// the elimination owns them).  Thread T wants register k = T >> 2 of the lanes 8 (T & 3) + i; in step j it reads lane
// 8 (T & 3) + (j ^ (T >> 2)), which offers its register j ^ (lane & 7) -- i.e. register T >> 2: eight shuffles in which
// every lane sends and receives something useful, framed by two XOR permutations of the registers.
__device__ __forceinline__ void warp_rows_to_threads(cplx (&v)[8], int lane) {
    xor_permute8(v, lane & 7);
    cplx w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = shfl_from(v[j], 8 * (lane & 3) + (j ^ (lane >> 2)));
    xor_permute8(w, lane >> 2);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = w[j];
}

template <class Source>
constexpr int pipe_threads() {
    return kEBThreads + kIThreads + (Source::kStoreBands ? kSThreads : 0);
}

#ifdef FH_TRACE
// role traces: E/B warp leaders stamp slots 0..15 (elimination 0..5, back-substitution 6..11), the first I warp 16..27
#define FH_TRP(w, k)                                                                                          \
    do {                                                                                                      \
        if (g_trace != nullptr && it == (uint32_t)g_trace_iter && blockIdx.x < 2 && (t & 31) == 0) {          \
            unsigned long long gt__;                                                                          \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__));                                          \
            g_trace[((blockIdx.x * 16 + (w)) * 16 + (k)) * 2] = clock64();                                    \
            g_trace[((blockIdx.x * 16 + (w)) * 16 + (k)) * 2 + 1] = (long long)gt__;                          \
        }                                                                                                     \
    } while (0)
#else
#define FH_TRP(w, k) do { } while (0)
#endif

template <int CLUSTER, class Source>
__global__ void __launch_bounds__(pipe_threads<Source>(), 1)
tridiag_pipe_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ StoreMaps umaps, Source src,
                    cplx *__restrict__ u, int64_t n, int64_t batch, int levels, int32_t *__restrict__ status) {
    static_assert(Source::kStaged || Source::kUniform, "generated rows: the pipeline covers uniform meshes only");
    extern __shared__ unsigned char smem_raw[];
    PipeSmem &sm = *reinterpret_cast<PipeSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    const int t = threadIdx.x, warp = t >> 5;
    int rank = 0;
    int64_t first_sys = blockIdx.x, sys_stride = gridDim.x;
    if (CLUSTER > 1) {
        rank = (int)cg::this_cluster().block_rank();
        first_sys = blockIdx.x / CLUSTER;
        sys_stride = gridDim.x / CLUSTER;
    }
    const LocalMap map = make_map<CLUSTER>(n, rank);
    const bool has_cut = CLUSTER > 1 && (n & 1);
    const int64_t cut_row = n / 2;

    for (int i = t; i < 2 * kPcrEntries; i += (int)blockDim.x) {  // zero rows beyond either end of the PCR arrays
        const int c = i % kPcrEntries;
        if (c < kHalfPad || (c - kHalfPad) % kPcrArr >= kH) (&sm.half[0][0])[i] = czero();
    }
    if (t < 4) sm.full[t][kT] = czero();
    if (t < 2) sm.xi[t][0] = czero();
    if (t == 0) {
        mbar_init(&sm.full_bar, 1);
        mbar_init(&sm.xchg_bar[0], 1);
        mbar_init(&sm.xchg_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t == 0) sm.i_done = 0;
    if (warp == 0) tmem_alloc(&sm.tmem_base, 512);  // one CTA per SM: all of tensor memory
    if constexpr (Source::kUniform) {
        if (t < 3) sm.bgeom[t] = row_geom(src.p, t == 0 ? 0 : t == 1 ? src.p.n_elem : 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CLUSTER > 1) cg::this_cluster().sync();  // partner's barriers exist before anything is sent
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    auto matrix_of = [&](int64_t sys) -> int64_t {
        if constexpr (Source::kSubset) return src.list[sys];
        else return sys;
    };
    if constexpr (Source::kSubset) batch = *src.count < batch ? *src.count : batch;

    if (t < kEBThreads) {
        // =============================================================================================== E/B
        if constexpr (Source::kStoreBands) reg_inc<kRegsEB1>();
        else reg_inc<kRegsEB>();
        const bool active = t >= map.first_active;
        const CUtensorMap *tmaps = maps.band[rank];
        const CUtensorMap *map_lo = map.dir > 0 ? &tmaps[0] : &tmaps[2];
        const CUtensorMap *map_hi = map.dir > 0 ? &tmaps[2] : &tmaps[0];
        const int feeder = map.first_active % kT;
        // this warp's lane quadrant of tensor memory; warps w and w + 4 share a quadrant and split its columns
        const uint32_t tpark = sm.tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(warp >> 2) * 256u;
        // Generated rows on a uniform mesh: every interior row of a system holds the same three REAL numbers and a zero
        // right-hand side, so every chunk that holds interior rows only is eliminated to the same seven (a', c') pairs
        // and the same interface row.  A warp whose 32 chunks are all of that kind ("plain warp") carries the sweeps out
        // on the real parts alone -- the imaginary lanes of the complex arithmetic would only move zeros, so the real
        // parts come out bit for bit as in the complex sweeps -- and keeps the 17 numbers its back-substitution needs
        // in registers instead of parking 50 complex ones in tensor memory.  The warp that holds the CTA's first
        // active chunk (mesh row 0 / N, or the head rows folded into it), warps with inactive threads and, single-CTA,
        // the last warp (mesh row N) take the general complex path below.
        bool plain = false;
        if constexpr (Source::kUniform) plain = 32 * warp > map.first_active && (CLUSTER > 1 || warp != kEBThreads / 32 - 1);

        auto prefetch = [&](int64_t sys) {
            if constexpr (Source::kStaged) {
                if (t == feeder) {
                    const int64_t sysm = matrix_of(sys);
                    mbar_expect_tx(&sm.full_bar, 4 * kTileBytes);
                    tma_load_3d(sm.stage[0], map_lo, 0, map.box_c1, (int)sysm, &sm.full_bar);
                    tma_load_3d(sm.stage[1], &tmaps[1], 0, map.box_c1, (int)sysm, &sm.full_bar);
                    tma_load_3d(sm.stage[2], map_hi, 0, map.box_c1, (int)sysm, &sm.full_bar);
                    const int64_t sysb = Source::kRefine ? sysm : sys;  // right-hand sides: compact unless refining in place
                    tma_load_3d(sm.stage[3], &tmaps[3], 0, map.box_c1, (int)sysb, &sm.full_bar);
                    const cplx *lo = map.dir > 0 ? src.dl : src.du, *hi = map.dir > 0 ? src.du : src.dl;
                    for (int j = 0; j < map.head; ++j) {
                        const int64_t g = sysm * n + map.global_row(j);
                        cp_async16(&sm.headst[0][j], lo + g);
                        cp_async16(&sm.headst[1][j], src.d + g);
                        cp_async16(&sm.headst[2][j], hi + g);
                        cp_async16(&sm.headst[3][j], src.b + sysb * n + map.global_row(j));
                    }
                    cp_async_commit();
                }
            }
        };
        if (first_sys < batch) prefetch(first_sys);

        // the loop over the systems, compiled once for plain warps and once for the others: what a plain warp carries
        // from a system's elimination to its back-substitution must not weigh on the register budget of the general path
        auto eb_loop = [&](auto plain_tag) {
        constexpr bool kPlain = decltype(plain_tag)::value;
        constexpr int kCarry = kPlain ? kM - 1 : 1;
        double ca[kCarry], cc[kCarry], c3[3];  // plain warps: a'_i, c'_i (i < 7) and the interior row (a, b, c) of `sys` ...
        double pa[kCarry], pc[kCarry], p3[3];  // ... and of `prev`
#pragma unroll
        for (int i = 0; i < kCarry; ++i) ca[i] = cc[i] = pa[i] = pc[i] = 0.0;
        c3[0] = c3[1] = c3[2] = p3[0] = p3[1] = p3[2] = 0.0;
        // refinement in one pass: solution value of local row j of the system whose solution starts at `us`
        [[maybe_unused]] auto u_at = [&](const cplx *us, int j) -> cplx {
            const int64_t g = map.global_row(j);
            return g >= 0 && g < n ? ldg_cg(us + g) : czero();
        };
        uint32_t it = 0;
        int64_t prev = -1;
        for (int64_t sys = first_sys;; sys += sys_stride, ++it) {
            const bool have = sys < batch;
            if (have) {
                // ------------------------------------------------------------------------ E: elimination of `sys`
                const uint32_t parity = it & 1;
                double hk = 0.0, hk2 = 0.0;
                cplx il = czero(), id = cone(), iu = czero(), ib = czero();  // the interior row (generated, uniform mesh)
                if constexpr (Source::kUniform) {
                    hk = __ldg(src.ks + sys), hk2 = __ldg(src.k2s + sys);
                    row_values(src.p, sm.bgeom[2], 1, hk, hk2, il, id, iu, ib);
                }
                FH_TRP(warp, 0);
                cplx a7, b7, c7, d7;  // the thread's interface row after the sweeps (b7 not inverted)
                if constexpr (kPlain) {
                    const double ra = map.dir > 0 ? il.re : iu.re, rb = id.re, rc = map.dir > 0 ? iu.re : il.re;
                    FH_TRP(warp, 1);
                    double A[kM], B[kM], C[kM];
                    A[0] = ra, C[0] = rc;
                    B[0] = rb * fast_rcp(rb * rb);  // (the real lane of frecip)
#pragma unroll
                    for (int i = 1; i < kM; ++i) {
                        const double w = ra * B[i - 1];
                        B[i] = fma(-w, rc, rb);
                        A[i] = -(w * A[i - 1]);
                        C[i] = rc;
                        if (i < kM - 1) B[i] = B[i] * fast_rcp(B[i] * B[i]);
                    }
                    FH_TRP(warp, 2);
                    FH_TRP(warp, 3);
#pragma unroll
                    for (int i = kM - 3; i >= 0; --i) {
                        const double w = C[i] * B[i + 1];
                        A[i] = fma(-w, A[i + 1], A[i]);
                        C[i] = -(w * C[i + 1]);
                    }
#pragma unroll
                    for (int i = 0; i < kM - 1; ++i) ca[i] = A[i] * B[i], cc[i] = C[i] * B[i];
                    c3[0] = ra, c3[1] = rb, c3[2] = rc;
                    sm.row0[0][t] = mk(ca[0], 0.0);
                    sm.row0[1][t] = mk(cc[0], 0.0);
                    sm.row0[2][t] = czero();
                    a7 = mk(A[kM - 1], 0.0), b7 = mk(B[kM - 1], 0.0), c7 = mk(rc, 0.0), d7 = czero();
                } else {
                    cplx a[kM], b[kM], c[kM], d[kM];
                    uint32_t tile_addr = 0, tile_x = 0;
                    // refinement in one pass: the solution values the thread's rows touch (local rows L - 1 .. L + 8).
                    // A thread's rows are 128 contiguous bytes, so fetching them thread by thread costs a warp ten
                    // requests of 32 scattered sectors each (measured: 3.7 us per system, as much as the solve).  The
                    // warp therefore reads its 256 rows with eight coalesced requests (lane l: rows 32 k + l), issued
                    // before the wait for the tiles, and transposes them in registers (warp_rows_to_threads: 32
                    // shuffles); the two rows beyond the warp's ends are fetched by its first and last lane.
                    // rhs_of() turns a row's b into b - A u with the operation order of residual_subset_kernel
                    // (diagonal, sub-diagonal, super-diagonal of the GLOBAL numbering).
                    [[maybe_unused]] cplx uo[kM + 2], ur[kM];
                    [[maybe_unused]] const cplx *uold = nullptr;
                    [[maybe_unused]] int64_t m_next = -1;
                    [[maybe_unused]] const int Lrow = map.head + kM * (t - map.first_active);
                    [[maybe_unused]] auto u_local = [&](int j) -> cplx { return u_at(uold, j); };
                    [[maybe_unused]] auto rhs_of = [&](int64_t g, cplx ra, cplx rb, cplx rc, cplx rd, cplx um, cplx u0, cplx up) -> cplx {
                        cplx v = nmsub(rd, rb, u0);
                        if (g > 0) v = nmsub(v, map.dir > 0 ? ra : rc, map.dir > 0 ? um : up);
                        if (g < n - 1) v = nmsub(v, map.dir > 0 ? rc : ra, map.dir > 0 ? up : um);
                        return v;
                    };
                    if constexpr (Source::kRefine) {
                        uold = src.u_old + matrix_of(sys) * n;
                        if (sys + sys_stride < batch) m_next = src.list[sys + sys_stride];
                        const int lane = t & 31, Lw = map.head + kM * (32 * warp - map.first_active);
#pragma unroll
                        for (int k = 0; k < kM; ++k) ur[k] = u_local(Lw + 32 * k + lane);
                        uo[0] = lane == 0 ? u_local(Lw - 1) : czero();
                        uo[kM + 1] = lane == 31 ? u_local(Lw + 32 * kM) : czero();
                    }
                    if constexpr (Source::kStaged) {
                        mbar_wait(&sm.full_bar, parity);
                        if (map.head > 0 && t == feeder) cp_async_wait_all();
                        const int tt = map.dir > 0 ? t : kT - 1 - t;
                        const int r = tt / kPer8;
                        tile_addr = smem_u32(&sm.stage[0][0]) + (uint32_t)r * 128u;
                        tile_x = (uint32_t)(r & 7);
                    }
                    if constexpr (Source::kRefine) {
                        warp_rows_to_threads(ur, t & 31);
#pragma unroll
                        for (int i = 0; i < kM; ++i) uo[i + 1] = ur[i];
                        const cplx up = shfl_up1(ur[kM - 1]), dn = shfl_down1(ur[0]);
                        if ((t & 31) != 0) uo[0] = up;
                        if ((t & 31) != 31) uo[kM + 1] = dn;
                    }
                    FH_TRP(warp, 1);
                    // generated rows: mesh row g in the CTA's local orientation (only the threads that own a mesh boundary
                    // row or head rows come here)
                    auto gen_row = [&](int64_t g, cplx &ra, cplx &rb, cplx &rc, cplx &rd) {
                        if constexpr (Source::kUniform) {
                            cplx lo = il, hi = iu;
                            rb = id, rd = ib;
                            if (g == 0 || g == n - 1) row_values(src.p, sm.bgeom[g == 0 ? 0 : 1], g, hk, hk2, lo, rb, hi, rd);
                            ra = map.dir > 0 ? lo : hi, rc = map.dir > 0 ? hi : lo;
                        }
                    };
                    auto load_row = [&](int i) {
                        if (!active) {
                            a[i] = czero(), b[i] = cone(), c[i] = czero(), d[i] = czero();
                        } else if constexpr (Source::kStaged) {
                            const int tt = map.dir > 0 ? t : kT - 1 - t;
                            const uint32_t e = (uint32_t)((map.dir > 0 ? i : kM - 1 - i) + kM * (tt % kPer8));
                            const uint32_t ad = tile_addr + ((e ^ tile_x) << 4);
                            a[i] = lds_volatile(ad), b[i] = lds_volatile(ad + kTileBytes);
                            c[i] = lds_volatile(ad + 2 * kTileBytes), d[i] = lds_volatile(ad + 3 * kTileBytes);
                            if constexpr (Source::kRefine)
                                d[i] = rhs_of(map.global_row(Lrow + i), a[i], b[i], c[i], d[i], uo[i], uo[i + 1], uo[i + 2]);
                        } else {
                            a[i] = map.dir > 0 ? il : iu, b[i] = id, c[i] = map.dir > 0 ? iu : il, d[i] = ib;
                            // a mesh boundary row can only sit in the first slot of the first active thread (local row 0,
                            // when there are no head rows) or, single-CTA, in the last slot of the last thread
                            if (i == 0 && map.head == 0 && t == map.first_active) gen_row(map.global_row(0), a[i], b[i], c[i], d[i]);
                            if (i == kM - 1 && CLUSTER == 1 && t == kT - 1) gen_row(n - 1, a[i], b[i], c[i], d[i]);
                        }
                        if (i == 0 && map.head == 0 && t == map.first_active) a[0] = czero();
                        if (i == kM - 1 && CLUSTER == 1 && t == kT - 1) c[kM - 1] = czero();
                    };
                    load_row(0);
                    load_row(1);
                    // head rows (far end, at most 8): serial Thomas by one thread, folded into its row 0; the factors wait
                    // in shared memory for the back-substitution
                    if (map.head > 0 && t == feeder) {
                        cplx cp = czero(), dp = czero();
#pragma unroll 1
                        for (int j = 0; j < map.head; ++j) {
                            cplx ha, hb, hc, hd;
                            if constexpr (Source::kStaged) {
                                ha = sm.headst[0][j], hb = sm.headst[1][j], hc = sm.headst[2][j], hd = sm.headst[3][j];
                                if constexpr (Source::kRefine)
                                    hd = rhs_of(map.global_row(j), ha, hb, hc, hd, u_local(j - 1), u_local(j), u_local(j + 1));
                            } else {
                                gen_row(map.global_row(j), ha, hb, hc, hd);
                            }
                            if (j == 0) ha = czero();
                            if (CLUSTER == 1 && j == map.n_loc - 1) hc = czero();
                            const cplx inv = frecip(nmsub(hb, ha, cp));
                            dp = nmsub(hd, ha, dp) * inv;
                            cp = hc * inv;
                            sm.headf[parity][0][j] = cp, sm.headf[parity][1][j] = dp;
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
                    FH_TRP(warp, 2);
                    if constexpr (Source::kRefine) {  // the next system's solution rows on their way into L2 (one line per lane:
                        if (m_next >= 0) {            // carrying the values themselves across the iteration spilled, and was slower)
                            const int64_t g = map.global_row(map.head + kM * (t - map.first_active) + (map.dir > 0 ? 0 : kM - 1));
                            if (g >= 0 && g < n)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(src.u_old + m_next * n + g) : "memory");
                        }
                    }
                    if constexpr (Source::kStaged) {
                        named_bar_sync(kBarEB, kEBThreads);  // every thread has drained the staging tiles ...
                        if (sys + sys_stride < batch) prefetch(sys + sys_stride);  // ... refill them
                    }
                    FH_TRP(warp, 3);
#pragma unroll
                    for (int i = kM - 3; i >= 0; --i) {
                        const cplx w = c[i] * b[i + 1];
                        a[i] = nmsub(a[i], w, a[i + 1]);
                        d[i] = nmsub(d[i], w, d[i + 1]);
                        c[i] = -(w * c[i + 1]);
                    }
                    // rows 0..6 in the form the back-substitution wants: x_i = d'_i - a'_i x_left - c'_i x_7
                    uint32_t pk[kParkCols];
#pragma unroll
                    for (int i = 0; i < kM - 1; ++i) {
                        a[i] = a[i] * b[i], c[i] = c[i] * b[i], d[i] = d[i] * b[i];
                        put2(pk, 3 * i, a[i]), put2(pk, 3 * i + 1, c[i]), put2(pk, 3 * i + 2, d[i]);
                    }
                    put2(pk, 21, a7o), put2(pk, 22, b7o), put2(pk, 23, c7o), put2(pk, 24, d7o);
                    sm.row0[0][t] = a[0];
                    sm.row0[1][t] = c[0];
                    sm.row0[2][t] = d[0];
                    tmem_wait_st();  // (the previous system's park has long completed: orders it before its un-park below)
                    park100(tpark + parity * kParkStride, pk);
                    a7 = a[kM - 1], b7 = b[kM - 1], c7 = c[kM - 1], d7 = d[kM - 1];
                }
                if (t == kStoreThread) tma_store_wait_read();  // the previous TMA store has finished reading xs
                FH_TRP(warp, 4);
                // all E/B threads + the I group's "full consumed": row0 is visible, xs and full are free
                named_bar_sync(kBarFullFree, kEBThreads + kIThreads);
                FH_TRP(warp, 5);
                {
                    cplx rb = b7;
                    cplx rA = a7, rC = czero(), rD = d7, rV = czero();
                    if (t < kT - 1) {
                        const cplx nA = sm.row0[0][t + 1], nC = sm.row0[1][t + 1], nD = sm.row0[2][t + 1];
                        rb = nmsub(rb, c7, nA);
                        rC = -(c7 * nC);
                        rD = nmsub(rD, c7, nD);
                    } else if (CLUSTER > 1) {
                        rV = c7;  // coupling to the partner CTA's interface unknown / the cut unknown
                    }
                    const cplx inv = frecip(rb);
                    sm.full[0][t] = rA * inv, sm.full[1][t] = rC * inv, sm.full[2][t] = rD * inv;
                    if (CLUSTER > 1) sm.full[3][t] = rV * inv;
                }
                named_bar_arrive(kBarFullReady, kEBThreads + kIThreads);
                FH_TRP(warp, 6);
            }
            if (prev >= 0) {
                // ------------------------------------------------------------------------ B: back-substitution of `prev`
                const uint32_t pp = (it - 1) & 1;
                const int64_t pm = Source::kRefine ? matrix_of(prev) : prev;  // where the solution and its status live
                uint32_t pk[kParkCols];
                if (!have) {  // last system: no elimination ran in this iteration (its waits are made up for here)
                    tmem_wait_st();
                    if (t == kStoreThread) tma_store_wait_read();
                    named_bar_sync(kBarEB, kEBThreads);  // (orders the store wait before the writes to xs)
                }
                if constexpr (!kPlain) unpark100(tpark + pp * kParkStride, pk);
                named_bar_sync(pp ? kBarX1 : kBarX0, kEBThreads + kIThreads);  // the I group has solved the interface
                FH_TRP(warp, 7);
                const cplx X = sm.xi[pp][1 + t], xl = sm.xi[pp][t], x_partner = sm.xi[pp][kT + 1];
                bool bad = !cfinite(X) || !cfinite(x_partner);
                cplx xr[kM];
                cplx a7o, b7o, c7o, d7o;
                if constexpr (kPlain) {
#pragma unroll
                    for (int i = 0; i < kM - 1; ++i)  // (the real coefficients' share of the complex nmsub chain)
                        xr[i] = mk(fma(-pc[i], X.re, fma(-pa[i], xl.re, 0.0)), fma(-pc[i], X.im, fma(-pa[i], xl.im, 0.0)));
                    a7o = mk(p3[0], 0.0), b7o = mk(p3[1], 0.0), c7o = mk(p3[2], 0.0), d7o = czero();
                } else {
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < kM - 1; ++i)
                        xr[i] = nmsub(nmsub(get2(pk, 3 * i + 2), get2(pk, 3 * i), xl), get2(pk, 3 * i + 1), X);
                    a7o = get2(pk, 21), b7o = get2(pk, 22), c7o = get2(pk, 23), d7o = get2(pk, 24);
                }
                xr[kM - 1] = X;
                const cplx x0 = xr[0], x6 = xr[kM - 2];
                if (active) {
                    const int tt = map.dir > 0 ? t - map.first_active : kT - 1 - t;
                    const int r = tt / kPer8;
                    const uint32_t base = smem_u32(&sm.xs[0]) + (uint32_t)r * 128u, x = (uint32_t)(r & 7);
#pragma unroll
                    for (int i = 0; i < kM; ++i) {
                        const uint32_t e = (uint32_t)((map.dir > 0 ? i : kM - 1 - i) + kM * (tt % kPer8));
                        sts_volatile(base + ((e ^ x) << 4), xr[i]);
                    }
                }
                if (map.head > 0 && t == feeder) {
                    cplx xn = x0;
                    cplx *out = u + pm * n;
#pragma unroll 1
                    for (int j = map.head - 1; j >= 0; --j) {
                        const cplx hcp = sm.headf[pp][0][j], hdp = sm.headf[pp][1][j];
                        xn = (map.chunks == 0 && j == map.head - 1) ? hdp : nmsub(hdp, hcp, xn);
                        bad |= !cfinite(xn);
                        if constexpr (Source::kRefine) red_add(out + map.global_row(j), xn);
                        else stg_stream(out + map.global_row(j), xn);
                    }
                }
                fence_proxy_async();  // make this thread's xs writes visible to the TMA engine
                const int any_nonfinite = named_bar_or(kBarEB, kEBThreads, bad);
                FH_TRP(warp, 8);
                if (t == kStoreThread) {
                    if (map.chunks > 0) {
                        if constexpr (Source::kRefine) tma_reduce_add_3d(&umaps.r[rank], 0, 0, (int)pm, sm.xs);
                        else tma_store_3d(&umaps.r[rank], 0, 0, (int)pm, sm.xs);
                        tma_store_commit();
                    }
                    if constexpr (Source::kRefine) atomicAnd(&status[pm], ~2);  // refined: the small-pivot flag goes
                    if (status != nullptr && any_nonfinite) atomicOr(&status[pm], 1);
                }
                {   // probe: defect of the original interface row with the computed solution
                    cplx xn = x_partner;
                    if (active && t < kT - 1) xn = sm.xs[map.xs_slot(t + 1, 0)];
                    const cplx r = nmsub(nmsub(nmsub(d7o, a7o, x6), b7o, X), c7o, xn);
                    const double scale = abs1(a7o) * abs1(x6) + abs1(b7o) * abs1(X) + abs1(c7o) * abs1(xn) + abs1(d7o);
                    const bool off = active && abs1(r) > kProbeTol * scale;
                    const bool severe = active && abs1(r) > kSevereTol * scale;
                    const bool any_off = __any_sync(0xffffffffu, off), any_severe = __any_sync(0xffffffffu, severe);
                    if constexpr (Source::kRefine) {  // a correction solve that breaks down again: pivoted re-solve
                        if (any_severe && (t & 31) == 0) atomicOr(&status[pm], 1);
                    } else {
                        if (any_off && (t & 31) == 0 && status != nullptr) atomicOr(&status[pm], any_severe ? 3 : 2);
                    }
                }
                FH_TRP(warp, 9);
            }
            if (!have) break;
            prev = sys;
            if constexpr (kPlain) {
#pragma unroll
                for (int i = 0; i < kM - 1; ++i) pa[i] = ca[i], pc[i] = cc[i];
                p3[0] = c3[0], p3[1] = c3[1], p3[2] = c3[2];
            }
        }
        };
        if constexpr (Source::kUniform) {
            if (plain) eb_loop(std::true_type{});
            else eb_loop(std::false_type{});
        } else {
            eb_loop(std::false_type{});
        }
        if (t == kStoreThread) tma_store_wait_read();
    } else if (t < kEBThreads + kIThreads) {
        // =============================================================================================== I
        if constexpr (Source::kStoreBands) reg_dec<kRegsI1>();
        else reg_dec<kRegsI>();
        const int j = t - kEBThreads;  // row among the kH odd interface rows: interface row r = 2 j + 1
        const int r = 2 * j + 1;
        constexpr uint32_t kArr = kPcrArr * sizeof(cplx);
        const uint32_t hb0 = smem_u32(&sm.half[0][kHalfPad + j]), hb1 = smem_u32(&sm.half[1][kHalfPad + j]);
        named_bar_arrive(kBarFullFree, kEBThreads + kIThreads);  // nothing to protect before the first system
        uint32_t it = 0;
        for (int64_t sys = first_sys; sys < batch; sys += sys_stride, ++it) {
            const uint32_t parity = it & 1;
            if (CLUSTER > 1 && j == 0) mbar_expect_tx(&sm.xchg_bar[parity], 2 * sizeof(cplx));
            // the cut row (lo x(rank 0 side) + diag x_cut + hi x(rank 1 side) = rhs): requested before the wait
            cplx cl = czero(), cb = cone(), ch = czero(), cd = czero();
            if (has_cut) {
                if constexpr (Source::kStaged) {
                    const int64_t g = matrix_of(sys) * n + cut_row;
                    cl = ldg_stream(src.dl + g), cb = ldg_stream(src.d + g), ch = ldg_stream(src.du + g);
                    if constexpr (Source::kRefine) {
                        const cplx *uc = src.u_old + g;
                        cd = nmsub(nmsub(nmsub(ldg_stream(src.b + g), cb, ldg_cg(uc)), cl, ldg_cg(uc - 1)), ch, ldg_cg(uc + 1));
                    } else {
                        cd = ldg_stream(src.b + sys * n + cut_row);
                    }
                } else {  // uniform mesh: the cut row is an interior row (n >= 2057 here)
                    row_values(src.p, sm.bgeom[2], 1, __ldg(src.ks + sys), __ldg(src.k2s + sys), cl, cb, ch, cd);
                }
            }
            FH_TRP(8 + (j >> 5), 0);
            named_bar_sync(kBarFullReady, kEBThreads + kIThreads);
            FH_TRP(8 + (j >> 5), 1);
            // cyclic-reduction step: the odd row r absorbs its even neighbours r - 1 and r + 1 (row kT is a zero row);
            // the even row below stays in registers and is solved from its own equation at the end
            cplx A, C, D, V = czero();
            const cplx eA = sm.full[0][r - 1], eC = sm.full[1][r - 1], eD = sm.full[2][r - 1];
            {
                const cplx oA = sm.full[0][r], oC = sm.full[1][r], oD = sm.full[2][r];
                const cplx pA = sm.full[0][r + 1], pC = sm.full[1][r + 1], pD = sm.full[2][r + 1];
                const PcrPivot pv = pcr_pivot(oA, eC, oC, pA);
                A = -pcr_scale(oA * eA, pv);
                C = -pcr_scale(oC * pC, pv);
                D = pcr_scale(nmsub(nmsub(oD, oA, eD), oC, pD), pv);
                if (CLUSTER > 1) V = pcr_scale(sm.full[3][r], pv);  // (the even rows have no V)
            }
            sts_volatile(hb0, A), sts_volatile(hb0 + kArr, C), sts_volatile(hb0 + 2 * kArr, D);
            if (CLUSTER > 1) sts_volatile(hb0 + 3 * kArr, V);
            if (sys + sys_stride < batch) named_bar_arrive(kBarFullFree, kEBThreads + kIThreads);  // `full` consumed
            named_bar_sync(kBarI, kIThreads);
            FH_TRP(8 + (j >> 5), 2);
            int cur = 0;
#pragma unroll 1
            for (int lv = 0, sb = (int)sizeof(cplx); lv < levels; ++lv, sb <<= 1) {  // rows beyond either end: zero rows
                const uint32_t rd = cur ? hb1 : hb0, wr = cur ? hb0 : hb1;
                const cplx Cm = lds_volatile(rd + kArr - sb), Ap = lds_volatile(rd + sb);
                const cplx Am = lds_volatile(rd - sb), Cp = lds_volatile(rd + kArr + sb);
                const cplx Dm = lds_volatile(rd + 2 * kArr - sb), Dp = lds_volatile(rd + 2 * kArr + sb);
                const PcrPivot pv = pcr_pivot(A, Cm, C, Ap);
                const cplx nD = pcr_scale(nmsub(nmsub(D, A, Dm), C, Dp), pv);
                if (CLUSTER > 1) {
                    const cplx Vm = lds_volatile(rd + 3 * kArr - sb), Vp = lds_volatile(rd + 3 * kArr + sb);
                    V = pcr_scale(nmsub(nmsub(V, A, Vm), C, Vp), pv);
                    sts_volatile(wr + 3 * kArr, V);
                }
                A = -pcr_scale(A * Am, pv);
                C = -pcr_scale(C * Cp, pv);
                D = nD;
                sts_volatile(wr, A), sts_volatile(wr + kArr, C), sts_volatile(wr + 2 * kArr, D);
                cur ^= 1;
                named_bar_sync(kBarI, kIThreads);
            }
            FH_TRP(8 + (j >> 5), 3);
            // now x(row 2j+1) = D - V * x_partner
            cplx x_partner = czero();
            if (CLUSTER > 1) {
                if (j == kH - 1) {
                    const uint32_t rbar = map_to_rank(&sm.xchg_bar[parity], rank ^ 1);
                    st_async_remote(map_to_rank(&sm.xchg[parity][0], rank ^ 1), D, rbar);
                    st_async_remote(map_to_rank(&sm.xchg[parity][1], rank ^ 1), V, rbar);
                }
                const cplx Ym = sm.half[cur][kHalfPad + 2 * kPcrArr + kH - 1], Vm = sm.half[cur][kHalfPad + 3 * kPcrArr + kH - 1];
                mbar_wait(&sm.xchg_bar[parity], (it >> 1) & 1);
                const cplx Yp = sm.xchg[parity][0], Vp = sm.xchg[parity][1];
                if (has_cut) {
                    // x_0 = Y0 - V0 x_cut, x_1 = Y1 - V1 x_cut (subscript = cluster rank); same operand order on both
                    // CTAs, so both obtain the same bits
                    const cplx Y0 = rank == 0 ? Ym : Yp, V0 = rank == 0 ? Vm : Vp;
                    const cplx Y1 = rank == 0 ? Yp : Ym, V1 = rank == 0 ? Vp : Vm;
                    const cplx den = nmsub(nmsub(cb, cl, V0), ch, V1);
                    x_partner = nmsub(nmsub(cd, cl, Y0), ch, Y1) * frecip(den);
                    if (rank == 0 && j == kH - 1) {
                        if constexpr (Source::kRefine) red_add(u + matrix_of(sys) * n + cut_row, x_partner);
                        else stg_stream(u + sys * n + cut_row, x_partner);
                    }
                } else {
                    const cplx x_me = nmsub(Ym, Vm, Yp) * frecip(nmsub(cone(), Vm, Vp));
                    x_partner = nmsub(Yp, Vp, x_me);
                }
            }
            FH_TRP(8 + (j >> 5), 4);
            const cplx xo = nmsub(D, V, x_partner);
            sm.xi[parity][1 + r] = xo;
            if (j == 0) sm.xi[parity][kT + 1] = x_partner;
            named_bar_sync(kBarI, kIThreads);
            const cplx xlo = sm.xi[parity][r - 1];  // interface unknown of row r - 2 (entry 0 is a zero)
            sm.xi[parity][r] = nmsub(nmsub(eD, eA, xlo), eC, xo);
            named_bar_arrive(parity ? kBarX1 : kBarX0, kEBThreads + kIThreads);
            if constexpr (Source::kStoreBands) {
                if (j == 0) *(volatile uint32_t *)&sm.i_done = it + 1;
            }
            FH_TRP(8 + (j >> 5), 5);
        }
    } else {
        // =============================================================================================== S
        if constexpr (Source::kStoreBands) {
            reg_dec<kRegsS1>();
            // A is kept: on a uniform mesh every interior row of a system holds the same four numbers, so what is
            // written need not come from the rows the other roles eliminate.  The S warps hold the interior row in
            // registers and stream it over the CTA's share of the four band arrays with plain 16-byte stores (a warp
            // covers 512 contiguous bytes per instruction); the thread whose store covered mesh row 0 / N then overwrites
            // it with that row's own values (same thread, same address: program order).  No shared memory, no barrier, no
            // dependence on the other roles: the stores run ahead of the solves and fill the write bandwidth the
            // latency-bound solver phases leave idle.  (Round 2 first sent tiles of identical rows by TMA stores: they
            // queued in front of the solution store in the SM's one bulk-copy queue, and the E/B warps waited ~5,000
            // cycles per system for their solution tile to come free.  A second attempt paced the queue -- one warp
            // filling [row 0 | 512 interior rows | row N] tiles and sending contiguous 8 KB bulk copies, at most 3 .. 15
            // of them pending: the interface warps, whose shared-memory traffic no longer shared the LSU with 1,024 store
            // instructions per system, ran their PCR levels in 480 instead of 800 cycles, but the bulk-copy engine of an
            // SM moved no more than ~18 B/cycle: 4.05 ms against 3.68, whatever the queue depth; two bands each way
            // 3.76 ms.)  Values = fh_assemble_1d_c128's, bit for bit.
            const int ts = t - (kEBThreads + kIThreads);
            const HostMap hm = host_map(n, CLUSTER);
            const int64_t lo = rank == 0 ? 0 : hm.n_loc[0] + hm.cut;                 // first row this CTA writes ...
            const int64_t cnt = rank == 0 ? hm.n_loc[0] + hm.cut : n - lo;           // ... and how many (rank 0: + cut row)
            uint32_t it = 0;
            for (int64_t sys = first_sys; sys < batch; sys += sys_stride, ++it) {
                const double k = __ldg(src.ks + sys), k2 = __ldg(src.k2s + sys);
                cplx il, id, iu, ib;
                row_values(src.p, sm.bgeom[2], 1, k, k2, il, id, iu, ib);
                FH_TRP(12 + (ts >> 5), 0);
#ifndef FH_S_AHEAD
#define FH_S_AHEAD 2
#endif
                if (FH_S_AHEAD >= 0) {  // pacing: at most FH_S_AHEAD systems ahead of the interface solves
                    if ((ts & 31) == 0)
                        while ((int)(it - *(volatile uint32_t *)&sm.i_done) > FH_S_AHEAD) __nanosleep(100);
                    __syncwarp();
                }
                const int64_t o0 = sys * n + lo, oe = o0 + cnt;
#pragma unroll 4
                for (int64_t o = o0 + ts; o < oe; o += kSThreads) {
                    stg_stream(src.out_dl + o, il), stg_stream(src.out_d + o, id);
                    stg_stream(src.out_du + o, iu), stg_stream(src.out_b + o, ib);
                }
#pragma unroll 1
                for (int e = 0; e < 2; ++e) {  // the two mesh-boundary rows
                    const int64_t g = e == 0 ? 0 : n - 1;
                    if (g >= lo && g < lo + cnt && (int)((g - lo) % kSThreads) == ts) {
                        cplx vl, vd, vu, vb;
                        row_values(src.p, sm.bgeom[e], g, k, k2, vl, vd, vu, vb);
                        const int64_t o = sys * n + g;
                        stg_stream(src.out_dl + o, vl), stg_stream(src.out_d + o, vd);
                        stg_stream(src.out_du + o, vu), stg_stream(src.out_b + o, vb);
                    }
                }
                FH_TRP(12 + (ts >> 5), 4);
            }
        }
    }
    if (CLUSTER > 1) cg::this_cluster().sync();  // no CTA leaves while its partner may still address it
    else __syncthreads();
    if (warp == 0) tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace fh
static_assert(sizeof(fh::PipeSmem) + 1024 <= 232448, "PipeSmem exceeds the 227 KB a CTA can have on sm_100");
