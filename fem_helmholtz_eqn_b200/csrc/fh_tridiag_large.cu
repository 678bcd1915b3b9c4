// One very large complex128 tridiagonal system (BASELINE configs 2, 4, 5: N = 1e6 .. 2^27) and the
// per-rank pieces of the multi-GPU SPIKE-style partitioned solve.
//
// Recursive partition, two rows per chunk:
//   reduce   : every thread owns a chunk of 16 consecutive rows and eliminates its 14 interior unknowns
//              with one downward and one upward sweep (O(1) registers), leaving two interface rows in
//              the chunk's FIRST and LAST unknown.  The interface rows of all chunks, in order, form a
//              tridiagonal system 8x smaller.  No chunk needs data of another chunk, so the same code
//              works across CTA, level and GPU boundaries.
//   coarsest : <= 4112 rows -> the batched cluster kernel (fh_tridiag_batched.cu); in the multi-GPU
//              path every rank reduces its slab down to exactly 2 rows, the 2P rows are all-gathered
//              (128 B per rank, NCCL) and solved redundantly by a tiny pivoted kernel.
//   back-sub : with its first and last unknown known, every chunk solves its 14 interior rows by a
//              local Thomas sweep and writes x.
// A CTA stages a tile of 64 chunks (1024 rows, 64 KB) in shared memory with coalesced 16-byte cp.async;
// the slot index is XOR-swizzled so that the per-thread sweeps (stride 256 B) are bank-conflict free.
// Three CTAs are resident per SM, so loads of one tile overlap the sweeps of the others.
//
// HBM traffic per row (level 0): reduce 64 B read; back-sub 64 B read + 16 B written; interface rows
// add 2*(64+64+16)/16 = 18 B at level 1 and 1/8 of that per further level.  The fused Helmholtz
// variants generate level-0 rows from (h, k) instead of reading them: 16 B written per row + interface.
#include "fh_helm1d.cuh"

namespace fh {

int batched_solve_bands(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n,
                        int64_t batch, int32_t *status, cudaStream_t stream);  // fh_tridiag_batched.cu
constexpr int64_t kCoarseMax = 4112;                                           // == kBatchedMaxN

constexpr int kLT = 64;              // threads (= chunks) per CTA
constexpr int kLM = 16;              // rows per chunk
constexpr int kLTile = kLT * kLM;    // 1024 rows per tile
constexpr int kMaxLevels = 16;

__device__ __forceinline__ int lswz(int r) { return r ^ ((r >> 4) & 7); }

// A-posteriori accuracy probe, the large-system counterpart of the batched kernel's (fh_tridiag_dev.cuh): the chunked
// elimination is unpivoted; what it can lose shows up as a defect of a chunk's two INTERFACE equations (first and last
// row of the chunk, which became rows of the next level) when they are re-evaluated with the computed solution --
// interior rows are satisfied by construction of the back-substitution.  A defect of a level-l row is a residual of the
// ORIGINAL system at that row (the reduced row is the original one minus multiples of interior rows, which hold), so it
// is measured in units of the original system: |r| <= tau (||A||_inf max|x_neighbourhood| + |d|) on every probed row
// bounds the normwise backward error by ~2 tau.  (Measured against the reduced row's own entries -- which shrink by 16x
// per level on a fine mesh -- the same defects look up to 1e6 times larger than they are: tools/probe_large_stats.py.)
// ||A||_inf: the level-0 reduction records the largest row sum it sees; generated rows take the interior row's.
// Status bits as fh_tridiag_solve_batched_c128: bit 1 (2) defect > kLargeProbeTol -> refine; bit 0 (1) as well above
// kLargeSevereTol or if not finite.
constexpr double kLargeProbeTol = 5e-14, kLargeSevereTol = 1e-8;
__device__ __forceinline__ double labs1(cplx z) { return fabs(z.re) + fabs(z.im); }
#ifdef FH_PROBE_STAT  // development: the largest relative defect any probe has seen (tools/probe_large_stats.py)
__device__ unsigned long long g_probe_max = 0ull;
#endif
__device__ __forceinline__ int probe_row(cplx a, cplx b, cplx c, cplx d, cplx xm, cplx x0, cplx xp, double anorm) {
    const cplx r = nmsub(nmsub(nmsub(d, a, xm), b, x0), c, xp);
    const double scale = anorm * fmax(fmax(labs1(xm), labs1(x0)), labs1(xp)) + labs1(d);
    const double v = labs1(r);
#ifdef FH_PROBE_STAT
    if (scale > 0.0 && v == v) atomicMax(&g_probe_max, (unsigned long long)__double_as_longlong(v / scale));
#endif
    return v > kLargeSevereTol * scale ? 3 : (v > kLargeProbeTol * scale ? 2 : 0);  // (NaN: reported by the finite test)
}

struct LargeSmem {  // +8: a trailing 1-row tile is merged into its predecessor (1025 rows)
    cplx a[kLTile + 8], b[kLTile + 8], c[kLTile + 8], d[kLTile + 8];
};

// Tiling rule shared by host and device (and by tests/models/recursive_partition_model.py): tiles of 1024
// rows, chunks of 16 rows; a trailing 1-row tile / chunk is merged into its predecessor so that every
// chunk has at least 2 rows (its first and last unknown are distinct).
__host__ __device__ __forceinline__ int64_t tile_count(int64_t n) {
    int64_t t = (n + kLTile - 1) / kLTile;
    if (n % kLTile == 1 && n > kLTile) --t;
    return t;
}
__host__ __device__ __forceinline__ int tile_rows(int64_t n, int64_t n_tiles, int64_t tile) {
    return tile == n_tiles - 1 ? (int)(n - tile * kLTile) : kLTile;
}
__host__ __device__ __forceinline__ int chunk_count(int rows) {
    int c = (rows + kLM - 1) / kLM;
    if (rows % kLM == 1 && rows > kLM) --c;
    return c;
}
__host__ __device__ __forceinline__ int64_t reduced_rows(int64_t n) {
    const int64_t nt = tile_count(n);
    return 2 * ((nt - 1) * kLT + chunk_count(tile_rows(n, nt, nt - 1)));
}

// ---------------------------------------------------------------------------------------------- sources
struct GlobalBands {  // rows of a level stored in global memory
    const cplx *dl, *d, *du, *b;
    int64_t n;
    int open_first, open_last;  // keep the coupling of the first / last row (slab of a larger system)
    __device__ __forceinline__ bool open_first_row() const { return open_first != 0; }
    __device__ __forceinline__ bool open_last_row() const { return open_last != 0; }
    __device__ __forceinline__ void fill(LargeSmem &sm, int64_t row0, int rows) const {
        for (int j = threadIdx.x; j < rows; j += kLT) {
            const int s = lswz(j);
            cp_async16(&sm.a[s], dl + row0 + j);
            cp_async16(&sm.b[s], d + row0 + j);
            cp_async16(&sm.c[s], du + row0 + j);
            cp_async16(&sm.d[s], b + row0 + j);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (row0 == 0 && !open_first) sm.a[lswz(0)] = czero();
            if (row0 + rows == n && !open_last) sm.c[lswz(rows - 1)] = czero();
        }
        __syncthreads();
    }
};

struct HelmRows {  // level-0 rows generated from the Helmholtz problem: local row j <-> mesh row g_off + j
    Problem1d p;
    double k, k2;
    const double *kk = nullptr;  // device-resident (k, k**2): the periodic recursion only (see UniformRows)
    int64_t g_off, n;
    __device__ __forceinline__ bool open_first_row() const { return g_off > 0; }
    __device__ __forceinline__ bool open_last_row() const { return g_off + n - 1 < p.n_elem; }
    // STORE: the generated rows are also written out (out_* = the slab's dl, d, du, b: what fh_assemble_1d_slab_c128 would
    // have stored, bit for bit), so that "assemble the slab, keep A, solve" costs 64 B per row written here instead of
    // 64 written by the assembly kernel + 64 read back by this one.
    template <bool STORE>
    __device__ __forceinline__ void fill_impl(LargeSmem &sm, int64_t row0, int rows, cplx *out_dl, cplx *out_d,
                                              cplx *out_du, cplx *out_b) const {
        // scalar-h (the reference's mode): every interior row is the same -- evaluate it once per thread
        const bool scalar = p.h_mode == FH_H_SCALAR && p.n_elem >= 2;
        double kv = k, k2v = k2;
        if (kk != nullptr) kv = __ldg(kk), k2v = __ldg(kk + 1);
        cplx il = czero(), id = czero(), iu = czero(), ib = czero();
        if (scalar) row_values(p, row_geom(p, 1), 1, kv, k2v, il, id, iu, ib);
        for (int j = threadIdx.x; j < rows; j += kLT) {
            const int64_t g = g_off + row0 + j;
            cplx lo = il, di = id, hi = iu, rhs = ib;
            if (!scalar || g == 0 || g == p.n_elem) row_values(p, row_geom(p, g), g, kv, k2v, lo, di, hi, rhs);
            const int s = lswz(j);
            sm.a[s] = lo, sm.b[s] = di, sm.c[s] = hi, sm.d[s] = rhs;
            if (STORE) {
                const int64_t o = row0 + j;
                stg_stream(out_dl + o, lo), stg_stream(out_d + o, di), stg_stream(out_du + o, hi), stg_stream(out_b + o, rhs);
            }
        }
        __syncthreads();
    }
    __device__ __forceinline__ void fill(LargeSmem &sm, int64_t row0, int rows) const {
        fill_impl<false>(sm, row0, rows, nullptr, nullptr, nullptr, nullptr);
    }
};
struct HelmRowsStore : HelmRows {  // level 0 of "assemble + keep A + solve" on one long mesh / one slab of it
    cplx *out_dl, *out_d, *out_du, *out_b;
    __device__ __forceinline__ void fill(LargeSmem &sm, int64_t row0, int rows) const {
        fill_impl<true>(sm, row0, rows, out_dl, out_d, out_du, out_b);
    }
};

// ---------------------------------------------------------------------------------------------- reduce
template <class Source>
__global__ void __launch_bounds__(kLT, 3)
large_reduce_kernel(Source src, cplx *__restrict__ ra, cplx *__restrict__ rb, cplx *__restrict__ rc,
                    cplx *__restrict__ rd, unsigned long long *__restrict__ anorm_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LargeSmem &sm = *reinterpret_cast<LargeSmem *>(smem_raw);
    const int64_t n_tiles = tile_count(src.n);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kLTile;
        const int rows = tile_rows(src.n, n_tiles, tile);
        src.fill(sm, row0, rows);
        const int t = threadIdx.x, r0 = t * kLM;
        const int nch = chunk_count(rows);
        const int len = t < nch ? (t == nch - 1 ? rows - r0 : kLM) : 0;  // 2..17 (1 only if n == 1)
        if (anorm_out != nullptr) {  // level 0: ||A||_inf of the system, the unit of the back-substitution's probe
            double rs = 0.0;
            for (int i = 0; i < len; ++i) {
                const int s = lswz(r0 + i);
                rs = fmax(rs, labs1(sm.a[s]) + labs1(sm.b[s]) + labs1(sm.c[s]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rs = fmax(rs, __shfl_xor_sync(0xffffffffu, rs, o));
            if ((t & 31) == 0) atomicMax(anorm_out, (unsigned long long)__double_as_longlong(rs));  // (>= 0: bit order)
        }
        if (len >= 1) {
            cplx a1, b1, c1, d1, a2, b2, c2, d2;  // interface rows in the chunk's first / last unknown
            if (len == 1) {
                const int s = lswz(r0);
                a1 = sm.a[s], b1 = sm.b[s], c1 = sm.c[s], d1 = sm.d[s];
                a2 = czero(), b2 = cone(), c2 = czero(), d2 = czero();  // n == 1 only: placeholder row
                c1 = czero();
            } else {
                {  // downward from row 1: fill-in couples to x_first
                    int s = lswz(r0 + 1);
                    cplx pa = sm.a[s], pb = sm.b[s], pc = sm.c[s], pd = sm.d[s];
                    for (int i = 2; i < len; ++i) {
                        s = lswz(r0 + i);
                        const cplx w = sm.a[s] * frecip(pb);
                        const cplx nb = nmsub(sm.b[s], w, pc), nd = nmsub(sm.d[s], w, pd);
                        pa = -(w * pa), pb = nb, pc = sm.c[s], pd = nd;
                    }
                    a2 = pa, b2 = pb, c2 = pc, d2 = pd;
                }
                {  // upward from row len-2: fill-in couples to x_last
                    int s = lswz(r0 + len - 2);
                    cplx qa = sm.a[s], qb = sm.b[s], qc = sm.c[s], qd = sm.d[s];
                    for (int i = len - 3; i >= 0; --i) {
                        s = lswz(r0 + i);
                        const cplx w = sm.c[s] * frecip(qb);
                        const cplx nb = nmsub(sm.b[s], w, qa), nd = nmsub(sm.d[s], w, qd);
                        qc = -(w * qc), qb = nb, qa = sm.a[s], qd = nd;
                    }
                    a1 = qa, b1 = qb, c1 = qc, d1 = qd;
                }
            }
            const int64_t o = 2 * (tile * kLT + t);
            ra[o] = a1, rb[o] = b1, rc[o] = c1, rd[o] = d1;
            ra[o + 1] = a2, rb[o + 1] = b2, rc[o + 1] = c2, rd[o + 1] = d2;
        }
        __syncthreads();  // tile buffers are refilled by the next iteration
    }
}

// ---------------------------------------------------------------------------------------------- back-substitution
template <class Source>
__global__ void __launch_bounds__(kLT, 3)
large_backsub_kernel(Source src, const cplx *__restrict__ xr, cplx *__restrict__ x, int32_t *__restrict__ status,
                     const double *__restrict__ anorm_ptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LargeSmem &sm = *reinterpret_cast<LargeSmem *>(smem_raw);
    const double anorm = *anorm_ptr;
    const int64_t n_tiles = tile_count(src.n);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kLTile;
        const int rows = tile_rows(src.n, n_tiles, tile);
        src.fill(sm, row0, rows);
        const int t = threadIdx.x, r0 = t * kLM;
        const int nch = chunk_count(rows);
        const int len = t < nch ? (t == nch - 1 ? rows - r0 : kLM) : 0;
        bool bad = false;
        int flag = 0;
        if (len >= 1) {
            const int64_t o = 2 * (tile * kLT + t);
            const cplx x0 = xr[o], xl = len > 1 ? xr[o + 1] : x0;
            // (the neighbours' unknowns, for the probe: requested together with the chunk's own, used at the end)
            const int64_t n_red = reduced_rows(src.n);
            const cplx xprev = o > 0 ? xr[o - 1] : czero(), xnext = o + 2 < n_red ? xr[o + 2] : czero();
            bad = !cfinite(x0) || !cfinite(xl);
            // the two interface equations of the chunk, kept for the probe (their right-hand sides are overwritten below)
            const int sF = lswz(r0), sL = lswz(r0 + len - 1);
            const cplx fa = sm.a[sF], fb = sm.b[sF], fc = sm.c[sF], fd = sm.d[sF];
            const cplx la = sm.a[sL], lb = sm.b[sL], lc = sm.c[sL], ld = sm.d[sL];
            // interior rows 1 .. len-2: Thomas with the two known neighbours; c', d' overwrite the tile
            cplx cp = czero(), dp = x0;  // "row 0" acts as x_0 = x0
            for (int i = 1; i <= len - 2; ++i) {
                const int s = lswz(r0 + i);
                const cplx ai = sm.a[s];
                const cplx inv = frecip(nmsub(sm.b[s], ai, cp));
                dp = nmsub(sm.d[s], ai, dp) * inv;
                cp = sm.c[s] * inv;
                sm.c[s] = cp, sm.d[s] = dp;
            }
            cplx xn = xl;
            if (len > 1) sm.d[lswz(r0 + len - 1)] = xl;
            for (int i = len - 2; i >= 1; --i) {
                const int s = lswz(r0 + i);
                xn = nmsub(sm.d[s], sm.c[s], xn);
                bad |= !cfinite(xn);
                sm.d[s] = xn;
            }
            sm.d[lswz(r0)] = x0;
            if (len >= 2) {
                const cplx x1 = len > 2 ? sm.d[lswz(r0 + 1)] : xl, xm2 = len > 2 ? sm.d[lswz(r0 + len - 2)] : x0;
                // the neighbours' unknowns: the previous chunk's last / the next chunk's first (a slab's outermost
                // rows couple to another rank's unknown, which is not here: those two equations are skipped)
                if (o > 0) flag |= probe_row(fa, fb, fc, fd, xprev, x0, x1, anorm);
                else if (!src.open_first_row()) flag |= probe_row(czero(), fb, fc, fd, czero(), x0, x1, anorm);
                if (o + 2 < n_red) flag |= probe_row(la, lb, lc, ld, xm2, xl, xnext, anorm);
                else if (!src.open_last_row()) flag |= probe_row(la, lb, czero(), ld, xm2, xl, czero(), anorm);
            }
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        for (int j = threadIdx.x; j < rows; j += kLT) stg_stream(x + row0 + j, sm.d[lswz(j)]);
        if (any_bad && status != nullptr && threadIdx.x == 0) atomicOr(status, 1);
        if (flag != 0 && status != nullptr) atomicOr(status, flag);  // (rare)
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- uniform meshes
// A GENERATED system on a uniform mesh (scalar h: the reference's only mode, nb:158) is periodic at every level of
// the recursion.  Level 0: every interior row holds the same four numbers, so every full interior chunk reduces to
// the same two interface rows (E1, O1).  Level 1 therefore reads E1 O1 E1 O1 ..., every full chunk of THAT reduces to
// the same (E2, O2), and so on: at level l the "plain" rows alternate between two constant rows (E_l at even, O_l at
// odd positions; E_0 = O_0).  Only the chunks that touch a mesh end or the ragged end of a level ("special" chunks: at
// most one at the front and two or three at the back of each level) hold anything else.  Hence
//   reduce   : nothing is streamed.  ONE single-CTA kernel walks down the levels; per level one lane reduces the plain
//              chunk (-> E_{l+1}, O_{l+1}), one lane per special chunk reduces that chunk (rows = constants, generated
//              mesh-end rows, or special rows of the level, which are the only entries of the level arrays ever written),
//              and one lane carries out the chunk's Thomas sweep symbolically in its two end values:
//                  x_i = alpha_i x_first + beta_i x_last + gamma_i,      i = 0 .. 15;
//   back-sub : plain chunks are a streaming map of the two end values through (alpha, beta, gamma) -- 2 B read + 16 B
//              written per row --, special chunks take the generic sweep.  Levels of <= 2^14 rows are handled inside
//              one single-CTA kernel, the larger ones take one launch each.
// The elimination arithmetic is that of large_reduce_kernel / large_backsub_kernel; it is just not repeated 2^23 times.
struct UniformRows {  // what HelmRows provides
    Problem1d p;
    double k, k2;
    const double *kk;  // != NULL: (k, k**2) are read from device memory when the kernel starts (CUDA-graph replays)
    int64_t g_off, n;
    __device__ __forceinline__ void fetch_k() {
        if (kk != nullptr) k = kk[0], k2 = kk[1];
    }
    __device__ __forceinline__ void interior(cplx &lo, cplx &di, cplx &hi, cplx &rhs) const {
        row_values(p, row_geom(p, 1), 1, k, k2, lo, di, hi, rhs);
    }
    __device__ __forceinline__ void row(int64_t g, const cplx in[4], cplx &lo, cplx &di, cplx &hi, cplx &rhs) const {
        lo = in[0], di = in[1], hi = in[2], rhs = in[3];
        if (g == 0 || g == p.n_elem) row_values(p, row_geom(p, g), g, k, k2, lo, di, hi, rhs);
    }
};

struct ChunkPos {
    int64_t r0;  // first row of the chunk inside the level
    int len;     // rows (2 .. 17; 1 only if the level has a single row)
    bool exists;
};
__host__ __device__ __forceinline__ ChunkPos chunk_pos(int64_t n, int64_t n_tiles, int64_t c) {
    const int64_t tile = c / kLT;
    const int t = (int)(c - tile * kLT);
    const int rows = tile_rows(n, n_tiles, tile);
    const int nch = chunk_count(rows);
    ChunkPos cp;
    cp.exists = t < nch;
    cp.r0 = tile * kLTile + (int64_t)t * kLM;
    cp.len = cp.exists ? (t == nch - 1 ? rows - t * kLM : kLM) : 0;
    return cp;
}

// the two interface rows of a chunk whose rows come from `rowfn(i, a, b, c, d)` -- the arithmetic of large_reduce_kernel,
// one sweep per call (the two sweeps are independent: they run in different warps)
template <class RowFn>
__device__ __forceinline__ void reduce_chunk_down(RowFn rowfn, int len, cplx out2[4]) {  // -> row in the LAST unknown
    cplx ra, rb, rc, rd, pa, pb, pc, pd;
    rowfn(1, pa, pb, pc, pd);
    for (int i = 2; i < len; ++i) {  // fill-in couples to x_first
        rowfn(i, ra, rb, rc, rd);
        const cplx w = ra * frecip(pb);
        const cplx nb = nmsub(rb, w, pc), nd = nmsub(rd, w, pd);
        pa = -(w * pa), pb = nb, pc = rc, pd = nd;
    }
    out2[0] = pa, out2[1] = pb, out2[2] = pc, out2[3] = pd;
}
template <class RowFn>
__device__ __forceinline__ void reduce_chunk_up(RowFn rowfn, int len, cplx out1[4]) {  // -> row in the FIRST unknown
    cplx ra, rb, rc, rd, qa, qb, qc, qd;
    rowfn(len - 2, qa, qb, qc, qd);
    for (int i = len - 3; i >= 0; --i) {  // fill-in couples to x_last
        rowfn(i, ra, rb, rc, rd);
        const cplx w = rc * frecip(qb);
        const cplx nb = nmsub(rb, w, qa), nd = nmsub(rd, w, qd);
        qc = -(w * qc), qb = nb, qa = ra, qd = nd;
    }
    out1[0] = qa, out1[1] = qb, out1[2] = qc, out1[3] = qd;
}

struct LevelPtrs {
    cplx *a, *b, *c, *d, *x;
};

#ifdef FH_DEEP_TRACE
__device__ long long g_deep_trace[64];
#define FH_DT(k) do { if (threadIdx.x == 0) g_deep_trace[k] = clock64(); } while (0)
#else
#define FH_DT(k) do { } while (0)
#endif
constexpr int kDeepThreads = 256;
__device__ __forceinline__ void named_bar_sync_large(int id, int threads) {  // hardware barrier `id` over `threads` threads
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
constexpr int kDeepSpecialMax = 2;        // special chunks per level: one at the front, one at the back (tests/test_large_plan_cpu.py)
constexpr int kDeepSmallLevelsMax = 8;    // levels the single-CTA back-substitution stages at once (2^14 rows down to 4: five)
constexpr int64_t kDeepSmall = 1 << 14;   // levels up to this many rows are back-substituted inside the single-CTA kernel

struct DeepLevel {
    int64_t n, chunks;  // rows, chunks
    int64_t plo, phi;   // plain rows  [plo, phi): row r holds E (r even) / O (r odd)
    int64_t clo, chi;   // plain chunks [clo, chi): 16 plain rows each
};
struct DeepPlan {  // host-computed geometry of the recursion + where the level arrays live (kernel parameter)
    int levels;                          // level `levels` has exactly 2 rows
    DeepLevel lv[kMaxLevels + 1];
    LevelPtrs ptr[kMaxLevels + 1];       // levels >= 1 (ptr[0] is unused: level-0 rows are generated, its x is the caller's)
};
static_assert(sizeof(UniformRows) + sizeof(DeepPlan) + 64 <= 4096, "the plan travels in kernel parameter space");
struct DeepConsts {  // per level, written by the reduce kernel, read by the back-substitution
    cplx eo[2][4];        // E, O: (a, b, c, d)
    cplx tab[3][kLM];     // alpha, beta, gamma
};

__device__ __forceinline__ int deep_special_count(const DeepLevel &L) { return (int)(L.clo + (L.chunks - L.chi)); }
__device__ __forceinline__ int64_t deep_special_chunk(const DeepLevel &L, int s) {
    return s < L.clo ? (int64_t)s : L.chi + (s - L.clo);
}
// What one level of the single-CTA reduction hands to the next, kept in shared memory (no global round trip between
// levels): the level's two constant rows and its special rows -- the interface rows (first, last) of the special chunks
// of the level above, in order.
struct DeepCarry {
    cplx eo[2][4];
    cplx spec[2 * kDeepSpecialMax][4];
};
// row r of level l; carry != NULL: the level's constants and special rows come from shared memory
__device__ __forceinline__ void deep_row(const UniformRows &u, const DeepPlan &dp, const DeepConsts *cs,
                                         const DeepCarry *carry, int l, int64_t r, cplx &a, cplx &b, cplx &c, cplx &d) {
    const DeepLevel &L = dp.lv[l];
    const cplx(*eo)[4] = carry != nullptr ? carry->eo : cs[l].eo;
    if (r >= L.plo && r < L.phi) {
        const cplx *e = eo[r & 1];
        a = e[0], b = e[1], c = e[2], d = e[3];
    } else if (l == 0) {
        u.row(u.g_off + r, eo[0], a, b, c, d);
    } else if (carry != nullptr) {
        const DeepLevel &P = dp.lv[l - 1];  // row r is an interface row of chunk r / 2 of the level above
        const int64_t pc = r >> 1;
        const int ps = (int)(pc < P.clo ? pc : P.clo + (pc - P.chi));
        const cplx *e = carry->spec[2 * ps + (int)(r & 1)];
        a = e[0], b = e[1], c = e[2], d = e[3];
    } else {
        const LevelPtrs &p = dp.ptr[l];
        a = p.a[r], b = p.b[r], c = p.c[r], d = p.d[r];
    }
}

typedef cplx DeepRows[kLM + 1][4];  // the (a, b, c, d) rows of one chunk, staged in shared memory
struct DeepStage {
    DeepRows rows[kDeepSpecialMax + 1];
    int len[kDeepSpecialMax + 1];
};
// All threads: the rows of the special chunks of level l -- and, with_plain, of a plain chunk as entry number ns -- go to
// shared memory with every load in flight at once (inside the sweeps each row would cost a dependent L2 round trip).
// first_thread: the thread that takes entry 0 (several levels staged back to back use different threads).
__device__ __forceinline__ void deep_stage(const UniformRows &u, const DeepPlan &dp, const DeepConsts *cs,
                                           const DeepCarry *carry, int l, bool with_plain, DeepStage &sg,
                                           int first_thread = 0, int tid = -1, int nthreads = 0) {
    const DeepLevel &L = dp.lv[l];
    const int ns = deep_special_count(L);
    const int total = (ns + (with_plain ? 1 : 0)) * (kLM + 1);
    const int nt = nthreads > 0 ? nthreads : (int)blockDim.x;  // (a subset of the CTA's threads: tid in [0, nthreads))
    if (tid < 0) tid = (int)threadIdx.x;
    for (int idx = (tid + nt - first_thread % nt) % nt; idx < total; idx += nt) {
        const int s = idx / (kLM + 1), i = idx - s * (kLM + 1);
        cplx a, b, c, d;
        if (s == ns) {
            if (i == 0) sg.len[s] = kLM;
            if (i >= kLM) continue;
            const cplx *e = carry != nullptr ? carry->eo[i & 1] : cs[l].eo[i & 1];
            a = e[0], b = e[1], c = e[2], d = e[3];
        } else {
            const ChunkPos cp = chunk_pos(L.n, tile_count(L.n), deep_special_chunk(L, s));
            if (i == 0) sg.len[s] = cp.len;
            if (i >= cp.len) continue;
            deep_row(u, dp, cs, carry, l, cp.r0 + i, a, b, c, d);
        }
        sg.rows[s][i][0] = a, sg.rows[s][i][1] = b, sg.rows[s][i][2] = c, sg.rows[s][i][3] = d;
    }
}

// All reductions, level 0 down to the 2-row level, by ONE CTA.  Warp 0 runs the downward sweeps (lane s < ns: special
// chunk s, lane ns: the plain chunk), warp 1 the upward sweeps, lane 0 of warp 2 the symbolic Thomas sweep.
// cs: per-level constants in global memory (visible to the whole CTA after each level's barrier).
__device__ void deep_reduce_levels(const UniformRows &u, const DeepPlan &dp, DeepConsts *cs, DeepStage &sg,
                                   DeepCarry carry[2]) {
    // The symbolic Thomas sweep of a level's plain chunk (forward + backward: 28 dependent steps, ~3,000 cycles) is
    // longer than the two reductions (14 steps, ~2,100) and nothing below a level depends on its tables: it runs on
    // its own warp, one lane, DECOUPLED from the level barriers -- it starts level l as soon as the constants (E_l, O_l)
    // are published and may trail the reductions by a level or two; two such warps take the even and the odd levels (the
    // round-2 start had the sweep inside the level barrier: 2.4 us per level, 19.7 us for the eight levels of a 2^24-row
    // slab).
    __shared__ int levels_published;  // cs[l].eo is valid for l <= levels_published
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    constexpr int kTableWarp = 2, kTableWarps = 2, kOthers = kDeepThreads - 32 * kTableWarps, kBarLevels = 1;
    if (t == 0) {
        cplx in[4];
        u.interior(in[0], in[1], in[2], in[3]);
        for (int q = 0; q < 4; ++q) cs[0].eo[0][q] = cs[0].eo[1][q] = carry[0].eo[0][q] = carry[0].eo[1][q] = in[q];
        levels_published = 0;
    }
    __syncthreads();
    if (warp >= kTableWarp && warp < kTableWarp + kTableWarps) {
        if (lane == 0) {
            for (int l = warp - kTableWarp; l < dp.levels; l += kTableWarps) {
                while (*(volatile int *)&levels_published < l) __nanosleep(40);
                __threadfence_block();
                cplx R[2][4];  // the plain rows of the level: E (even positions), O (odd positions)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const volatile double *e0 = &cs[l].eo[0][q].re, *e1 = &cs[l].eo[1][q].re;
                    R[0][q] = mk(e0[0], e0[1]), R[1][q] = mk(e1[0], e1[1]);
                }
                // x_i = alpha_i x_first + beta_i x_last + gamma_i for a plain chunk of this level
                cplx fc[kLM], fP[kLM], fG[kLM];
                cplx cp = czero(), Pp = cone(), Gp = czero();
#pragma unroll
                for (int i = 1; i <= kLM - 2; ++i) {  // forward sweep with x_first as a symbol: dp_i = P_i x_first + G_i
                    const cplx *r = R[i & 1];
                    const cplx inv = frecip(nmsub(r[1], r[0], cp));
                    Pp = -((r[0] * Pp) * inv);
                    Gp = nmsub(r[3], r[0], Gp) * inv;
                    cp = r[2] * inv;
                    fc[i] = cp, fP[i] = Pp, fG[i] = Gp;
                }
                cplx al = czero(), be = cone(), ga = czero();  // backward sweep with x_last as a symbol
                cs[l].tab[0][kLM - 1] = al, cs[l].tab[1][kLM - 1] = be, cs[l].tab[2][kLM - 1] = ga;  // x_15 = x_last
#pragma unroll
                for (int i = kLM - 2; i >= 1; --i) {
                    al = nmsub(fP[i], fc[i], al);
                    be = -(fc[i] * be);
                    ga = nmsub(fG[i], fc[i], ga);
                    cs[l].tab[0][i] = al, cs[l].tab[1][i] = be, cs[l].tab[2][i] = ga;
                }
                cs[l].tab[0][0] = cone(), cs[l].tab[1][0] = czero(), cs[l].tab[2][0] = czero();  // x_0 = x_first
            }
        }
    } else {
        const int tid = t < 32 * kTableWarp ? t : t - 32 * kTableWarps;  // the other 192 threads, numbered 0 .. 191
        for (int l = 0; l < dp.levels; ++l) {
            const DeepLevel &L = dp.lv[l];
            const int ns = deep_special_count(L);
            const DeepCarry &now = carry[l & 1];
            DeepCarry &next = carry[(l + 1) & 1];
            FH_DT(3 * l);
            deep_stage(u, dp, cs, &now, l, true, sg, 0, tid, kOthers);
            named_bar_sync_large(kBarLevels, kOthers);
            FH_DT(3 * l + 1);
            if (warp < 2 && lane <= ns) {
                cplx out[4];
                const DeepRows &R = sg.rows[lane];
                auto rowfn = [&](int i, cplx &a, cplx &b, cplx &c, cplx &d) { a = R[i][0], b = R[i][1], c = R[i][2], d = R[i][3]; };
                if (warp == 0) reduce_chunk_down(rowfn, sg.len[lane], out);
                else reduce_chunk_up(rowfn, sg.len[lane], out);
                const int which = warp == 0 ? 1 : 0;  // downward sweep -> the row in the LAST unknown
                if (lane == ns) {
                    for (int q = 0; q < 4; ++q) cs[l + 1].eo[which][q] = next.eo[which][q] = out[q];
                } else {  // (to global memory for the back-substitution, to shared memory for the next level)
                    const LevelPtrs &p = dp.ptr[l + 1];
                    const int64_t o = 2 * deep_special_chunk(L, lane) + which;
                    p.a[o] = out[0], p.b[o] = out[1], p.c[o] = out[2], p.d[o] = out[3];
                    for (int q = 0; q < 4; ++q) next.spec[2 * lane + which][q] = out[q];
                }
                __threadfence_block();
            }
            named_bar_sync_large(kBarLevels, kOthers);
            FH_DT(3 * l + 2);
            if (t == 0) *(volatile int *)&levels_published = l + 1;  // (E, O) of level l + 1 are in cs[l + 1].eo
        }
    }
    FH_DT(60);
    __syncthreads();  // the tables of every level are complete
    FH_DT(61);
    // the 2-row level in full: plain positions are filled in, so that its four arrays (a[2] b[2] c[2] d[2], contiguous)
    // are the slab's interface rows
    const DeepLevel &T = dp.lv[dp.levels];
    if (t < T.n && t >= T.plo && t < T.phi) {
        const LevelPtrs &p = dp.ptr[dp.levels];
        const cplx *e = cs[dp.levels].eo[t & 1];
        p.a[t] = e[0], p.b[t] = e[1], p.c[t] = e[2], p.d[t] = e[3];
    }
    __syncthreads();
}

// Plain chunks [c_begin, c_end) of one level: x(16 c + i) = alpha_i xr(2 c) + beta_i xr(2 c + 1) + gamma_i.  256 threads
// = 16 chunks per pass, 8 passes in flight (the end-value loads overlap): one "round" writes 128 chunks = 32 KB of
// contiguous solution.  A thread keeps the same row-in-chunk i throughout.
// (xr is NOT __restrict__ / read-only-cached: inside the single-CTA kernel it was written a level earlier)
// Probe of a plain chunk's two interface equations (rows 16 c and 16 c + 15 of the level: the constants E and O), by the
// two threads that hold x(16 c + 1) and x(16 c + 14); the neighbouring chunks' unknowns are entries 2 c - 1 and 2 c + 2
// of the level below (n_next entries), passed in by the caller.
__device__ __forceinline__ int deep_probe_plain(const DeepConsts &C, double anorm, int i, cplx xi, cplx x_first,
                                                cplx x_last, bool has_prev, cplx x_prev, bool has_next, cplx x_next) {
    if (i == 1 && has_prev) {
        const cplx *e = C.eo[0];
        return probe_row(e[0], e[1], e[2], e[3], x_prev, x_first, xi, anorm);
    }
    if (i == kLM - 2 && has_next) {
        const cplx *e = C.eo[1];
        return probe_row(e[0], e[1], e[2], e[3], xi, x_last, x_next, anorm);
    }
    return 0;
}
// ||A||_inf of a generated system: the interior row's sum (the mesh-end rows hold half its diagonal)
__device__ __forceinline__ double deep_anorm(const DeepConsts *cs) {
    const cplx *e = cs[0].eo[0];
    return labs1(e[0]) + labs1(e[1]) + labs1(e[2]);
}
__device__ __forceinline__ int deep_backsub_plain(const DeepConsts &C, int64_t c_begin, int64_t c_end, const cplx *xr,
                                                  cplx *x) {
    const int i = threadIdx.x & (kLM - 1);
    const cplx al = C.tab[0][i], be = C.tab[1][i], ga = C.tab[2][i];
    int flag = 0;
    constexpr int kStep = kDeepThreads / kLM, kFly = 8;
    for (int64_t c0 = c_begin + (threadIdx.x >> 4); c0 < c_end; c0 += kFly * kStep) {
        cplx x0[kFly], xl[kFly];
#pragma unroll
        for (int q = 0; q < kFly; ++q) {
            const int64_t c = c0 + q * kStep;
            if (c < c_end) x0[q] = xr[2 * c], xl[q] = xr[2 * c + 1];  // (16 lanes read the same pair: one request)
        }
#pragma unroll
        for (int q = 0; q < kFly; ++q) {
            const int64_t c = c0 + q * kStep;
            if (c < c_end) {
                const cplx xi = cmadd(cmadd(ga, al, x0[q]), be, xl[q]);
                flag |= cfinite(xi) ? 0 : 1;
                stg_stream(x + c * kLM + i, xi);
            }
        }
    }
    return flag;
}
// The probe of the plain chunks.  They all go through the same three tables and the same two constant rows; what
// differs from chunk to chunk is only the pair of end values.  A loss of accuracy in the tables or constants of a level
// -- the failure mode of the unpivoted symbolic sweep -- therefore shows in every chunk, and a SAMPLE of them, spread
// evenly over the level, is probed: one chunk per thread of one CTA (<= 256 chunks), both interface equations, the two
// unknowns next to the end values re-evaluated from the tables.  (Probing every chunk inside the streaming loop made the
// store stream instruction-bound: 2^27 rows took 1.2 ms instead of 0.5 ms.)
__device__ __forceinline__ int deep_probe_samples(const DeepConsts &C, double anorm, const DeepLevel &L, const cplx *xr,
                                                  int64_t n_next, int tid, int n_threads) {  // sample `tid` of `n_threads`
    const int64_t count = L.chi - L.clo;
    if (count <= 0 || tid < 0) return 0;
    const int64_t stride = count > (int64_t)n_threads ? count / (int64_t)n_threads : 1;
    const int64_t c = L.clo + (int64_t)tid * stride;
    if (c >= L.chi) return 0;
    const cplx xf = xr[2 * c], xl = xr[2 * c + 1];
    int flag = 0;
    if (c > 0) {
        const cplx x1 = cmadd(cmadd(C.tab[2][1], C.tab[0][1], xf), C.tab[1][1], xl);
        flag |= deep_probe_plain(C, anorm, 1, x1, xf, xl, true, xr[2 * c - 1], false, czero());
    }
    if (2 * c + 2 < n_next) {
        const cplx x14 = cmadd(cmadd(C.tab[2][kLM - 2], C.tab[0][kLM - 2], xf), C.tab[1][kLM - 2], xl);
        flag |= deep_probe_plain(C, anorm, kLM - 2, x14, xf, xl, false, czero(), true, xr[2 * c + 2]);
    }
    return flag;
}
// special chunk number s of level l (rows staged by deep_stage): Thomas sweep between the two known end values, one
// thread; c', d' overwrite the staged rows
__device__ __forceinline__ int deep_backsub_special(const DeepLevel &L, double anorm, DeepStage &sg, int s,
                                                    const cplx *xr, int64_t n_next, cplx *x) {
    const int64_t c = deep_special_chunk(L, s), r0 = c * kLM;
    const int len = sg.len[s];
    DeepRows &R = sg.rows[s];
    const cplx x0 = xr[2 * c], xl = xr[2 * c + 1];
    int flag = (cfinite(x0) && cfinite(xl)) ? 0 : 1;
    cplx cpv = czero(), dpv = x0;  // "row 0" acts as x_0 = x0
    for (int i = 1; i <= len - 2; ++i) {
        const cplx a = R[i][0];
        const cplx inv = frecip(nmsub(R[i][1], a, cpv));
        dpv = nmsub(R[i][3], a, dpv) * inv;
        cpv = R[i][2] * inv;
        R[i][2] = cpv, R[i][3] = dpv;
    }
    cplx xn = xl, x1 = xl, xm2 = x0;  // x(r0 + 1), x(r0 + len - 2)
    x[r0 + len - 1] = xl;
    for (int i = len - 2; i >= 1; --i) {
        xn = nmsub(R[i][3], R[i][2], xn);
        flag |= cfinite(xn) ? 0 : 1;
        x[r0 + i] = xn;
        if (i == len - 2) xm2 = xn;
        x1 = xn;
    }
    x[r0] = x0;
    if (len >= 2) {  // probe: the chunk's two interface equations (rows 0 and len - 1 are untouched by the sweep)
        // rows without a neighbour inside this slab: a mesh end has an exact zero there, a slab end couples to another
        // rank's unknown (then the equation is skipped: its coefficient is non-zero)
        const bool hp = c > 0, hn = 2 * c + 2 < n_next;
        if (hp || (R[0][0].re == 0.0 && R[0][0].im == 0.0))
            flag |= probe_row(R[0][0], R[0][1], R[0][2], R[0][3], hp ? xr[2 * c - 1] : czero(), x0, x1, anorm);
        const cplx *q = R[len - 1];
        if (hn || (q[2].re == 0.0 && q[2].im == 0.0))
            flag |= probe_row(q[0], q[1], q[2], q[3], xm2, xl, hn ? xr[2 * c + 2] : czero(), anorm);
    }
    return flag;
}
__device__ __forceinline__ const cplx *deep_xr(const DeepPlan &dp, int l) { return dp.ptr[l + 1].x; }
__device__ __forceinline__ cplx *deep_x(const DeepPlan &dp, int l, cplx *u_out) { return l == 0 ? u_out : dp.ptr[l].x; }

// levels `levels - 1`, `levels - 2`, ... while they are small (<= kDeepSmall rows), by ONE CTA.  The special rows of all
// of them are staged first (one round of L2 latency for the lot).
struct DeepStageSet {
    DeepStage lv[kDeepSmallLevelsMax];
};
__device__ void deep_backsub_small_levels(const UniformRows &u, const DeepPlan &dp, const DeepConsts *cs, DeepStageSet &set,
                                          cplx *u_out, int32_t *status) {
    int flag = 0;
    int j = 0;
    for (int l = dp.levels - 1; l >= 0 && dp.lv[l].n <= kDeepSmall; --l, ++j)
        deep_stage(u, dp, cs, nullptr, l, false, set.lv[j], 32 * j);
    __syncthreads();
    const double anorm = deep_anorm(cs);
    j = 0;
    for (int l = dp.levels - 1; l >= 0 && dp.lv[l].n <= kDeepSmall; --l, ++j) {
        const DeepLevel &L = dp.lv[l];
        const cplx *xr = deep_xr(dp, l);
        cplx *x = deep_x(dp, l, u_out);
        const int64_t n_next = dp.lv[l + 1].n;
        // the special chunks go to lane 0 of the last warps, ahead of their share of the streaming part (the sweep is
        // the critical path of the level)
        const int w = (kDeepThreads >> 5) - 1 - (int)(threadIdx.x >> 5);
        if (w < deep_special_count(L) && (threadIdx.x & 31) == 0)
            flag |= deep_backsub_special(L, anorm, set.lv[j], w, xr, n_next, x);
        __syncwarp();
        flag |= deep_backsub_plain(cs[l], L.clo, L.chi, xr, x);
        __syncthreads();
    }
    // the probes of all these levels at once, off the critical path of the level-by-level chain: warp w samples 32
    // chunks of the w-th small level (every level's solution is still in the workspace)
    {
        const int l = dp.levels - 1 - (int)(threadIdx.x >> 5);
        if (l >= 0 && dp.lv[l].n <= kDeepSmall)
            flag |= deep_probe_samples(cs[l], anorm, dp.lv[l], deep_xr(dp, l), dp.lv[l + 1].n, (int)(threadIdx.x & 31), 32);
    }
    flag = (int)__reduce_or_sync(0xffffffffu, (unsigned)flag);
    if (flag != 0 && (threadIdx.x & 31) == 0 && status != nullptr) atomicOr(status, flag);
}

// multi-GPU piece 1: reduce the slab to its two interface rows (iface8 = a[2] b[2] c[2] d[2])
__global__ void __launch_bounds__(kDeepThreads)
deep_reduce_kernel(UniformRows u, const __grid_constant__ DeepPlan dp, DeepConsts *cs, cplx *__restrict__ iface8) {
    __shared__ DeepStage sg;
    __shared__ DeepCarry carry[2];
    u.fetch_k();
    deep_reduce_levels(u, dp, cs, sg, carry);
    if (threadIdx.x < 8) iface8[threadIdx.x] = dp.ptr[dp.levels].a[threadIdx.x];
}
// multi-GPU piece 2: seed the 2-row level with the slab's end values, back-substitute the small levels
__global__ void __launch_bounds__(kDeepThreads)
deep_backsub_small_kernel(UniformRows u, const __grid_constant__ DeepPlan dp, const DeepConsts *cs,
                          const cplx *__restrict__ x_first_last, cplx *__restrict__ u_out, int32_t *__restrict__ status) {
    __shared__ DeepStageSet set;
    u.fetch_k();
    if (threadIdx.x < 2) dp.ptr[dp.levels].x[threadIdx.x] = x_first_last[threadIdx.x];
    __syncthreads();
    deep_backsub_small_levels(u, dp, cs, set, u_out, status);
}
// single GPU: both pieces in one kernel; the 2-row system in between is solved with partial pivoting
__global__ void __launch_bounds__(kDeepThreads)
deep_solve_small_kernel(UniformRows u, const __grid_constant__ DeepPlan dp, DeepConsts *cs, cplx *__restrict__ u_out,
                        int32_t *__restrict__ status) {
    __shared__ DeepStage sg;
    __shared__ DeepCarry carry[2];
    __shared__ DeepStageSet set;
    u.fetch_k();
    deep_reduce_levels(u, dp, cs, sg, carry);
    if (threadIdx.x == 0) {  // [b0 c0; a1 b1] x = [d0; d1]  (a0 and c1 are the exact zeros of the mesh ends)
        const LevelPtrs &p = dp.ptr[dp.levels];
        cplx b0 = p.b[0], c0 = p.c[0], d0 = p.d[0], a1 = p.a[1], b1 = p.b[1], d1 = p.d[1];
        auto abs2 = [](cplx z) { return fma(z.re, z.re, z.im * z.im); };
        if (abs2(a1) > abs2(b0)) {
            cplx tmp = b0; b0 = a1, a1 = tmp;
            tmp = c0, c0 = b1, b1 = tmp;
            tmp = d0, d0 = d1, d1 = tmp;
        }
        const cplx inv0 = crecip(b0), m = a1 * inv0;
        const cplx x1 = nmsub(d1, m, d0) * crecip(nmsub(b1, m, c0));
        const cplx x0 = nmsub(d0, c0, x1) * inv0;
        p.x[0] = x0, p.x[1] = x1;
        if ((!cfinite(x0) || !cfinite(x1)) && status != nullptr) atomicOr(status, 1);
    }
    __syncthreads();
    deep_backsub_small_levels(u, dp, cs, set, u_out, status);
}
// The streaming part of one LARGE level.  The end values of 128 chunks (256 consecutive entries of xr, 4 KB) arrive by
// one cp.async per thread, kDeepStages rounds ahead of their use, so the kernel is a pure store stream: with the loads
// issued by the consuming threads themselves (deep_backsub_plain) every round paid an L2 / DRAM round trip -- under
// write load a long one -- before its 32 KB of stores could go out (measured: 4.3 TB/s on the 2.1 GB of N = 2^27).
constexpr int kDeepStages = 4;
struct DeepRing {
    cplx buf[kDeepStages][2 * 128];
};
__device__ __forceinline__ int deep_backsub_plain_staged(const DeepConsts &C, int64_t c_begin, int64_t c_end,
                                                         const cplx *__restrict__ xr, cplx *__restrict__ x,
                                                         DeepRing &ring) {
    const int t = threadIdx.x, i = t & (kLM - 1), cl = t >> 4;
    const cplx al = C.tab[0][i], be = C.tab[1][i], ga = C.tab[2][i];
    const int64_t rounds = (c_end - c_begin + 127) / 128;
    auto fetch = [&](int64_t r) {  // (an empty group keeps the group count in step with the rounds)
        const int64_t e = 2 * (c_begin + r * 128) + t;
        if (r < rounds && e < 2 * c_end) cp_async16(&ring.buf[r % kDeepStages][t], xr + e);
        cp_async_commit();
    };
    for (int r = 0; r < kDeepStages - 1; ++r) fetch(r);
    bool bad = false;
    for (int64_t r = 0; r < rounds; ++r) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kDeepStages - 2) : "memory");  // this thread's share of round r
        __syncthreads();  // everybody's share has landed, and everybody is done with round r - 1 ...
        fetch(r + kDeepStages - 1);  // ... whose buffer is refilled
        const cplx *b = ring.buf[r % kDeepStages];
        const int64_t c0 = c_begin + r * 128 + cl;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t c = c0 + 16 * q;
            if (c < c_end) {
                const cplx xi = cmadd(cmadd(ga, al, b[2 * (cl + 16 * q)]), be, b[2 * (cl + 16 * q) + 1]);
                bad |= !cfinite(xi);
                stg_stream(x + c * kLM + i, xi);
            }
        }
    }
    return bad ? 1 : 0;
}

// one large level: CTA 0 takes the special chunks, every other CTA streams `per` contiguous rounds of plain chunks
__global__ void __launch_bounds__(kDeepThreads)
deep_backsub_level_kernel(UniformRows u, const __grid_constant__ DeepPlan dp, const DeepConsts *cs, int l, int64_t per,
                          cplx *__restrict__ u_out, int32_t *__restrict__ status) {
    __shared__ DeepStage sg;
    __shared__ DeepRing ring;
    u.fetch_k();
    const DeepLevel &L = dp.lv[l];
    const cplx *xr = deep_xr(dp, l);
    cplx *x = deep_x(dp, l, u_out);
    const int64_t n_next = dp.lv[l + 1].n;
    const double anorm = deep_anorm(cs);
    int flag = 0;
    if (blockIdx.x == 0) {
        deep_stage(u, dp, cs, nullptr, l, false, sg);
        __syncthreads();
        if ((int)threadIdx.x < deep_special_count(L))
            flag = deep_backsub_special(L, anorm, sg, threadIdx.x, xr, n_next, x);
        // (warp 0 holds the special chunks' serial sweeps; the other warps of this CTA take the samples meanwhile)
        flag |= deep_probe_samples(cs[l], anorm, L, xr, n_next, (int)threadIdx.x - 32, kDeepThreads - 32);
    } else {
        const int64_t b = blockIdx.x - 1;
        int64_t c_end = L.clo + (b + 1) * per * 128;
        if (c_end > L.chi) c_end = L.chi;
        flag = deep_backsub_plain_staged(cs[l], L.clo + b * per * 128, c_end, xr, x, ring);
    }
    flag = (int)__reduce_or_sync(0xffffffffu, (unsigned)flag);
    if (flag != 0 && (threadIdx.x & 31) == 0 && status != nullptr) atomicOr(status, flag);
}

// Tiny dense solve of the gathered interface system (2P rows, P = number of ranks): one thread,
// Gaussian elimination with partial pivoting.  iface: [P][4][2] = per rank (a, b, c, d) x (first, last).
constexpr int kIfaceMax = 32;  // up to 16 ranks (one NVSwitch domain)
typedef cplx IfaceMat[kIfaceMax][kIfaceMax + 1];
// Called by every thread of the CTA (A in shared memory; barriers inside); the solution is written by thread 0 to x
// (shared or global), *bad_out (thread 0's) tells whether it is finite.  LOAD(p, q) = entry q of rank p's 8 numbers.
template <class Load>
__device__ __forceinline__ void spike_interface_solve(Load load, int P, IfaceMat &A, cplx *x, bool *bad_out) {
    const int n = 2 * P;
    for (int i = threadIdx.x; i < n * (n + 1); i += blockDim.x) A[i / (n + 1)][i % (n + 1)] = czero();
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int p = 0; p < P; ++p)
            for (int r = 0; r < 2; ++r) {
                const int i = 2 * p + r;
                if (i > 0) A[i][i - 1] = load(p, 0 + r);
                A[i][i] = load(p, 2 + r);
                if (i < n - 1) A[i][i + 1] = load(p, 4 + r);
                A[i][n] = load(p, 6 + r);
            }
        bool bad = false;
        for (int k = 0; k < n; ++k) {
            int piv = k;
            double best = A[k][k].re * A[k][k].re + A[k][k].im * A[k][k].im;
            for (int i = k + 1; i < min(n, k + 3); ++i) {  // tridiagonal + fill: only two rows below can be non-zero
                const double v = A[i][k].re * A[i][k].re + A[i][k].im * A[i][k].im;
                if (v > best) best = v, piv = i;
            }
            if (best == 0.0) bad = true;
            if (piv != k)
                for (int j = k; j <= n; ++j) {
                    const cplx tmp = A[k][j];
                    A[k][j] = A[piv][j], A[piv][j] = tmp;
                }
            const cplx inv = crecip(A[k][k]);
            for (int i = k + 1; i < min(n, k + 3); ++i) {
                const cplx w = A[i][k] * inv;
                for (int j = k; j <= n; ++j) A[i][j] = nmsub(A[i][j], w, A[k][j]);
            }
        }
        for (int k = n - 1; k >= 0; --k) {
            cplx sum = A[k][n];
            for (int j = k + 1; j < min(n, k + 4); ++j) sum = nmsub(sum, A[k][j], x[j]);
            x[k] = sum * crecip(A[k][k]);
            bad |= !cfinite(x[k]);
        }
        *bad_out = bad;
    }
    __syncthreads();
}
__global__ void spike_interface_kernel(const cplx *__restrict__ iface, int P, cplx *__restrict__ x,
                                       int32_t *__restrict__ status) {
    __shared__ IfaceMat A;
    bool bad = false;
    spike_interface_solve([&](int p, int q) { return iface[(size_t)p * 8 + q]; }, P, A, x, &bad);
    if (threadIdx.x == 0 && bad && status != nullptr) atomicOr(status, 1);
}

// ---- the exchange over peer memory ---------------------------------------------------------------------------------
// Config 4's one exchange step is 128 bytes per rank, and what surrounds it on a rank are single-CTA kernels a few
// microseconds long: a library all-gather between them costs more than they do (NCCL's small-message latency plus
// two kernel boundaries).  So the exchange is part of the kernel: after its reduction, the CTA stores its two interface
// rows straight into every peer's exchange buffer (symmetric memory: the same allocation on every rank, mapped over
// NVLink), publishes a sequence number per peer with a system-scope release, waits -- acquire -- until every peer's
// number has arrived in its own buffer, solves the 2P-row interface system and goes on with the small back-substitution
// levels.  One kernel per rank from the first level down to the interface and back up to the 2^14-row level.
// Slots and flags are double-buffered by the parity of the solve counter (a rank can be at most one solve ahead of a
// peer: it cannot pass the wait of solve s before every peer has published s, i.e. has left solve s - 1 behind); the
// counter lives in device memory and is advanced by the kernel, so a captured graph can be replayed.
struct PeerSlots {
    cplx iface[2][kIfaceMax / 2][8];              // [parity][source rank][a0 a1 b0 b1 c0 c1 d0 d1]
    unsigned long long flag[2][kIfaceMax / 2];    // [parity][source rank] = solve counter of the data in the slot
};
struct PeerExchange {
    void *const *peer_bufs;     // device array [n_ranks]: every rank's PeerSlots as mapped into this process
    int rank, n_ranks;
    unsigned long long *seq;    // this rank's solve counter (device memory, starts at 0)
    unsigned long long timeout_ns;
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ cplx ld_volatile_global(const cplx *p) {  // (a peer wrote it: never from a stale L1 line)
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return mk(v.x, v.y);
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kDeepThreads)
deep_solve_exchange_kernel(UniformRows u, const __grid_constant__ DeepPlan dp, DeepConsts *cs, PeerExchange ex,
                           cplx *__restrict__ x_iface_out, cplx *__restrict__ u_out, int32_t *__restrict__ status) {
    __shared__ DeepStage sg;
    __shared__ DeepCarry carry[2];
    // the interface matrix is needed before the small levels are staged, the stage set after: one piece of memory
    constexpr size_t kRaw = sizeof(DeepStageSet) > sizeof(IfaceMat) ? sizeof(DeepStageSet) : sizeof(IfaceMat);
    __shared__ __align__(16) unsigned char raw[kRaw];
    __shared__ cplx xs[kIfaceMax];
    IfaceMat &A = *reinterpret_cast<IfaceMat *>(raw);
    DeepStageSet &set = *reinterpret_cast<DeepStageSet *>(raw);
    const int t = threadIdx.x, P = ex.n_ranks;
    u.fetch_k();
    deep_reduce_levels(u, dp, cs, sg, carry);
    const unsigned long long seq = *ex.seq + 1;
    const int par = (int)(seq & 1ull);
    // publish: 8 numbers to each of the P ranks (its own buffer included), then one flag per rank
    if (t < 8 * P) {
        const int p = t >> 3, e = t & 7;
        PeerSlots *dst = static_cast<PeerSlots *>(ex.peer_bufs[p]);
        dst->iface[par][ex.rank][e] = dp.ptr[dp.levels].a[e];  // (a[2] b[2] c[2] d[2] are contiguous)
        __threadfence_system();
    }
    __syncthreads();
    if (t < P) st_release_sys(&static_cast<PeerSlots *>(ex.peer_bufs[t])->flag[par][ex.rank], seq);
    // wait for every rank's rows in this rank's buffer
    const PeerSlots *mine = static_cast<const PeerSlots *>(ex.peer_bufs[ex.rank]);
    if (t < P) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(&mine->flag[par][t]) < seq) {
            if (global_timer_ns() - t0 > ex.timeout_ns) {  // a peer never arrived: report, do not hang the GPU
                if (status != nullptr) atomicOr(status, 5);
                break;
            }
        }
    }
    __syncthreads();
    bool bad = false;
    spike_interface_solve([&](int p, int q) { return ld_volatile_global(&mine->iface[par][p][q]); }, P, A, xs, &bad);
    if (t == 0 && bad && status != nullptr) atomicOr(status, 1);
    if (t < 2 * P) x_iface_out[t] = xs[t];
    if (t < 2) dp.ptr[dp.levels].x[t] = xs[2 * ex.rank + t];
    __syncthreads();  // (also: A is dead, `set` may be written)
    deep_backsub_small_levels(u, dp, cs, set, u_out, status != nullptr ? status + 1 : nullptr);
    if (t == 0) *ex.seq = seq;
}

// ---------------------------------------------------------------------------------------------- host side
struct LevelPlan {
    int levels;                     // number of reduced levels (level 0 is the caller's system)
    int64_t n[kMaxLevels + 1];      // rows per level
    size_t off[kMaxLevels + 1];     // byte offset of level l >= 1 inside the workspace (5 arrays of n[l])
    size_t off_anorm;               // one double: ||A||_inf of level 0, recorded by its reduction (the probe's unit)
    size_t total;
};

static LevelPlan plan_levels(int64_t n0, int64_t stop_at) {
    LevelPlan pl;
    pl.levels = 0, pl.n[0] = n0, pl.off[0] = 0, pl.total = 0;
    while (pl.n[pl.levels] > stop_at && pl.levels < kMaxLevels) {
        const int64_t nn = reduced_rows(pl.n[pl.levels]);
        pl.levels++;
        pl.n[pl.levels] = nn;
        pl.off[pl.levels] = pl.total;
        pl.total += 5 * sizeof(cplx) * (size_t)nn;
        pl.total = (pl.total + 255) & ~(size_t)255;
    }
    pl.off_anorm = pl.total;
    pl.total += 256;
    return pl;
}

static LevelPtrs level_ptrs(void *ws, const LevelPlan &pl, int l) {
    cplx *base = reinterpret_cast<cplx *>(static_cast<char *>(ws) + pl.off[l]);
    const int64_t n = pl.n[l];
    return {base, base + n, base + 2 * n, base + 3 * n, base + 4 * n};
}

static unsigned grid_for(int64_t n) {
    const int64_t tiles = tile_count(n);
    const int64_t cap = 3ll * sm_count() * 4;  // 3 resident CTAs per SM, a few waves, then grid-stride
    return (unsigned)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

static unsigned long long *anorm_slot(void *ws, const LevelPlan &pl) {
    return reinterpret_cast<unsigned long long *>(static_cast<char *>(ws) + pl.off_anorm);
}
template <class Source>
static int launch_reduce(const Source &src, const LevelPtrs &out, unsigned long long *anorm_out, cudaStream_t st) {
    auto kern = large_reduce_kernel<Source>;
    FH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LargeSmem)));
    if (anorm_out != nullptr) FH_CUDA(cudaMemsetAsync(anorm_out, 0, sizeof(unsigned long long), st));
    kern<<<grid_for(src.n), kLT, sizeof(LargeSmem), st>>>(src, out.a, out.b, out.c, out.d, anorm_out);
    FH_LAUNCH_CHECK("large_reduce_kernel");
    return FH_OK;
}
template <class Source>
static int launch_backsub(const Source &src, const cplx *xr, cplx *x, int32_t *status, const double *anorm,
                          cudaStream_t st) {
    auto kern = large_backsub_kernel<Source>;
    FH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LargeSmem)));
    kern<<<grid_for(src.n), kLT, sizeof(LargeSmem), st>>>(src, xr, x, status, anorm);
    FH_LAUNCH_CHECK("large_backsub_kernel");
    return FH_OK;
}

// ---- generated rows on a uniform mesh: the periodic recursion ---------------------------------------------------------
static bool uniform_rows(const HelmRows &src) { return src.p.h_mode == FH_H_SCALAR && src.p.n_elem >= 2; }
static UniformRows as_uniform(const HelmRows &src) {
    UniformRows u;
    u.p = src.p, u.k = src.k, u.k2 = src.k2, u.kk = src.kk, u.g_off = src.g_off, u.n = src.n;
    return u;
}
// the per-level constants live behind the level arrays of the workspace
static size_t deep_consts_offset(const LevelPlan &pl) { return (pl.total + 255) & ~(size_t)255; }
static size_t deep_workspace_bytes(int64_t n) {
    return deep_consts_offset(plan_levels(n, 2)) + sizeof(DeepConsts) * (kMaxLevels + 1);
}
// Geometry of the recursion for the rows [g_off, g_off + n) of a mesh with n_elem elements: which rows / chunks of every
// level are plain.  Only the last chunk of a level can be irregular (short, or 17 rows after a merge), and a chunk is
// plain when it has 16 rows, all of them plain.  false: not applicable (the caller takes the generic kernels).
static bool deep_geometry(int64_t g_off, int64_t n, int64_t n_elem, const LevelPlan &pl, DeepPlan &dp) {
    if (pl.levels < 1 || pl.n[pl.levels] != 2) return false;
    memset(&dp, 0, sizeof(dp));
    dp.levels = pl.levels;
    int64_t plo = g_off == 0 ? 1 : 0;                         // mesh row 0 has its own coefficients
    int64_t phi = n - (g_off + n - 1 == n_elem ? 1 : 0);      // ... and so has mesh row N
    for (int l = 0; l <= pl.levels; ++l) {
        DeepLevel &L = dp.lv[l];
        L.n = pl.n[l];
        const int64_t nt = tile_count(L.n);
        const int last_rows = tile_rows(L.n, nt, nt - 1), nch_last = chunk_count(last_rows);
        L.chunks = (nt - 1) * kLT + nch_last;
        const int last_len = last_rows - (nch_last - 1) * kLM;
        L.plo = plo, L.phi = phi;
        int64_t clo = (plo + kLM - 1) / kLM, chi = phi / kLM;
        const int64_t regular = L.chunks - (last_len != kLM ? 1 : 0);
        if (chi > regular) chi = regular;
        if (clo > L.chunks) clo = L.chunks;
        if (chi < clo) chi = clo;
        L.clo = clo, L.chi = chi;
        if (l < pl.levels && clo + (L.chunks - chi) > kDeepSpecialMax) return false;
        plo = 2 * clo, phi = 2 * chi;
    }
    int small = 0;
    for (int l = 0; l < pl.levels; ++l) small += dp.lv[l].n <= kDeepSmall ? 1 : 0;
    return small <= kDeepSmallLevelsMax;
}
static bool deep_plan(const HelmRows &src, void *ws, size_t ws_bytes, DeepPlan &dp, DeepConsts *&cs) {
    if (!uniform_rows(src) || ws == nullptr) return false;
    const LevelPlan pl = plan_levels(src.n, 2);
    if (ws_bytes < deep_workspace_bytes(src.n) || !deep_geometry(src.g_off, src.n, src.p.n_elem, pl, dp)) return false;
    for (int l = 1; l <= pl.levels; ++l) dp.ptr[l] = level_ptrs(ws, pl, l);
    cs = reinterpret_cast<DeepConsts *>(static_cast<char *>(ws) + deep_consts_offset(pl));
    return true;
}
// the levels the single-CTA kernel leaves over, largest last
static int deep_backsub_big_levels(const UniformRows &u, const DeepPlan &dp, const DeepConsts *cs, cplx *u_out,
                                   int32_t *status, cudaStream_t st) {
    int l = dp.levels - 1;
    while (l >= 0 && dp.lv[l].n <= kDeepSmall) --l;
    for (; l >= 0; --l) {
        const int64_t rounds = (dp.lv[l].chi - dp.lv[l].clo + 127) / 128;  // 32 KB of solution each
        const int64_t want = 16ll * sm_count();                              // a few waves of 2 resident CTAs per SM
        const int64_t per = rounds > want ? (rounds + want - 1) / want : 1;
        const int64_t ctas = (rounds + per - 1) / per;
        deep_backsub_level_kernel<<<(unsigned)(ctas + 1), kDeepThreads, 0, st>>>(u, dp, cs, l, per, u_out, status);
    }
    FH_LAUNCH_CHECK("deep_backsub_level_kernel");
    return FH_OK;
}

static GlobalBands bands_of(const LevelPtrs &p, int64_t n, int open_first, int open_last) {
    GlobalBands g;
    g.dl = p.a, g.d = p.b, g.du = p.c, g.b = p.d, g.n = n, g.open_first = open_first, g.open_last = open_last;
    return g;
}

// reduce level 0 (any source) and every further level down to the plan's last one
template <class Source0>
static int reduce_all(const Source0 &src0, const LevelPlan &pl, void *ws, int open_first, int open_last,
                      cudaStream_t st) {
    for (int l = 0; l < pl.levels; ++l) {
        const LevelPtrs out = level_ptrs(ws, pl, l + 1);
        int rc = l == 0 ? launch_reduce(src0, out, anorm_slot(ws, pl), st)
                        : launch_reduce(bands_of(level_ptrs(ws, pl, l), pl.n[l], open_first, open_last), out, nullptr, st);
        if (rc != FH_OK) return rc;
    }
    return FH_OK;
}
// back-substitute from the plan's last level (whose x is already filled) up to level 0
template <class Source0>
static int backsub_all(const Source0 &src0, const LevelPlan &pl, void *ws, int open_first, int open_last,
                       cplx *u, int32_t *status, cudaStream_t st) {
    for (int l = pl.levels - 1; l >= 0; --l) {
        const cplx *xr = level_ptrs(ws, pl, l + 1).x;
        const double *anorm = reinterpret_cast<const double *>(anorm_slot(ws, pl));
        int rc = l == 0 ? launch_backsub(src0, xr, u, status, anorm, st)
                        : launch_backsub(bands_of(level_ptrs(ws, pl, l), pl.n[l], open_first, open_last), xr,
                                         level_ptrs(ws, pl, l).x, status, anorm, st);
        if (rc != FH_OK) return rc;
    }
    return FH_OK;
}

template <class Source0>
static int solve_large_any(const Source0 &src0, cplx *u, int64_t n, int32_t *status, void *ws, size_t ws_bytes,
                           cudaStream_t st) {
    const LevelPlan pl = plan_levels(n, kCoarseMax);
    if (ws_bytes < pl.total || (pl.total > 0 && ws == nullptr)) {
        set_error("workspace too small: need %zu bytes, got %zu", pl.total, ws_bytes);
        return FH_E_WORKSPACE;
    }
    if (status != nullptr) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    int rc = reduce_all(src0, pl, ws, 0, 0, st);
    if (rc != FH_OK) return rc;
    const LevelPtrs top = level_ptrs(ws, pl, pl.levels);
    // (the batched kernel zeroes status[0] itself -- nothing has been flagged yet: the reductions do not write it -- and
    // ORs in its own probe / breakdown bits; the back-substitution levels OR theirs on top)
    rc = batched_solve_bands(top.a, top.b, top.c, top.d, top.x, pl.n[pl.levels], 1, status, st);
    if (rc != FH_OK) return rc;
    return backsub_all(src0, pl, ws, 0, 0, u, status, st);
}

size_t solve_large_workspace_bytes(int64_t n) {  // stored bands, or generated rows (periodic recursion)
    const size_t a = plan_levels(n, kCoarseMax).total, b = deep_workspace_bytes(n);
    return a > b ? a : b;
}

int solve_large(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n, int32_t *status,
                void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n <= kCoarseMax) return batched_solve_bands(dl, d, du, b, u, n, 1, status, stream);
    GlobalBands src;
    src.dl = dl, src.d = d, src.du = du, src.b = b, src.n = n, src.open_first = 0, src.open_last = 0;
    return solve_large_any(src, u, n, status, workspace, workspace_bytes, stream);
}

static int helm_rows_from(const fh_problem1d *prob_host, double k, double k2, int64_t row_begin, int64_t n_rows,
                          HelmRows &src) {
    int rc = problem_from_host(prob_host, src.p);
    if (rc != FH_OK) return rc;
    FH_REQUIRE(row_begin >= 0 && n_rows >= 1 && row_begin + n_rows <= src.p.n_elem + 1,
               "row range [%lld, %lld) outside the mesh of %lld nodes", (long long)row_begin,
               (long long)(row_begin + n_rows), (long long)(src.p.n_elem + 1));
    src.k = k, src.k2 = k2, src.g_off = row_begin, src.n = n_rows;
    return FH_OK;
}
static int kdev_unsupported() {
    set_error("a device-resident wavenumber needs a uniform mesh (scalar h, >= 2 elements) and a workspace of "
              "fh_spike_workspace_bytes / fh_tridiag_solve_large_workspace_bytes");
    return FH_E_UNSUPPORTED;
}

}  // namespace fh

extern "C" {

#ifdef FH_DEEP_TRACE
__attribute__((visibility("default"))) int fh_debug_deep_trace(long long *out64) {
    return (int)cudaMemcpyFromSymbol(out64, fh::g_deep_trace, sizeof(long long) * 64);
}
#endif
#ifdef FH_PROBE_STAT
__attribute__((visibility("default"))) double fh_debug_probe_max(int reset) {
    unsigned long long v = 0ull, z = 0ull;
    cudaMemcpyFromSymbol(&v, fh::g_probe_max, sizeof(v));
    if (reset) cudaMemcpyToSymbol(fh::g_probe_max, &z, sizeof(z));
    double d;
    memcpy(&d, &v, sizeof(d));
    return d;
}
#endif

size_t fh_tridiag_solve_large_workspace_bytes(int64_t n) { return fh::solve_large_workspace_bytes(n); }

int fh_large_uniform_plan(int64_t n_elem, int64_t row_begin, int64_t n_rows, int64_t *out6, int max_levels) {
    using namespace fh;
    FH_REQUIRE(out6 != nullptr && n_elem >= 2 && row_begin >= 0 && n_rows > 2 && row_begin + n_rows <= n_elem + 1,
               "bad row range or NULL output");
    DeepPlan dp;
    if (!deep_geometry(row_begin, n_rows, n_elem, plan_levels(n_rows, 2), dp)) {
        set_error("the periodic recursion does not apply to this row range");
        return FH_E_UNSUPPORTED;
    }
    FH_REQUIRE(max_levels > dp.levels, "out6 holds %d levels, the plan has %d", max_levels, dp.levels + 1);
    for (int l = 0; l <= dp.levels; ++l) {
        const DeepLevel &L = dp.lv[l];
        int64_t *o = out6 + 6 * l;
        o[0] = L.n, o[1] = L.chunks, o[2] = L.plo, o[3] = L.phi, o[4] = L.clo, o[5] = L.chi;
    }
    return dp.levels + 1;
}

int fh_tridiag_solve_large_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b, fh_c128 *u,
                                int64_t n, int32_t *status, void *workspace, size_t workspace_bytes,
                                fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return FH_OK;
    FH_REQUIRE(dl && d && du && b && u, "NULL device pointer");
    return solve_large(reinterpret_cast<const cplx *>(dl), reinterpret_cast<const cplx *>(d),
                       reinterpret_cast<const cplx *>(du), reinterpret_cast<const cplx *>(b),
                       reinterpret_cast<cplx *>(u), n, status, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

// kk != NULL: (k, k**2) live in device memory (uniform meshes only: the periodic recursion)
static int helm_solve_large_fused(const fh_problem1d *prob_host, double k, double k2, const double *kk, fh_c128 *u,
                                  int32_t *status, void *workspace, size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host != nullptr && u != nullptr, "NULL pointer");
    HelmRows src;
    int rc = helm_rows_from(prob_host, k, k2, 0, prob_host->n_elem + 1, src);
    if (rc != FH_OK) return rc;
    src.kk = kk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeepPlan dp;
    DeepConsts *cs = nullptr;
    if (deep_plan(src, workspace, workspace_bytes, dp, cs)) {  // uniform mesh: the periodic recursion
        if (status != nullptr) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
        const UniformRows ur = as_uniform(src);
        deep_solve_small_kernel<<<1, kDeepThreads, 0, st>>>(ur, dp, cs, reinterpret_cast<cplx *>(u), status);
        FH_LAUNCH_CHECK("deep_solve_small_kernel");
        return deep_backsub_big_levels(ur, dp, cs, reinterpret_cast<cplx *>(u), status, st);
    }
    if (kk != nullptr) return kdev_unsupported();
    return solve_large_any(src, reinterpret_cast<cplx *>(u), src.n, status, workspace, workspace_bytes, st);
}
int fh_helmholtz1d_solve_large_fused_c128(const fh_problem1d *prob_host, double k, double k2, fh_c128 *u,
                                          int32_t *status, void *workspace, size_t workspace_bytes,
                                          fh_stream_t stream) {
    return helm_solve_large_fused(prob_host, k, k2, nullptr, u, status, workspace, workspace_bytes, stream);
}
int fh_helmholtz1d_solve_large_fused_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev, fh_c128 *u,
                                               int32_t *status, void *workspace, size_t workspace_bytes,
                                               fh_stream_t stream) {
    FH_REQUIRE(k_k2_dev != nullptr, "NULL pointer");
    return helm_solve_large_fused(prob_host, 0.0, 0.0, k_k2_dev, u, status, workspace, workspace_bytes, stream);
}

// One long mesh, A kept: the level-0 reduction generates the rows, stores them as the bands fh_assemble_1d_c128 would
// have stored (bit for bit) and reduces them in the same pass; everything else -- the further levels, the coarse cluster
// solve, the back-substitution (which reads the stored bands) -- is fh_tridiag_solve_large_c128's.  162 instead of 226
// bytes per row, one launch less, and the bits of the two-call sequence.
int fh_helmholtz1d_assemble_solve_large_c128(const fh_problem1d *prob_host, double k, double k2, const double *k_k2_dev,
                                             fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b, fh_c128 *u, int32_t *status,
                                             void *workspace, size_t workspace_bytes, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host && dl && d && du && b && u, "NULL pointer");
    HelmRowsStore src;
    int rc = helm_rows_from(prob_host, k, k2, 0, prob_host->n_elem + 1, src);
    if (rc != FH_OK) return rc;
    if (src.n <= kCoarseMax) {
        set_error("meshes of up to %lld nodes: fh_helmholtz1d_sweep_assemble_solve_c128", (long long)kCoarseMax);
        return FH_E_UNSUPPORTED;
    }
    src.kk = k_k2_dev;
    src.out_dl = reinterpret_cast<cplx *>(dl), src.out_d = reinterpret_cast<cplx *>(d);
    src.out_du = reinterpret_cast<cplx *>(du), src.out_b = reinterpret_cast<cplx *>(b);
    const LevelPlan pl = plan_levels(src.n, kCoarseMax);
    if (workspace_bytes < pl.total || workspace == nullptr) {
        set_error("workspace too small: need %zu bytes, got %zu (fh_tridiag_solve_large_workspace_bytes)", pl.total,
                  workspace_bytes);
        return FH_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (status != nullptr) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    rc = reduce_all(src, pl, workspace, 0, 0, st);
    if (rc != FH_OK) return rc;
    const LevelPtrs top = level_ptrs(workspace, pl, pl.levels);
    rc = batched_solve_bands(top.a, top.b, top.c, top.d, top.x, pl.n[pl.levels], 1, status, st);
    if (rc != FH_OK) return rc;
    GlobalBands stored;
    stored.dl = src.out_dl, stored.d = src.out_d, stored.du = src.out_du, stored.b = src.out_b;
    stored.n = src.n, stored.open_first = 0, stored.open_last = 0;
    return backsub_all(stored, pl, workspace, 0, 0, reinterpret_cast<cplx *>(u), status, st);
}

// ---- multi-GPU SPIKE pieces: every rank owns the contiguous rows [row_begin, row_begin + n_rows) ------------
size_t fh_spike_workspace_bytes(int64_t n_rows) {
    const size_t a = fh::plan_levels(n_rows, 2).total, b = fh::deep_workspace_bytes(n_rows);
    return a > b ? a : b;
}

int fh_spike_reduce_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b, int64_t n_rows,
                         int open_first, int open_last, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                         fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(n_rows >= 2, "a slab needs at least 2 rows");
    FH_REQUIRE(dl && d && du && b && iface8, "NULL device pointer");
    const LevelPlan pl = plan_levels(n_rows, 2);
    if (workspace_bytes < pl.total || (pl.total > 0 && !workspace)) {
        set_error("workspace too small: need %zu bytes", pl.total);
        return FH_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GlobalBands src;
    src.dl = reinterpret_cast<const cplx *>(dl), src.d = reinterpret_cast<const cplx *>(d);
    src.du = reinterpret_cast<const cplx *>(du), src.b = reinterpret_cast<const cplx *>(b);
    src.n = n_rows, src.open_first = open_first, src.open_last = open_last;
    if (pl.levels == 0) {  // the slab already is its own interface
        const cplx *arrs[4] = {src.dl, src.d, src.du, src.b};
        for (int i = 0; i < 4; ++i)
            FH_CUDA(cudaMemcpyAsync(reinterpret_cast<cplx *>(iface8) + 2 * i, arrs[i], 2 * sizeof(cplx),
                                    cudaMemcpyDeviceToDevice, st));
        return FH_OK;
    }
    int rc = reduce_all(src, pl, workspace, open_first, open_last, st);
    if (rc != FH_OK) return rc;
    const LevelPtrs top = level_ptrs(workspace, pl, pl.levels);  // 2 rows: a[2] b[2] c[2] d[2] contiguous
    FH_CUDA(cudaMemcpyAsync(iface8, top.a, 8 * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    return FH_OK;
}

static int helm_spike_reduce(const fh_problem1d *prob_host, double k, double k2, const double *kk, int64_t row_begin,
                             int64_t n_rows, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                             fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host != nullptr && iface8 != nullptr, "NULL pointer");
    FH_REQUIRE(n_rows > 2, "a generated slab needs more than 2 rows");
    HelmRows src;
    int rc = helm_rows_from(prob_host, k, k2, row_begin, n_rows, src);
    if (rc != FH_OK) return rc;
    src.kk = kk;
    const LevelPlan pl = plan_levels(n_rows, 2);
    if (workspace_bytes < pl.total || !workspace) {
        set_error("workspace too small: need %zu bytes", pl.total);
        return FH_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeepPlan dp;
    DeepConsts *cs = nullptr;
    if (deep_plan(src, workspace, workspace_bytes, dp, cs)) {  // uniform mesh: the periodic recursion, one launch
        deep_reduce_kernel<<<1, kDeepThreads, 0, st>>>(as_uniform(src), dp, cs, reinterpret_cast<cplx *>(iface8));
        FH_LAUNCH_CHECK("deep_reduce_kernel");
        return FH_OK;
    }
    if (kk != nullptr) return kdev_unsupported();
    rc = reduce_all(src, pl, workspace, 1, 1, st);  // generated rows carry exact zeros at the mesh ends
    if (rc != FH_OK) return rc;
    const LevelPtrs top = level_ptrs(workspace, pl, pl.levels);
    FH_CUDA(cudaMemcpyAsync(iface8, top.a, 8 * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    return FH_OK;
}
int fh_helmholtz1d_spike_reduce_c128(const fh_problem1d *prob_host, double k, double k2, int64_t row_begin,
                                     int64_t n_rows, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                                     fh_stream_t stream) {
    return helm_spike_reduce(prob_host, k, k2, nullptr, row_begin, n_rows, workspace, workspace_bytes, iface8, stream);
}
int fh_helmholtz1d_spike_reduce_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev, int64_t row_begin,
                                          int64_t n_rows, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                                          fh_stream_t stream) {
    FH_REQUIRE(k_k2_dev != nullptr, "NULL pointer");
    return helm_spike_reduce(prob_host, 0.0, 0.0, k_k2_dev, row_begin, n_rows, workspace, workspace_bytes, iface8,
                             stream);
}

size_t fh_spike_peer_buffer_bytes(void) { return sizeof(fh::PeerSlots); }

int fh_helmholtz1d_spike_solve_peer_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev, int64_t row_begin,
                                              int64_t n_rows, void *workspace, size_t workspace_bytes,
                                              void *const *peer_bufs_dev, int rank, int n_ranks,
                                              unsigned long long *seq_dev, double timeout_s, fh_c128 *x_iface,
                                              fh_c128 *u, int32_t *status2, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host && k_k2_dev && workspace && peer_bufs_dev && seq_dev && x_iface && u, "NULL pointer");
    FH_REQUIRE(n_ranks >= 1 && 2 * n_ranks <= kIfaceMax && rank >= 0 && rank < n_ranks, "bad rank / n_ranks (<= %d ranks)",
               kIfaceMax / 2);
    FH_REQUIRE(n_rows > 2, "a generated slab needs more than 2 rows");
    HelmRows src;
    int rc = helm_rows_from(prob_host, 0.0, 0.0, row_begin, n_rows, src);
    if (rc != FH_OK) return rc;
    src.kk = k_k2_dev;
    DeepPlan dp;
    DeepConsts *cs = nullptr;
    if (!deep_plan(src, workspace, workspace_bytes, dp, cs)) return kdev_unsupported();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PeerExchange ex;
    ex.peer_bufs = peer_bufs_dev, ex.rank = rank, ex.n_ranks = n_ranks, ex.seq = seq_dev;
    ex.timeout_ns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 2.0) * 1e9);
    const UniformRows ur = as_uniform(src);
    deep_solve_exchange_kernel<<<1, kDeepThreads, 0, st>>>(ur, dp, cs, ex, reinterpret_cast<cplx *>(x_iface),
                                                           reinterpret_cast<cplx *>(u), status2);
    FH_LAUNCH_CHECK("deep_solve_exchange_kernel");
    return deep_backsub_big_levels(ur, dp, cs, reinterpret_cast<cplx *>(u), status2 != nullptr ? status2 + 1 : nullptr, st);
}

int fh_helmholtz1d_spike_reduce_store_c128(const fh_problem1d *prob_host, double k, double k2, const double *k_k2_dev,
                                           int64_t row_begin, int64_t n_rows, fh_c128 *dl, fh_c128 *d, fh_c128 *du,
                                           fh_c128 *b, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                                           fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host && dl && d && du && b && iface8, "NULL pointer");
    FH_REQUIRE(n_rows > 2, "a generated slab needs more than 2 rows");
    HelmRowsStore src;
    int rc = helm_rows_from(prob_host, k, k2, row_begin, n_rows, src);
    if (rc != FH_OK) return rc;
    src.kk = k_k2_dev;
    src.out_dl = reinterpret_cast<cplx *>(dl), src.out_d = reinterpret_cast<cplx *>(d);
    src.out_du = reinterpret_cast<cplx *>(du), src.out_b = reinterpret_cast<cplx *>(b);
    const LevelPlan pl = plan_levels(n_rows, 2);
    if (workspace_bytes < pl.total || !workspace) {
        set_error("workspace too small: need %zu bytes", pl.total);
        return FH_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // level 0: rows generated, stored and reduced in one pass; the levels below as for stored bands (the slab's outermost
    // rows keep their coupling to the neighbouring ranks: generated rows carry exact zeros at the mesh ends)
    rc = reduce_all(src, pl, workspace, 1, 1, st);
    if (rc != FH_OK) return rc;
    const LevelPtrs top = level_ptrs(workspace, pl, pl.levels);
    FH_CUDA(cudaMemcpyAsync(iface8, top.a, 8 * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    return FH_OK;
}

int fh_spike_solve_interface_c128(const fh_c128 *iface_all, int n_ranks, fh_c128 *x_iface, int32_t *status,
                                  fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(iface_all && x_iface, "NULL device pointer");
    FH_REQUIRE(n_ranks >= 1 && 2 * n_ranks <= kIfaceMax, "n_ranks must be in [1, %d]", kIfaceMax / 2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (status != nullptr) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    auto kern = spike_interface_kernel;
    kern<<<1, 128, 0, st>>>(reinterpret_cast<const cplx *>(iface_all), n_ranks, reinterpret_cast<cplx *>(x_iface),
                            status);
    FH_LAUNCH_CHECK("spike_interface_kernel");
    return FH_OK;
}

static int spike_seed_top(const fh::LevelPlan &pl, void *workspace, const fh_c128 *x_first_last, cudaStream_t st) {
    const fh::LevelPtrs top = fh::level_ptrs(workspace, pl, pl.levels);
    FH_CUDA(cudaMemcpyAsync(top.x, x_first_last, 2 * sizeof(fh::cplx), cudaMemcpyDeviceToDevice, st));
    return FH_OK;
}

int fh_spike_backsub_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b, int64_t n_rows,
                          int open_first, int open_last, void *workspace, size_t workspace_bytes,
                          const fh_c128 *x_first_last, fh_c128 *u, int32_t *status, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(dl && d && du && b && x_first_last && u, "NULL device pointer");
    FH_REQUIRE(n_rows >= 2, "a slab needs at least 2 rows");
    const LevelPlan pl = plan_levels(n_rows, 2);
    FH_REQUIRE(workspace_bytes >= pl.total, "workspace too small: need %zu bytes", pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pl.levels == 0) {
        FH_CUDA(cudaMemcpyAsync(u, x_first_last, 2 * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
        return FH_OK;
    }
    int rc = spike_seed_top(pl, workspace, x_first_last, st);
    if (rc != FH_OK) return rc;
    GlobalBands src;
    src.dl = reinterpret_cast<const cplx *>(dl), src.d = reinterpret_cast<const cplx *>(d);
    src.du = reinterpret_cast<const cplx *>(du), src.b = reinterpret_cast<const cplx *>(b);
    src.n = n_rows, src.open_first = open_first, src.open_last = open_last;
    return backsub_all(src, pl, workspace, open_first, open_last, reinterpret_cast<cplx *>(u), status, st);
}

static int helm_spike_backsub(const fh_problem1d *prob_host, double k, double k2, const double *kk, int64_t row_begin,
                              int64_t n_rows, void *workspace, size_t workspace_bytes, const fh_c128 *x_first_last,
                              fh_c128 *u, int32_t *status, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host && x_first_last && u, "NULL pointer");
    HelmRows src;
    int rc = helm_rows_from(prob_host, k, k2, row_begin, n_rows, src);
    if (rc != FH_OK) return rc;
    src.kk = kk;
    const LevelPlan pl = plan_levels(n_rows, 2);
    FH_REQUIRE(workspace_bytes >= pl.total && pl.levels > 0, "workspace too small: need %zu bytes", pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeepPlan dp;
    DeepConsts *cs = nullptr;
    if (deep_plan(src, workspace, workspace_bytes, dp, cs)) {  // (the same decision as in the reduce call)
        const UniformRows ur = as_uniform(src);
        deep_backsub_small_kernel<<<1, kDeepThreads, 0, st>>>(ur, dp, cs, reinterpret_cast<const cplx *>(x_first_last),
                                                              reinterpret_cast<cplx *>(u), status);
        FH_LAUNCH_CHECK("deep_backsub_small_kernel");
        return deep_backsub_big_levels(ur, dp, cs, reinterpret_cast<cplx *>(u), status, st);
    }
    if (kk != nullptr) return kdev_unsupported();
    rc = spike_seed_top(pl, workspace, x_first_last, st);
    if (rc != FH_OK) return rc;
    return backsub_all(src, pl, workspace, 1, 1, reinterpret_cast<cplx *>(u), status, st);
}
int fh_helmholtz1d_spike_backsub_c128(const fh_problem1d *prob_host, double k, double k2, int64_t row_begin,
                                      int64_t n_rows, void *workspace, size_t workspace_bytes,
                                      const fh_c128 *x_first_last, fh_c128 *u, int32_t *status, fh_stream_t stream) {
    return helm_spike_backsub(prob_host, k, k2, nullptr, row_begin, n_rows, workspace, workspace_bytes, x_first_last, u,
                              status, stream);
}
int fh_helmholtz1d_spike_backsub_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev, int64_t row_begin,
                                           int64_t n_rows, void *workspace, size_t workspace_bytes,
                                           const fh_c128 *x_first_last, fh_c128 *u, int32_t *status,
                                           fh_stream_t stream) {
    FH_REQUIRE(k_k2_dev != nullptr, "NULL pointer");
    return helm_spike_backsub(prob_host, 0.0, 0.0, k_k2_dev, row_begin, n_rows, workspace, workspace_bytes, x_first_last,
                              u, status, stream);
}

}  // extern "C"
