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
    int64_t g_off, n;
    __device__ __forceinline__ void fill(LargeSmem &sm, int64_t row0, int rows) const {
        // scalar-h (the reference's mode): every interior row is the same -- evaluate it once per thread
        const bool scalar = p.h_mode == FH_H_SCALAR && p.n_elem >= 2;
        cplx il = czero(), id = czero(), iu = czero(), ib = czero();
        if (scalar) row_values(p, row_geom(p, 1), 1, k, k2, il, id, iu, ib);
        for (int j = threadIdx.x; j < rows; j += kLT) {
            const int64_t g = g_off + row0 + j;
            cplx lo = il, di = id, hi = iu, rhs = ib;
            if (!scalar || g == 0 || g == p.n_elem) row_values(p, row_geom(p, g), g, k, k2, lo, di, hi, rhs);
            const int s = lswz(j);
            sm.a[s] = lo, sm.b[s] = di, sm.c[s] = hi, sm.d[s] = rhs;
        }
        __syncthreads();
    }
};

// ---------------------------------------------------------------------------------------------- reduce
template <class Source>
__global__ void __launch_bounds__(kLT, 3)
large_reduce_kernel(Source src, cplx *__restrict__ ra, cplx *__restrict__ rb, cplx *__restrict__ rc,
                    cplx *__restrict__ rd) {
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
large_backsub_kernel(Source src, const cplx *__restrict__ xr, cplx *__restrict__ x, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LargeSmem &sm = *reinterpret_cast<LargeSmem *>(smem_raw);
    const int64_t n_tiles = tile_count(src.n);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kLTile;
        const int rows = tile_rows(src.n, n_tiles, tile);
        src.fill(sm, row0, rows);
        const int t = threadIdx.x, r0 = t * kLM;
        const int nch = chunk_count(rows);
        const int len = t < nch ? (t == nch - 1 ? rows - r0 : kLM) : 0;
        bool bad = false;
        if (len >= 1) {
            const int64_t o = 2 * (tile * kLT + t);
            const cplx x0 = xr[o], xl = len > 1 ? xr[o + 1] : x0;
            bad = !cfinite(x0) || !cfinite(xl);
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
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        for (int j = threadIdx.x; j < rows; j += kLT) stg_stream(x + row0 + j, sm.d[lswz(j)]);
        if (any_bad && status != nullptr && threadIdx.x == 0) atomicOr(status, 1);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- uniform meshes
// Level 0 of a GENERATED system on a uniform mesh (scalar h: the reference's only mode, nb:158).  Every interior row
// holds the same four numbers, so every full interior chunk reduces to the same two interface rows, and the
// back-substitution of such a chunk is the same linear map of its two known end values:
//     x_i = alpha_i x_first + beta_i x_last + gamma_i,      i = 1 .. 14
// (alpha, beta, gamma: the chunk's Thomas sweep carried out once, symbolically in x_first / x_last).  What is left of
// the two level-0 passes is a structured fill of the interface rows (8 B per row) and a streaming kernel that reads two
// values per chunk and writes the solution (18 B per row) -- the elimination itself is unchanged, it is just not
// repeated 2^23 times.  The chunks that are not "full interior" (those holding mesh row 0 or N, and the short / merged
// chunk at the end of the last tile) take the generic sweeps, with rows generated on the fly.
struct UniformRows {  // what HelmRows provides, with the interior row evaluated once on the host side of the launch
    Problem1d p;
    double k, k2;
    int64_t g_off, n;
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
__device__ __forceinline__ ChunkPos chunk_pos(int64_t n, int64_t n_tiles, int64_t c) {
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
// full chunk of 16 interior rows?
__device__ __forceinline__ bool chunk_is_plain(const UniformRows &u, const ChunkPos &cp) {
    const int64_t g0 = u.g_off + cp.r0;
    return cp.len == kLM && g0 != 0 && g0 + cp.len - 1 != u.p.n_elem;
}

// the two interface rows of a chunk whose rows come from `rowfn(i, a, b, c, d)` -- the arithmetic of large_reduce_kernel
template <class RowFn>
__device__ __forceinline__ void reduce_chunk(RowFn rowfn, int len, cplx out1[4], cplx out2[4]) {
    cplx ra, rb, rc, rd;
    if (len == 1) {
        rowfn(0, ra, rb, rc, rd);
        out1[0] = ra, out1[1] = rb, out1[2] = czero(), out1[3] = rd;
        out2[0] = czero(), out2[1] = cone(), out2[2] = czero(), out2[3] = czero();
        return;
    }
    {  // downward from row 1: fill-in couples to x_first
        cplx pa, pb, pc, pd;
        rowfn(1, pa, pb, pc, pd);
        for (int i = 2; i < len; ++i) {
            rowfn(i, ra, rb, rc, rd);
            const cplx w = ra * frecip(pb);
            const cplx nb = nmsub(rb, w, pc), nd = nmsub(rd, w, pd);
            pa = -(w * pa), pb = nb, pc = rc, pd = nd;
        }
        out2[0] = pa, out2[1] = pb, out2[2] = pc, out2[3] = pd;
    }
    {  // upward from row len-2: fill-in couples to x_last
        cplx qa, qb, qc, qd;
        rowfn(len - 2, qa, qb, qc, qd);
        for (int i = len - 3; i >= 0; --i) {
            rowfn(i, ra, rb, rc, rd);
            const cplx w = rc * frecip(qb);
            const cplx nb = nmsub(rb, w, qa), nd = nmsub(rd, w, qd);
            qc = -(w * qc), qb = nb, qa = ra, qd = nd;
        }
        out1[0] = qa, out1[1] = qb, out1[2] = qc, out1[3] = qd;
    }
}

constexpr int kUniThreads = 256;

// Which chunks the streaming kernels cover: every chunk of the full tiles (tiles 0 .. n_tiles - 2 hold exactly 64 chunks
// of 16 rows: chunk c = rows 16 c .. 16 c + 15), except chunk 0 when it holds mesh row 0.  The last tile (whose last
// chunk may be short or merged, and which holds mesh row N when the slab ends the mesh) and that first chunk go to the
// one-thread-per-chunk kernels below.
struct UniformSpan {
    int64_t c_lo, c_hi;  // streaming chunks [c_lo, c_hi)
};
__host__ __device__ __forceinline__ UniformSpan uniform_span(int64_t g_off, int64_t n) {
    UniformSpan sp;
    sp.c_lo = g_off == 0 ? 1 : 0;
    sp.c_hi = (tile_count(n) - 1) * kLT;
    if (sp.c_hi < sp.c_lo) sp.c_hi = sp.c_lo;
    return sp;
}
// special chunk number `idx` (0 .. kLT): the chunks of the last tile, then chunk 0 (when it is not a streaming chunk
// and not part of the last tile)
__device__ __forceinline__ int64_t special_chunk(const UniformRows &u, int idx) {
    const int64_t n_tiles = tile_count(u.n);
    if (idx < kLT) return (n_tiles - 1) * kLT + idx;
    if (idx == kLT && n_tiles > 1 && u.g_off == 0) return 0;
    return -1;
}

// Structured fill: the interface rows of a plain chunk, computed once per CTA, written for every streaming chunk;
// every CTA owns a contiguous range of chunks.
__global__ void __launch_bounds__(kUniThreads)
uniform_reduce_kernel(UniformRows u, int64_t per_block, cplx *__restrict__ ra, cplx *__restrict__ rb,
                      cplx *__restrict__ rc, cplx *__restrict__ rd) {
    __shared__ cplx plain[2][4];
    if (threadIdx.x == 0) {
        cplx in[4], p1[4], p2[4];
        u.interior(in[0], in[1], in[2], in[3]);
        reduce_chunk([&](int, cplx &a, cplx &b, cplx &c, cplx &d) { a = in[0], b = in[1], c = in[2], d = in[3]; }, kLM,
                     p1, p2);
        for (int q = 0; q < 4; ++q) plain[0][q] = p1[q], plain[1][q] = p2[q];
    }
    __syncthreads();
    // thread t writes entry o = 2 c + (t & 1) of the four interface arrays: consecutive lanes, consecutive entries
    const int half = threadIdx.x & 1;
    const cplx va = plain[half][0], vb = plain[half][1], vc = plain[half][2], vd = plain[half][3];
    const UniformSpan sp = uniform_span(u.g_off, u.n);
    const int64_t o_begin = 2 * sp.c_lo + (int64_t)blockIdx.x * per_block;
    int64_t o_end = o_begin + per_block;
    if (o_end > 2 * sp.c_hi) o_end = 2 * sp.c_hi;
    for (int64_t o = o_begin + threadIdx.x; o < o_end; o += kUniThreads)   // (per_block is even: parity = t & 1)
        stg_stream(ra + o, va), stg_stream(rb + o, vb), stg_stream(rc + o, vc), stg_stream(rd + o, vd);
}

// ... and the special chunks: generic sweeps, rows generated on the fly, one thread per chunk
__global__ void uniform_special_reduce_kernel(UniformRows u, cplx *__restrict__ ra, cplx *__restrict__ rb,
                                              cplx *__restrict__ rc, cplx *__restrict__ rd) {
    const int64_t c = special_chunk(u, (int)threadIdx.x);
    if (c < 0) return;
    const ChunkPos cp = chunk_pos(u.n, tile_count(u.n), c);
    if (!cp.exists) return;
    cplx in[4], o1[4], o2[4];
    u.interior(in[0], in[1], in[2], in[3]);
    const int64_t g0 = u.g_off + cp.r0;
    reduce_chunk([&](int i, cplx &a, cplx &b, cplx &c2, cplx &d) { u.row(g0 + i, in, a, b, c2, d); }, cp.len, o1, o2);
    const int64_t o = 2 * c;
    ra[o] = o1[0], rb[o] = o1[1], rc[o] = o1[2], rd[o] = o1[3];
    ra[o + 1] = o2[0], rb[o + 1] = o2[1], rc[o + 1] = o2[2], rd[o + 1] = o2[3];
}

// Streaming back-substitution: x(16 c + i) = alpha_i x_first(c) + beta_i x_last(c) + gamma_i.  Every CTA owns a
// contiguous range of rows that starts on a chunk boundary; 256 threads = 16 chunks per pass, so a thread keeps the same
// row-in-chunk i -- and with it (alpha_i, beta_i, gamma_i) in registers -- for its whole range.
__global__ void __launch_bounds__(kUniThreads)
uniform_backsub_kernel(UniformRows u, int64_t chunks_per_block, const cplx *__restrict__ xr, cplx *__restrict__ x,
                       int32_t *__restrict__ status) {
    __shared__ cplx tab[3][kLM], fwd[3][kLM];
    if (threadIdx.x == 0) {
        cplx in[4];
        u.interior(in[0], in[1], in[2], in[3]);
        // forward sweep with x_first as a symbol: dp_i = P_i x_first + G_i;  backward with x_last as a symbol
        cplx cp = czero(), Pp = cone(), Gp = czero();
        for (int i = 1; i <= kLM - 2; ++i) {
            const cplx inv = frecip(nmsub(in[1], in[0], cp));
            Pp = -((in[0] * Pp) * inv);
            Gp = nmsub(in[3], in[0], Gp) * inv;
            cp = in[2] * inv;
            fwd[0][i] = cp, fwd[1][i] = Pp, fwd[2][i] = Gp;
        }
        cplx al = czero(), be = cone(), ga = czero();
        tab[0][kLM - 1] = al, tab[1][kLM - 1] = be, tab[2][kLM - 1] = ga;  // x_15 = x_last
        for (int i = kLM - 2; i >= 1; --i) {
            al = nmsub(fwd[1][i], fwd[0][i], al);
            be = -(fwd[0][i] * be);
            ga = nmsub(fwd[2][i], fwd[0][i], ga);
            tab[0][i] = al, tab[1][i] = be, tab[2][i] = ga;
        }
        tab[0][0] = cone(), tab[1][0] = czero(), tab[2][0] = czero();       // x_0 = x_first
    }
    __syncthreads();
    const int i = threadIdx.x & (kLM - 1);
    const cplx al = tab[0][i], be = tab[1][i], ga = tab[2][i];
    const UniformSpan sp = uniform_span(u.g_off, u.n);
    const int64_t c_begin = sp.c_lo + (int64_t)blockIdx.x * chunks_per_block;
    int64_t c_end = c_begin + chunks_per_block;
    if (c_end > sp.c_hi) c_end = sp.c_hi;
    bool bad = false;
    constexpr int kStep = kUniThreads / kLM, kFly = 8;  // 8 passes in flight: the end-value loads overlap
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
                bad |= !cfinite(xi);
                stg_stream(x + c * kLM + i, xi);
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0 && status != nullptr) atomicOr(status, 1);
}

// ... and the special chunks: Thomas sweep with generated rows, one thread per chunk
__global__ void uniform_special_backsub_kernel(UniformRows u, const cplx *__restrict__ xr, cplx *__restrict__ x,
                                               int32_t *__restrict__ status) {
    const int64_t c = special_chunk(u, (int)threadIdx.x);
    if (c < 0) return;
    const ChunkPos cp = chunk_pos(u.n, tile_count(u.n), c);
    if (!cp.exists) return;
    cplx in[4];
    u.interior(in[0], in[1], in[2], in[3]);
    const int64_t g0 = u.g_off + cp.r0;
    const cplx x0 = xr[2 * c], xl = cp.len > 1 ? xr[2 * c + 1] : x0;
    bool bad = !cfinite(x0) || !cfinite(xl);
    cplx cs[kLM + 1], ds[kLM + 1];
    cplx cpv = czero(), dpv = x0;
    for (int i = 1; i <= cp.len - 2; ++i) {
        cplx a, b, cc, d;
        u.row(g0 + i, in, a, b, cc, d);
        const cplx inv = frecip(nmsub(b, a, cpv));
        dpv = nmsub(d, a, dpv) * inv;
        cpv = cc * inv;
        cs[i] = cpv, ds[i] = dpv;
    }
    cplx xn = xl;
    if (cp.len > 1) x[cp.r0 + cp.len - 1] = xl;
    for (int i = cp.len - 2; i >= 1; --i) {
        xn = nmsub(ds[i], cs[i], xn);
        bad |= !cfinite(xn);
        x[cp.r0 + i] = xn;
    }
    x[cp.r0] = x0;
    if (bad && status != nullptr) atomicOr(status, 1);
}

// Tiny dense solve of the gathered interface system (2P rows, P = number of ranks): one thread,
// Gaussian elimination with partial pivoting.  iface: [P][4][2] = per rank (a, b, c, d) x (first, last).
constexpr int kIfaceMax = 32;  // up to 16 ranks (one NVSwitch domain)
__global__ void spike_interface_kernel(const cplx *__restrict__ iface, int P, cplx *__restrict__ x,
                                       int32_t *__restrict__ status) {
    __shared__ cplx A[kIfaceMax][kIfaceMax + 1];
    const int n = 2 * P;
    for (int i = threadIdx.x; i < n * (n + 1); i += blockDim.x) A[i / (n + 1)][i % (n + 1)] = czero();
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int p = 0; p < P; ++p)
        for (int r = 0; r < 2; ++r) {
            const int i = 2 * p + r;
            const cplx *q = iface + (size_t)p * 8;
            if (i > 0) A[i][i - 1] = q[0 + r];
            A[i][i] = q[2 + r];
            if (i < n - 1) A[i][i + 1] = q[4 + r];
            A[i][n] = q[6 + r];
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
        cplx s = A[k][n];
        for (int j = k + 1; j < min(n, k + 4); ++j) s = nmsub(s, A[k][j], x[j]);
        x[k] = s * crecip(A[k][k]);
        bad |= !cfinite(x[k]);
    }
    if (bad && status != nullptr) atomicOr(status, 1);
}

// ---------------------------------------------------------------------------------------------- host side
struct LevelPlan {
    int levels;                     // number of reduced levels (level 0 is the caller's system)
    int64_t n[kMaxLevels + 1];      // rows per level
    size_t off[kMaxLevels + 1];     // byte offset of level l >= 1 inside the workspace (5 arrays of n[l])
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
    return pl;
}

struct LevelPtrs {
    cplx *a, *b, *c, *d, *x;
};
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

template <class Source>
static int launch_reduce(const Source &src, const LevelPtrs &out, cudaStream_t st) {
    auto kern = large_reduce_kernel<Source>;
    FH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LargeSmem)));
    kern<<<grid_for(src.n), kLT, sizeof(LargeSmem), st>>>(src, out.a, out.b, out.c, out.d);
    FH_LAUNCH_CHECK("large_reduce_kernel");
    return FH_OK;
}
template <class Source>
static int launch_backsub(const Source &src, const cplx *xr, cplx *x, int32_t *status, cudaStream_t st) {
    auto kern = large_backsub_kernel<Source>;
    FH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LargeSmem)));
    kern<<<grid_for(src.n), kLT, sizeof(LargeSmem), st>>>(src, xr, x, status);
    FH_LAUNCH_CHECK("large_backsub_kernel");
    return FH_OK;
}

// generated rows on a uniform mesh: the structured level-0 kernels
static bool uniform_rows(const HelmRows &src) { return src.p.h_mode == FH_H_SCALAR && src.p.n_elem >= 2; }
static UniformRows as_uniform(const HelmRows &src) {
    UniformRows u;
    u.p = src.p, u.k = src.k, u.k2 = src.k2, u.g_off = src.g_off, u.n = src.n;
    return u;
}
// contiguous ranges: `units` pieces of work over at most 8 CTAs per SM; returns the grid, sets units per CTA (a
// multiple of `granule`)
static unsigned uniform_grid(int64_t units, int64_t granule, int64_t *per_block) {
    const int64_t want = 8ll * sm_count();
    int64_t per = (units + want - 1) / want;
    per = ((per + granule - 1) / granule) * granule;
    if (per < granule) per = granule;
    *per_block = per;
    const int64_t blocks = (units + per - 1) / per;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}
static int launch_reduce(const HelmRows &src, const LevelPtrs &out, cudaStream_t st) {
    if (!uniform_rows(src)) return launch_reduce<HelmRows>(src, out, st);
    const UniformRows u = as_uniform(src);
    const UniformSpan sp = uniform_span(u.g_off, u.n);
    if (sp.c_hi > sp.c_lo) {
        int64_t per;
        const unsigned grid = uniform_grid(2 * (sp.c_hi - sp.c_lo), 2 * kUniThreads, &per);
        uniform_reduce_kernel<<<grid, kUniThreads, 0, st>>>(u, per, out.a, out.b, out.c, out.d);
    }
    uniform_special_reduce_kernel<<<1, 96, 0, st>>>(u, out.a, out.b, out.c, out.d);
    FH_LAUNCH_CHECK("uniform_reduce_kernel");
    return FH_OK;
}
static int launch_backsub(const HelmRows &src, const cplx *xr, cplx *x, int32_t *status, cudaStream_t st) {
    if (!uniform_rows(src)) return launch_backsub<HelmRows>(src, xr, x, status, st);
    const UniformRows u = as_uniform(src);
    const UniformSpan sp = uniform_span(u.g_off, u.n);
    if (sp.c_hi > sp.c_lo) {
        int64_t per;
        const unsigned grid = uniform_grid(sp.c_hi - sp.c_lo, kUniThreads / kLM, &per);
        uniform_backsub_kernel<<<grid, kUniThreads, 0, st>>>(u, per, xr, x, status);
    }
    uniform_special_backsub_kernel<<<1, 96, 0, st>>>(u, xr, x, status);
    FH_LAUNCH_CHECK("uniform_backsub_kernel");
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
        int rc = l == 0 ? launch_reduce(src0, out, st)
                        : launch_reduce(bands_of(level_ptrs(ws, pl, l), pl.n[l], open_first, open_last), out, st);
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
        int rc = l == 0 ? launch_backsub(src0, xr, u, status, st)
                        : launch_backsub(bands_of(level_ptrs(ws, pl, l), pl.n[l], open_first, open_last), xr,
                                         level_ptrs(ws, pl, l).x, status, st);
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
    int32_t *coarse_status = nullptr;  // the batched kernel zeroes status[] itself; keep the OR semantics
    rc = batched_solve_bands(top.a, top.b, top.c, top.d, top.x, pl.n[pl.levels], 1, coarse_status, st);
    if (rc != FH_OK) return rc;
    return backsub_all(src0, pl, ws, 0, 0, u, status, st);
}

size_t solve_large_workspace_bytes(int64_t n) { return plan_levels(n, kCoarseMax).total; }

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

}  // namespace fh

extern "C" {

size_t fh_tridiag_solve_large_workspace_bytes(int64_t n) { return fh::solve_large_workspace_bytes(n); }

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

int fh_helmholtz1d_solve_large_fused_c128(const fh_problem1d *prob_host, double k, double k2, fh_c128 *u,
                                          int32_t *status, void *workspace, size_t workspace_bytes,
                                          fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host != nullptr && u != nullptr, "NULL pointer");
    HelmRows src;
    int rc = helm_rows_from(prob_host, k, k2, 0, prob_host->n_elem + 1, src);
    if (rc != FH_OK) return rc;
    return solve_large_any(src, reinterpret_cast<cplx *>(u), src.n, status, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

// ---- multi-GPU SPIKE pieces: every rank owns the contiguous rows [row_begin, row_begin + n_rows) ------------
size_t fh_spike_workspace_bytes(int64_t n_rows) { return fh::plan_levels(n_rows, 2).total; }

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

int fh_helmholtz1d_spike_reduce_c128(const fh_problem1d *prob_host, double k, double k2, int64_t row_begin,
                                     int64_t n_rows, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                                     fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host != nullptr && iface8 != nullptr, "NULL pointer");
    FH_REQUIRE(n_rows > 2, "a generated slab needs more than 2 rows");
    HelmRows src;
    int rc = helm_rows_from(prob_host, k, k2, row_begin, n_rows, src);
    if (rc != FH_OK) return rc;
    const LevelPlan pl = plan_levels(n_rows, 2);
    if (workspace_bytes < pl.total || !workspace) {
        set_error("workspace too small: need %zu bytes", pl.total);
        return FH_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = reduce_all(src, pl, workspace, 1, 1, st);  // generated rows carry exact zeros at the mesh ends
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

int fh_helmholtz1d_spike_backsub_c128(const fh_problem1d *prob_host, double k, double k2, int64_t row_begin,
                                      int64_t n_rows, void *workspace, size_t workspace_bytes,
                                      const fh_c128 *x_first_last, fh_c128 *u, int32_t *status, fh_stream_t stream) {
    using namespace fh;
    FH_REQUIRE(prob_host && x_first_last && u, "NULL pointer");
    HelmRows src;
    int rc = helm_rows_from(prob_host, k, k2, row_begin, n_rows, src);
    if (rc != FH_OK) return rc;
    const LevelPlan pl = plan_levels(n_rows, 2);
    FH_REQUIRE(workspace_bytes >= pl.total && pl.levels > 0, "workspace too small: need %zu bytes", pl.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = spike_seed_top(pl, workspace, x_first_last, st);
    if (rc != FH_OK) return rc;
    return backsub_all(src, pl, workspace, 1, 1, reinterpret_cast<cplx *>(u), status, st);
}

}  // extern "C"
