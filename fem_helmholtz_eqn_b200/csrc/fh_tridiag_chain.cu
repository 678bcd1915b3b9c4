// Batched complex128 tridiagonal solve, "chain" variant: one system per cluster of C = 4 CTAs of 128 threads,
// 1024 rows per CTA, TWO CTAs resident per SM (101 KB of shared memory, <= 255 registers x 128 threads each).
//
// Why: the 2-CTA x 256-thread kernel (fh_tridiag_batched.cu) holds a whole SM with one CTA of 8 warps that
// march through the phases in lock-step; it is latency-bound (issue slots 34 % busy).  Two independent CTAs per
// SM overlap one CTA's dependent chains and barriers with the other's.
//
// All CTAs process their rows in forward order, so the cluster is a plain chain of 4 x 128 chunks.  Inside a
// CTA the algorithm is the one of fh_tridiag_batched.cu (8 rows per thread in registers, elimination to one
// interface row per thread, PCR over the CTA's 128 interface rows).  Across CTAs:
//   * pre-exchange : the last interface row of CTA q needs the normalised first row of CTA q+1's first chunk
//                    (3 complex numbers, st.async into q's shared memory, mbarrier complete_tx);
//   * PCR          : every CTA carries two extra right-hand sides, V (coupling of its last interface row to the
//                    next CTA's first interface unknown e_{q+1}) and W (coupling of its first interface row to
//                    the previous CTA's last interface unknown p_{q-1}):  X_t = Y_t - V_t e_{q+1} - W_t p_{q-1};
//   * post-exchange: every CTA broadcasts (Y, V, W) of its first and last interface row (6 numbers) to the other
//                    CTAs; each CTA then solves the 2C-unknown interface chain redundantly (a few dozen complex
//                    operations) and finishes locally.
// Data movement, accuracy probe and status flags are as in the 2-CTA kernel (TMA tensor tiles of 128 chunks = 16 KB).
#include <cooperative_groups.h>

#include "fh_tridiag_dev.cuh"

namespace cg = cooperative_groups;

namespace fh {

constexpr int cT = 128;                  // threads per CTA
constexpr int cM = 8;                    // rows per thread
constexpr int cSlots = cT * cM;          // 1024 rows per CTA
constexpr int cHeadMax = 8;
constexpr int cLevels = 7;               // log2(cT)
constexpr int cMaxCluster = 4;
constexpr uint32_t cTileBytes = cSlots * sizeof(cplx);  // 16 KB

struct __align__(1024) ChainSmem {
    cplx stage[4][cSlots];              // TMA destination: bands of the NEXT system
    cplx xs[cSlots];                    // TMA source: solution of the CURRENT system
    cplx pcr[2][5][cT];                 // PCR ping-pong: A, C, D, V, W
    cplx row0[3][cT];                   // normalised first row of every chunk: A0, C0, D0
    cplx headst[4][cHeadMax];           // head rows (n mod 8) of the next system, CTA 0 only
    cplx pre[2][3];                     // [parity]: A0, C0, D0 of the next CTA's first chunk
    cplx post[2][cMaxCluster][6];       // [parity][rank]: Y, V, W of the first and of the last interface row
    unsigned long long full_bar, pre_bar[2], post_bar[2];
};

struct ChainGeom {      // how a system of n rows maps onto the cluster (computed on the host)
    int head;           // n mod 8: rows in front of the first chunk, eliminated serially by one thread of CTA 0
    int chunks_total;   // n / 8
    int pad;            // leading chunk slots of CTA 0 that stay empty (identity rows): C * 128 - chunks_total
};

struct ChainStoreMaps {
    CUtensorMap first, rest;  // CTA 0 (left-aligned, clipped to its chunks) and the other CTAs
};

template <int C, class Source>
__global__ void __launch_bounds__(cT, 2)
tridiag_chain_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ ChainStoreMaps umaps, Source src,
                     cplx *__restrict__ u, int64_t n, int64_t batch, ChainGeom geo, int32_t *__restrict__ status) {
    extern __shared__ unsigned char smem_raw[];
    ChainSmem &sm = *reinterpret_cast<ChainSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    const int t = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int64_t first_sys = blockIdx.x / C, sys_stride = gridDim.x / C;

    const int mc = rank * cT + t - geo.pad;  // this thread's chunk index in memory order (< 0: empty slot)
    const bool active = mc >= 0;
    const int box_c1 = rank * cT - geo.pad;  // first chunk coordinate of this CTA's tile
    const int feeder = rank == 0 ? geo.pad : 0;      // issues the loads; in CTA 0 it also owns the head rows
    const bool has_head = rank == 0 && geo.head > 0;
    auto xs_slot = [&](int tt, int i) {
        const int r = rank == 0 ? tt - geo.pad : tt;  // CTA 0 stores left-aligned (TMA stores reject negative coords)
        return r * cM + (i ^ (r & 7));
    };

    if (t == 0) {
        mbar_init(&sm.full_bar, 1);
        for (int p = 0; p < 2; ++p) mbar_init(&sm.pre_bar[p], 1), mbar_init(&sm.post_bar[p], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();

    HelmState hs;
    if constexpr (!Source::kStaged) {
        if (src.p.h_mode == FH_H_SCALAR) {
            hs.g_first = row_geom(src.p, 0);
            hs.g_mid = row_geom(src.p, src.p.n_elem > 1 ? 1 : 0);
            hs.g_last = row_geom(src.p, src.p.n_elem);
        }
    }
    auto gen_row = [&](int64_t g, cplx &a, cplx &b, cplx &c, cplx &d) {
        if constexpr (!Source::kStaged) {
            if (src.p.h_mode == FH_H_SCALAR) {
                RowGeom rg = hs.g_mid;
                if (g == 0) rg = hs.g_first;
                if (g == src.p.n_elem) rg = hs.g_last;
                row_values(src.p, rg, g, hs.k, hs.k2, a, b, c, d);
            } else {
                row_values(src.p, row_geom(src.p, g), g, hs.k, hs.k2, a, b, c, d);
            }
        }
    };

    auto prefetch = [&](int64_t sys) {
        if constexpr (Source::kStaged) {
            if (t == feeder) {
                mbar_expect_tx(&sm.full_bar, 4 * cTileBytes);
                tma_load_3d(sm.stage[0], &maps.dl, 0, box_c1, (int)sys, &sm.full_bar);
                tma_load_3d(sm.stage[1], &maps.d, 0, box_c1, (int)sys, &sm.full_bar);
                tma_load_3d(sm.stage[2], &maps.du, 0, box_c1, (int)sys, &sm.full_bar);
                tma_load_3d(sm.stage[3], &maps.b, 0, box_c1, (int)sys, &sm.full_bar);
                if (has_head) {
                    for (int j = 0; j < geo.head; ++j) {
                        const int64_t g = sys * n + j;
                        cp_async16(&sm.headst[0][j], src.dl + g);
                        cp_async16(&sm.headst[1][j], src.d + g);
                        cp_async16(&sm.headst[2][j], src.du + g);
                        cp_async16(&sm.headst[3][j], src.b + g);
                    }
                    cp_async_commit();
                }
            }
        }
    };
    if (first_sys < batch) prefetch(first_sys);

    uint32_t it = 0;
    for (int64_t sys = first_sys; sys < batch; sys += sys_stride, ++it) {
        const uint32_t parity = it & 1, phase = (it >> 1) & 1;
        if constexpr (!Source::kStaged) hs.k = __ldg(src.ks + sys), hs.k2 = __ldg(src.k2s + sys);
        if (t == 0) {
            if (rank < C - 1) mbar_expect_tx(&sm.pre_bar[parity], 3 * sizeof(cplx));
            mbar_expect_tx(&sm.post_bar[parity], (C - 1) * 6 * sizeof(cplx));
        }

        // ---- bands -> registers ---------------------------------------------------------------------
        cplx a[cM], b[cM], c[cM], d[cM];
        if constexpr (Source::kStaged) {
            mbar_wait(&sm.full_bar, parity);
            if (active) {
#pragma unroll
                for (int i = 0; i < cM; ++i) {
                    const int s = t * cM + (i ^ (t & 7));
                    a[i] = sm.stage[0][s], b[i] = sm.stage[1][s], c[i] = sm.stage[2][s], d[i] = sm.stage[3][s];
                }
            }
        } else {
            if (active) {
#pragma unroll
                for (int i = 0; i < cM; ++i) gen_row((int64_t)geo.head + (int64_t)mc * cM + i, a[i], b[i], c[i], d[i]);
            }
        }
        if (!active) {
#pragma unroll
            for (int i = 0; i < cM; ++i) a[i] = czero(), b[i] = cone(), c[i] = czero(), d[i] = czero();
        }
        if (rank == 0 && geo.head == 0 && t == geo.pad) a[0] = czero();  // first row of the system
        if (rank == C - 1 && t == cT - 1) c[cM - 1] = czero();            // last row of the system

        // ---- head rows (n mod 8, in front of the first chunk): serial Thomas, folded into row 0 ----------
        cplx hcp[cHeadMax], hdp[cHeadMax];
        if (has_head && t == feeder) {
            if constexpr (Source::kStaged) cp_async_wait_all();
            cplx cp = czero(), dp = czero();
#pragma unroll
            for (int j = 0; j < cHeadMax; ++j) {
                if (j < geo.head) {
                    cplx ha, hb, hc, hd;
                    if constexpr (Source::kStaged) {
                        ha = sm.headst[0][j], hb = sm.headst[1][j], hc = sm.headst[2][j], hd = sm.headst[3][j];
                    } else {
                        gen_row(j, ha, hb, hc, hd);
                    }
                    if (j == 0) ha = czero();
                    const cplx inv = frecip(nmsub(hb, ha, cp));
                    dp = nmsub(hd, ha, dp) * inv;
                    cp = hc * inv;
                    hcp[j] = cp, hdp[j] = dp;
                }
            }
            b[0] = nmsub(b[0], a[0], cp);
            d[0] = nmsub(d[0], a[0], dp);
            a[0] = czero();
        }
        if constexpr (Source::kStaged) {
            __syncthreads();  // every thread has drained the staging tiles ...
            if (sys + sys_stride < batch) prefetch(sys + sys_stride);  // ... refill them
        }

        // ---- chunk elimination (as in the 2-CTA kernel): b[i] (i < 7) becomes 1 / pivot --------------------
        const cplx a7o = a[cM - 1], b7o = b[cM - 1], c7o = c[cM - 1], d7o = d[cM - 1];  // for the probe
        b[0] = frecip(b[0]);
#pragma unroll
        for (int i = 1; i < cM; ++i) {
            const cplx w = a[i] * b[i - 1];
            b[i] = nmsub(b[i], w, c[i - 1]);
            d[i] = nmsub(d[i], w, d[i - 1]);
            a[i] = -(w * a[i - 1]);
            if (i < cM - 1) b[i] = frecip(b[i]);
        }
#pragma unroll
        for (int i = cM - 3; i >= 0; --i) {
            const cplx w = c[i] * b[i + 1];
            a[i] = nmsub(a[i], w, a[i + 1]);
            d[i] = nmsub(d[i], w, d[i + 1]);
            c[i] = -(w * c[i + 1]);
        }
        // ---- normalised first row of the chunk; CTA q > 0 ships its thread-0 copy to CTA q - 1 ---------------
        {
            const cplx A0 = a[0] * b[0], C0 = c[0] * b[0], D0 = d[0] * b[0];
            sm.row0[0][t] = A0, sm.row0[1][t] = C0, sm.row0[2][t] = D0;
            if (t == 0 && rank > 0) {
                const uint32_t rbar = map_to_rank(&sm.pre_bar[parity], rank - 1);
                st_async_remote(map_to_rank(&sm.pre[parity][0], rank - 1), A0, rbar);
                st_async_remote(map_to_rank(&sm.pre[parity][1], rank - 1), C0, rbar);
                st_async_remote(map_to_rank(&sm.pre[parity][2], rank - 1), D0, rbar);
            }
        }
        __syncthreads();
        // ---- interface row of this chunk ----------------------------------------------------------------------
        cplx rA, rC, rD, rV = czero(), rW = czero();
        cplx nA = czero(), nC = czero(), nD = czero();  // normalised first row of the NEXT chunk
        if (t < cT - 1) {
            nA = sm.row0[0][t + 1], nC = sm.row0[1][t + 1], nD = sm.row0[2][t + 1];
        } else if (rank < C - 1) {
            mbar_wait(&sm.pre_bar[parity], phase);
            nA = sm.pre[parity][0], nC = sm.pre[parity][1], nD = sm.pre[parity][2];
        }
        {
            const cplx rb = nmsub(b[cM - 1], c[cM - 1], nA);
            const cplx cx = -(c[cM - 1] * nC);
            const cplx inv = frecip(rb);
            rA = a[cM - 1] * inv;
            rD = nmsub(d[cM - 1], c[cM - 1], nD) * inv;
            rC = czero();
            if (t < cT - 1) rC = cx * inv;
            else rV = cx * inv;                  // couples to e_{rank+1}
            if (t == 0) rW = rA, rA = czero();   // couples to p_{rank-1} (zero in CTA 0)
        }
        // ---- PCR over the CTA's 128 interface rows, three right-hand sides ---------------------------------------
        int cur = 0;
        sm.pcr[0][0][t] = rA, sm.pcr[0][1][t] = rC, sm.pcr[0][2][t] = rD, sm.pcr[0][3][t] = rV, sm.pcr[0][4][t] = rW;
        __syncthreads();
#pragma unroll 1
        for (int lv = 0, s = 1; lv < cLevels; ++lv, s <<= 1) {
            const bool last = lv == cLevels - 1;
            cplx Am = czero(), Cm = czero(), Dm = czero(), Vm = czero(), Wm = czero();
            cplx Ap = czero(), Cp = czero(), Dp = czero(), Vp = czero(), Wp = czero();
            if (t - s >= 0) {
                Am = sm.pcr[cur][0][t - s], Cm = sm.pcr[cur][1][t - s], Dm = sm.pcr[cur][2][t - s];
                Vm = sm.pcr[cur][3][t - s], Wm = sm.pcr[cur][4][t - s];
            }
            if (t + s < cT) {
                Ap = sm.pcr[cur][0][t + s], Cp = sm.pcr[cur][1][t + s], Dp = sm.pcr[cur][2][t + s];
                Vp = sm.pcr[cur][3][t + s], Wp = sm.pcr[cur][4][t + s];
            }
            const cplx inv = frecip(nmsub(nmsub(cone(), rA, Cm), rC, Ap));
            const cplx nDv = nmsub(nmsub(rD, rA, Dm), rC, Dp) * inv;
            const cplx nV = nmsub(nmsub(rV, rA, Vm), rC, Vp) * inv;
            const cplx nW = nmsub(nmsub(rW, rA, Wm), rC, Wp) * inv;
            cur ^= 1;
            if (!last) {
                const cplx nAv = -((rA * Am) * inv), nCv = -((rC * Cp) * inv);
                rA = nAv, rC = nCv;
                sm.pcr[cur][0][t] = rA, sm.pcr[cur][1][t] = rC;
            }
            rD = nDv, rV = nV, rW = nW;
            if (!last) sm.pcr[cur][2][t] = rD, sm.pcr[cur][3][t] = rV, sm.pcr[cur][4][t] = rW;
            if (!last) __syncthreads();
        }
        // now X_t = rD - rV * e_{rank+1} - rW * p_{rank-1}
        // ---- post-exchange: (Y, V, W) of the first and the last interface row go to every CTA -----------------
        if (t == 0 || t == cT - 1) {
            const int o = t == 0 ? 0 : 3;
            sm.post[parity][rank][o + 0] = rD, sm.post[parity][rank][o + 1] = rV, sm.post[parity][rank][o + 2] = rW;
#pragma unroll
            for (int q = 0; q < C; ++q) {
                if (q == rank) continue;
                const uint32_t rbar = map_to_rank(&sm.post_bar[parity], q);
                st_async_remote(map_to_rank(&sm.post[parity][rank][o + 0], q), rD, rbar);
                st_async_remote(map_to_rank(&sm.post[parity][rank][o + 1], q), rV, rbar);
                st_async_remote(map_to_rank(&sm.post[parity][rank][o + 2], q), rW, rbar);
            }
        }
        __syncthreads();  // own entries of post[] are visible
        mbar_wait(&sm.post_bar[parity], phase);
        // ---- interface chain, solved redundantly by every thread: e_q / p_q = first / last interface unknown of CTA q
        //   e_q = Ye_q - Ve_q e_{q+1} - We_q p_{q-1},   p_q = Yp_q - Vp_q e_{q+1} - Wp_q p_{q-1}
        cplx x_next = czero(), x_prev = czero();  // e_{rank+1}, p_{rank-1}
        {
            cplx alpha[C], beta[C], gamma[C], delta[C], e[C + 1];
            alpha[0] = sm.post[parity][0][3], beta[0] = sm.post[parity][0][4];  // p_0 = alpha - beta e_1
            gamma[0] = czero(), delta[0] = czero();
#pragma unroll
            for (int q = 1; q < C; ++q) {
                const cplx Ye = sm.post[parity][q][0], Ve = sm.post[parity][q][1], We = sm.post[parity][q][2];
                const cplx Yp = sm.post[parity][q][3], Vp = sm.post[parity][q][4], Wp = sm.post[parity][q][5];
                const cplx inv = frecip(nmsub(cone(), We, beta[q - 1]));
                gamma[q] = nmsub(Ye, We, alpha[q - 1]) * inv;  // e_q = gamma - delta e_{q+1}
                delta[q] = Ve * inv;
                const cplx wb = Wp * beta[q - 1];
                alpha[q] = nmsub(Yp, Wp, alpha[q - 1]) + wb * gamma[q];  // p_q = alpha - beta e_{q+1}
                beta[q] = Vp + wb * delta[q];
            }
            e[C] = czero();
#pragma unroll
            for (int q = C - 1; q >= 1; --q) e[q] = nmsub(gamma[q], delta[q], e[q + 1]);
#pragma unroll
            for (int q = 0; q < C; ++q) {
                if (q == rank) x_next = e[q + 1];
                if (q == rank - 1) x_prev = nmsub(alpha[q], beta[q], e[q + 1]);
            }
        }
        const cplx X = nmsub(nmsub(rD, rV, x_next), rW, x_prev);
        sm.pcr[cur][0][t] = X;  // (the last PCR level left pcr[cur] unwritten)
        if (t == 0) tma_store_wait_read();  // the previous system's TMA store has finished reading xs
        __syncthreads();
        const cplx xl = t > 0 ? sm.pcr[cur][0][t - 1] : x_prev;

        // ---- back-substitution + staging -------------------------------------------------------------------------
        bool bad = !cfinite(X);
        cplx x0 = X, x6 = X;
#pragma unroll
        for (int i = 0; i < cM; ++i) {
            cplx xi = X;
            if (i < cM - 1) xi = nmsub(nmsub(d[i], a[i], xl), c[i], X) * b[i];
            if (i == 0) x0 = xi;
            if (i == cM - 2) x6 = xi;
            bad |= !cfinite(xi);
            if (active) sm.xs[xs_slot(t, i)] = xi;
        }
        if (has_head && t == feeder) {
            cplx xn = x0;
            cplx *out = u + sys * n;
#pragma unroll
            for (int j = cHeadMax - 1; j >= 0; --j) {
                if (j < geo.head) {
                    xn = nmsub(hdp[j], hcp[j], xn);
                    bad |= !cfinite(xn);
                    stg_stream(out + j, xn);
                }
            }
        }
        fence_proxy_async();
        const int any_nonfinite = __syncthreads_or(bad ? 1 : 0);
        if (t == 0) {
            tma_store_3d(rank == 0 ? &umaps.first : &umaps.rest, 0, rank == 0 ? 0 : box_c1, (int)sys, sm.xs);
            tma_store_commit();
            if (status != nullptr && any_nonfinite) atomicOr(&status[sys], 1);
        }
        {   // probe: defect of the original interface row with the computed solution
            cplx xn = czero();
            if (active && t < cT - 1) xn = sm.xs[xs_slot(t + 1, 0)];
            if (t == cT - 1 && rank < C - 1) xn = nmsub(nmsub(nD, nA, X), nC, x_next);  // first row of the next CTA
            const cplx r = nmsub(nmsub(nmsub(d7o, a7o, x6), b7o, X), c7o, xn);
            const double scale = abs1(a7o) * abs1(x6) + abs1(b7o) * abs1(X) + abs1(c7o) * abs1(xn) + abs1(d7o);
            const bool off = active && abs1(r) > kProbeTol * scale;
            if (__any_sync(0xffffffffu, off) && (t & 31) == 0 && status != nullptr) atomicOr(&status[sys], 2);
        }
    }
    if (t == 0) tma_store_wait_read();
    cluster.sync();  // no CTA leaves while another may still address its shared memory
}

// ---------------------------------------------------------------------------------------------- host side
template <int C>
static bool chain_supports(int64_t n, ChainGeom &g) {
    g.head = (int)(n % cM);
    g.chunks_total = (int)(n / cM);
    g.pad = C * cT - g.chunks_total;
    return g.chunks_total > (C - 1) * cT && g.chunks_total <= C * cT;
}

template <int C, class Source>
static int launch_chain(const Source &src, const cplx *const bands[4], cplx *u, int64_t n, int64_t batch,
                        int32_t *status, cudaStream_t stream) {
    ChainGeom geo;
    if (!chain_supports<C>(n, geo)) {
        set_error("chain kernel does not cover n = %lld", (long long)n);
        return FH_E_UNSUPPORTED;
    }
    auto kernel = tridiag_chain_kernel<C, Source>;
    const size_t smem = sizeof(ChainSmem) + 1024;
    TensorMaps maps;
    ChainStoreMaps umaps;
    memset(&maps, 0, sizeof(maps));
    memset(&umaps, 0, sizeof(umaps));
    int rc = make_tensor_map(&umaps.first, u, n, batch, geo.head, cT - geo.pad, cT);
    if (rc != FH_OK) return rc;
    rc = make_tensor_map(&umaps.rest, u, n, batch, geo.head, geo.chunks_total, cT);
    if (rc != FH_OK) return rc;
    if (Source::kStaged) {
        CUtensorMap *dst[4] = {&maps.dl, &maps.d, &maps.du, &maps.b};
        for (int i = 0; i < 4; ++i) {
            rc = make_tensor_map(dst[i], bands[i], n, batch, geo.head, geo.chunks_total, cT);
            if (rc != FH_OK) return rc;
        }
    }
    FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (status != nullptr) FH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * batch, stream));
    int64_t groups = (2ll * sm_count()) / C;  // two CTAs per SM
    if (groups < 1) groups = 1;
    if (groups > batch) groups = batch;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * C));
    cfg.blockDim = dim3(cT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FH_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, umaps, src, u, n, batch, geo, status));
    return FH_OK;
}

bool chain_covers(int64_t n) {
    ChainGeom g;
    return chain_supports<4>(n, g);
}

int chain_solve_bands(const cplx *dl, const cplx *d, const cplx *du, const cplx *b, cplx *u, int64_t n, int64_t batch,
                      int32_t *status, cudaStream_t stream) {
    BandSource src;
    src.dl = dl, src.d = d, src.du = du, src.b = b;
    const cplx *const bands[4] = {dl, d, du, b};
    return launch_chain<4>(src, bands, u, n, batch, status, stream);
}

int chain_solve_helm(const HelmSource &src, cplx *u, int64_t n, int64_t batch, int32_t *status, cudaStream_t stream) {
    const cplx *const none[4] = {nullptr, nullptr, nullptr, nullptr};
    return launch_chain<4>(src, none, u, n, batch, status, stream);
}

}  // namespace fh
