/*
 * fem_helmholtz.h -- C ABI of the B200-native Helmholtz FEM hot path
 * (libfemhelmholtz_b200.so, built from fem_helmholtz_eqn_b200/csrc/ for sm_100a).
 *
 * The reference (anishasahu/fem-helmholtz-eqn) has no FFI layer: its seam is a Python call
 * surface.  Each entry point below names the reference interface it replaces
 * (nb:N = raw JSON line N of notebooks/1D_soln.ipynb; other paths are under src/).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.  The caller owns every
 *     buffer; the library allocates nothing persistent.  Scratch space is passed in and sized by
 *     the matching *_workspace_bytes() query.
 *   - complex128 is stored as interleaved (re, im) doubles, identical to numpy.complex128 /
 *     torch.complex128 / cuDoubleComplex.  Complex scalars cross the ABI as two doubles.
 *   - Every call is asynchronous and ordered on `stream` (a cudaStream_t passed as void*).
 *   - Return value: 0 = ok, < 0 = invalid argument (FH_E_*), > 0 = a cudaError_t.
 *     fh_last_error() returns a thread-local description of the last failure.
 *   - No exceptions cross the boundary; the library keeps no global mutable state besides the
 *     thread-local error string and one-time per-device kernel attribute setup, and it never allocates
 *     device memory (not even stream-ordered scratch).
 *   - Batched 1D arrays are batch-major: entry (system s, row j) lives at [s * n + j].
 *     Tridiagonal bands use n entries per system: dl[s][0] and du[s][n-1] are ignored.
 */
#ifndef FEM_HELMHOLTZ_H
#define FEM_HELMHOLTZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FH_ABI_VERSION 1

#if defined(__GNUC__)
#define FH_API __attribute__((visibility("default")))
#else
#define FH_API
#endif

typedef void *fh_stream_t;            /* cudaStream_t */
typedef struct { double re, im; } fh_c128;

enum { FH_OK = 0, FH_E_INVALID = -1, FH_E_UNSUPPORTED = -2, FH_E_WORKSPACE = -3 };

/* Boundary rows of the 1D problem.
 *   NEUMANN    (left)  : load b[0] += -i k cos(theta)                       nb:113-121, nb:174-176
 *   SOMMERFELD (right) : A[N,N] += i k                                      nb:124-129, nb:179-182
 *   DIRICHLET          : row replacement A[row]=0; A[row,row]=1; b[row]=g   solvers/fem2d_dirichlet.py:20-27 */
enum { FH_BC_NEUMANN = 0, FH_BC_SOMMERFELD = 1, FH_BC_DIRICHLET = 2 };

/* Element size: SCALAR = the reference's h = L/N for every element (nb:44, nb:158);
 * NODAL = h_e = x[e+1] - x[e] from the node array (general meshes; extension). */
enum { FH_H_SCALAR = 0, FH_H_NODAL = 1 };

/* Description of one 1D Helmholtz problem family (shared by assembly and the fused solver). */
typedef struct {
    int64_t n_elem;            /* N; the system has n = N + 1 rows                                  */
    double h;                  /* scalar element size (FH_H_SCALAR)                                  */
    const double *x;           /* device fp64[N+1] node coordinates (FH_H_NODAL) or NULL             */
    int h_mode;                /* FH_H_*                                                             */
    int bc_left, bc_right;     /* FH_BC_*: left in {NEUMANN, DIRICHLET}, right in {SOMMERFELD, DIRICHLET} */
    double g_left_re, g_left_im;   /* Dirichlet value at x = 0                                       */
    double g_right_re, g_right_im; /* Dirichlet value at x = L                                       */
    double cos_theta;          /* cos(incident angle), evaluated by the caller (nb:113)              */
    double gauss_xi, gauss_w;  /* the 1-point rule the reference uses: leggauss(1) = (0, 2)  nb:52   */
} fh_problem1d;

FH_API const char *fh_last_error(void);
FH_API int fh_abi_version(void);
/* Device facts the host layer needs for grid sizing / the bench (SM count etc.). Host pointers. */
FH_API int fh_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host, size_t *smem_optin_host);

/* ---- element routines (replace nb:55-67 shape_functions, nb:70-96 element_matrices,
 *      nb:99-131 boundary_matrices); outputs are tiny device arrays, row-major -------------------- */
FH_API int fh_shape_functions_1d(double xi, double *phi2, double *dphi2, fh_stream_t stream);
FH_API int fh_element_matrices_1d(double h, double xi, double w, double *Ke4, double *Me4, fh_stream_t stream);
FH_API int fh_boundary_matrices_1d(double k, double cos_theta, fh_c128 *Ne2, fh_c128 *Se4, fh_stream_t stream);

/* ---- assembly (replaces nb:134-187 `assembly`): A = -K + k^2 M + S as tridiagonal bands ---------
 * ks / k2s: device fp64[n_k]; k2s[i] must be the caller's k**2 (nb:185 evaluates it on the host -- on a SCALAR k, i.e.
 * libm's pow(k, 2.0), which differs from the correctly rounded k*k by one ulp for about one k in a thousand: square the
 * wavenumbers of a sweep one by one, as the reference's loop nb:300-317 does, not as an array).
 * dl, d, du, b: device complex128[n_k][N+1].  Values are bit-identical to the reference's dense A.
 * fh_assemble_1d_c128 allocates nothing and needs no scratch.  On a non-uniform mesh (FH_H_NODAL) with many wavenumbers
 * fh_assemble_1d_ws_c128 is faster (0.97 instead of 0.63 of the HBM write peak): the k-independent geometry of the rows
 * (48 B per mesh row) is tabulated once in a caller-supplied workspace of fh_assemble_1d_workspace_bytes() bytes (0 =
 * no use for one: both calls then do the same). */
FH_API int fh_assemble_1d_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s,
                        int64_t n_k, fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b,
                        fh_stream_t stream);
FH_API size_t fh_assemble_1d_workspace_bytes(int64_t n_elem, int h_mode, int64_t n_k);
FH_API int fh_assemble_1d_ws_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s,
                           int64_t n_k, fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b, void *workspace,
                           size_t workspace_bytes, fh_stream_t stream);
/* One wavenumber, the mesh rows [row_begin, row_begin + n_rows) only: dl, d, du, b = device complex128[n_rows] (the slab
 * of one rank of the partitioned solve with A materialised, BASELINE config 4; same values as the full assembly).
 * k_k2_dev != NULL: (k, k**2) are read from device memory when the kernel runs (CUDA-graph replays), k / k2 are ignored. */
FH_API int fh_assemble_1d_slab_c128(const fh_problem1d *prob_host, double k, double k2, const double *k_k2_dev,
                             int64_t row_begin, int64_t n_rows, fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b,
                             fh_stream_t stream);

/* ---- batched tridiagonal solve (replaces nb:204 `np.linalg.solve(A, b)` for tridiagonal A) ------
 * Thread-chunk partition (8 rows per thread in registers) + parallel cyclic reduction of the
 * per-thread interface rows in shared memory; a 2-CTA cluster solves one system when n > 2056.
 * status[batch] (device int32, may be NULL): bit 0 (FH_STATUS_NONFINITE) set if the unpivoted elimination broke
 * down -- the solution holds a non-finite value (zero pivot / singular system) or the a-posteriori probe found a
 * relative defect above 1e-8: see fh_tridiag_repair_flagged_c128 / fh_tridiag_solve_pivoted_c128; bit 1
 * (FH_STATUS_SMALL_PIVOT) set if the probe -- every thread re-evaluates the original equation of its interface row
 * with the computed solution -- found a relative defect above 5e-14 (backward error may exceed ~1e-13: refine that
 * system, see fh_tridiag_refine_flagged_c128).  n <= fh_tridiag_batched_max_n() takes the cluster kernel, larger n
 * is routed to the partitioned large-system solver using `workspace`. */
FH_API int64_t fh_tridiag_batched_max_n(void);
FH_API size_t fh_tridiag_solve_batched_workspace_bytes(int64_t n, int64_t batch);
FH_API int fh_tridiag_solve_batched_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du,
                                  const fh_c128 *b, fh_c128 *u, int64_t n, int64_t batch,
                                  int32_t *status, void *workspace, size_t workspace_bytes,
                                  fh_stream_t stream);

/* ---- one-pass assemble + solve that KEEPS A (nb:203-204 `A, b = assembly(...)`; `u = np.linalg.solve(A, b)` for many k):
 * the rows are generated in registers, written out as the bands fh_assemble_1d_c128 would have stored (bit for bit)
 * and solved from the registers -- 64 + 16 bytes of HBM traffic per row instead of 64 + 64 + 16, A is never read
 * back.  Uniform meshes (scalar h) of 3 .. fh_tridiag_batched_max_n() nodes; FH_E_UNSUPPORTED otherwise (use the two
 * calls).  status as for fh_tridiag_solve_batched_c128; flagged systems are refined from the stored bands. */
FH_API int fh_helmholtz1d_sweep_assemble_solve_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s,
                                             int64_t n_k, fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b,
                                             fh_c128 *u, int32_t *status, void *workspace, size_t workspace_bytes,
                                             fh_stream_t stream);

/* ---- fused assemble+solve sweep (nb:190-205 `solve_helmholtz_fem` for many k at once): the bands
 * are generated in registers and never stored; only u[n_k][N+1] is written. ----------------------- */
FH_API int fh_helmholtz1d_sweep_fused_c128(const fh_problem1d *prob_host, const double *ks, const double *k2s,
                                    int64_t n_k, fh_c128 *u, int32_t *status, void *workspace,
                                    size_t workspace_bytes, fh_stream_t stream);

/* ---- one very large system (single mesh of up to 2^31-2 rows per GPU): recursive partition ------
 * every level turns each chunk of 16 rows into 2 interface rows (8x smaller system); the coarsest level
 * (<= 4112 rows) is solved by the batched kernel; back-substitution re-reads the bands (2 passes). */
FH_API size_t fh_tridiag_solve_large_workspace_bytes(int64_t n);
FH_API int fh_tridiag_solve_large_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du,
                                const fh_c128 *b, fh_c128 *u, int64_t n, int32_t *status,
                                void *workspace, size_t workspace_bytes, fh_stream_t stream);

/* Same solver with level-0 rows generated from the Helmholtz problem instead of read (bands never stored). */
FH_API int fh_helmholtz1d_solve_large_fused_c128(const fh_problem1d *prob_host, double k, double k2, fh_c128 *u,
                                          int32_t *status, void *workspace, size_t workspace_bytes,
                                          fh_stream_t stream);

/* Host-side planning of that solver on a uniform mesh (no device work; what the CPU tests check against a brute-force
 * model): for the rows [row_begin, row_begin + n_rows) of a mesh with n_elem elements, level l of the recursion has
 * out6[6 l ..] = { rows, chunks, first plain row, end of the plain rows, first plain chunk, end of the plain chunks }
 * ("plain" rows alternate between two constant rows; everything else is a "special" chunk).  Returns the number of
 * levels written (the last one has 2 rows), or FH_E_UNSUPPORTED / FH_E_INVALID (negative). */
FH_API int fh_large_uniform_plan(int64_t n_elem, int64_t row_begin, int64_t n_rows, int64_t *out6, int max_levels);

/* ... and with the wavenumber in device memory (k_k2_dev = device fp64[2] = {k, k**2}, read when the kernels run), so
 * that a captured CUDA graph of the call can be replayed for another k.  Uniform meshes only (FH_E_UNSUPPORTED
 * otherwise).  The reference's sweep loop nb:300-317 / nb:469-471 on one long mesh. */
FH_API int fh_helmholtz1d_solve_large_fused_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev, fh_c128 *u,
                                               int32_t *status, void *workspace, size_t workspace_bytes,
                                               fh_stream_t stream);

/* assembly() + np.linalg.solve() (nb:134-187, nb:204) for ONE long mesh (n_elem + 1 > fh_tridiag_batched_max_n()) in
 * one pass that KEEPS A: the first reduction level generates the rows, stores them as the bands fh_assemble_1d_c128
 * would have stored (bit for bit) and reduces them; the rest is fh_tridiag_solve_large_c128's, so dl/d/du/b and u hold
 * exactly what the two calls would have produced, for 162 instead of 226 bytes of traffic per row.
 * k_k2_dev != NULL: (k, k**2) are read from device memory (then k, k2 are ignored), so that a captured CUDA graph serves
 * every wavenumber.  workspace: fh_tridiag_solve_large_workspace_bytes(n_elem + 1).  status as fh_tridiag_solve_large_c128. */
FH_API int fh_helmholtz1d_assemble_solve_large_c128(const fh_problem1d *prob_host, double k, double k2,
                                             const double *k_k2_dev, fh_c128 *dl, fh_c128 *d, fh_c128 *du, fh_c128 *b,
                                             fh_c128 *u, int32_t *status, void *workspace, size_t workspace_bytes,
                                             fh_stream_t stream);

/* ---- multi-GPU SPIKE-style partitioned solve (single mesh split into contiguous slabs, one per rank) ------
 * Per rank:  1. *_spike_reduce   : slab [row_begin, row_begin+n_rows) -> its 2 interface rows,
 *                                  iface8 = device complex128[4][2] = (dl, d, du, b) x (first row, last row);
 *            2. all-gather iface8 over the ranks (NCCL, 128 B per rank) -> complex128[P][4][2];
 *            3. fh_spike_solve_interface_c128 : every rank solves the 2P-row interface system redundantly,
 *                                  x_iface = complex128[2P] (rank r owns entries 2r, 2r+1);
 *            4. *_spike_backsub  : slab solution from its two interface values (x_first_last = &x_iface[2r]).
 * open_first / open_last: the slab's first / last row keeps its coupling to the neighbouring rank.
 * The workspace (fh_spike_workspace_bytes) must be the same, untouched buffer for steps 1 and 4. */
FH_API size_t fh_spike_workspace_bytes(int64_t n_rows);
FH_API int fh_spike_reduce_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                         int64_t n_rows, int open_first, int open_last, void *workspace,
                         size_t workspace_bytes, fh_c128 *iface8, fh_stream_t stream);
FH_API int fh_helmholtz1d_spike_reduce_c128(const fh_problem1d *prob_host, double k, double k2, int64_t row_begin,
                                     int64_t n_rows, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                                     fh_stream_t stream);
/* Step 1 for "assemble the slab, KEEP A, solve" (nb:203-204 on a mesh split over ranks; BASELINE config 4 with A
 * materialised): the level-0 rows are generated, written out as the slab's bands dl, d, du, b (complex128[n_rows] each:
 * what fh_assemble_1d_slab_c128 stores, bit for bit) and reduced in the same pass -- 64 B per row written instead of 64
 * written by the assembly + 64 read back by fh_spike_reduce_c128; step 4 is fh_spike_backsub_c128 on the stored bands
 * (open_first = slab does not start the mesh, open_last = does not end it).  162 instead of 226 B per row in all.
 * k_k2_dev != NULL: (k, k**2) from device memory (graph replays). */
FH_API int fh_helmholtz1d_spike_reduce_store_c128(const fh_problem1d *prob_host, double k, double k2,
                                           const double *k_k2_dev, int64_t row_begin, int64_t n_rows, fh_c128 *dl,
                                           fh_c128 *d, fh_c128 *du, fh_c128 *b, void *workspace,
                                           size_t workspace_bytes, fh_c128 *iface8, fh_stream_t stream);
FH_API int fh_spike_solve_interface_c128(const fh_c128 *iface_all, int n_ranks, fh_c128 *x_iface, int32_t *status,
                                  fh_stream_t stream);
FH_API int fh_spike_backsub_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                          int64_t n_rows, int open_first, int open_last, void *workspace,
                          size_t workspace_bytes, const fh_c128 *x_first_last, fh_c128 *u, int32_t *status,
                          fh_stream_t stream);
FH_API int fh_helmholtz1d_spike_backsub_c128(const fh_problem1d *prob_host, double k, double k2, int64_t row_begin,
                                      int64_t n_rows, void *workspace, size_t workspace_bytes,
                                      const fh_c128 *x_first_last, fh_c128 *u, int32_t *status,
                                      fh_stream_t stream);

/* Steps 1 and 4 on generated rows with the wavenumber in device memory (see fh_helmholtz1d_solve_large_fused_kdev_c128):
 * the whole reduce -> all-gather -> interface solve -> back-substitution sequence can be captured once as a CUDA graph
 * and replayed per wavenumber. */
FH_API int fh_helmholtz1d_spike_reduce_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev, int64_t row_begin,
                                          int64_t n_rows, void *workspace, size_t workspace_bytes, fh_c128 *iface8,
                                          fh_stream_t stream);
FH_API int fh_helmholtz1d_spike_backsub_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev,
                                           int64_t row_begin, int64_t n_rows, void *workspace,
                                           size_t workspace_bytes, const fh_c128 *x_first_last, fh_c128 *u,
                                           int32_t *status, fh_stream_t stream);

/* Steps 1-4 as ONE sequence without a library collective: the exchange is carried out by the kernel itself over peer
 * memory (the B200-native form of this path's one exchange step: 128 bytes per rank between kernels a few microseconds
 * long).  peer_bufs_dev: device array [n_ranks] of pointers to every rank's exchange buffer as mapped into this process
 * (symmetric memory over NVLink: cudaIpc / VMM / torch.distributed._symmetric_memory; fh_spike_peer_buffer_bytes() bytes
 * each, zero-initialised once, then owned by these calls).  After its reduction the single-CTA kernel stores its two
 * interface rows into every peer's buffer, publishes its solve counter with a system-scope release, waits (acquire) for
 * every peer's, solves the 2P-row interface system and back-substitutes.  seq_dev: this rank's solve counter (device
 * uint64, zero before the first call, advanced by the kernel -- a captured graph of the call can be replayed).  Every
 * rank must issue the same sequence of calls.  timeout_s (<= 0: 2 s): a peer that never arrives sets status2[0] |= 5
 * instead of hanging the GPU.  x_iface: complex128[2 n_ranks] (as fh_spike_solve_interface_c128); status2: int32[2]
 * (interface solve, back-substitution; may be NULL).  Uniform meshes (FH_E_UNSUPPORTED otherwise). */
FH_API size_t fh_spike_peer_buffer_bytes(void);
FH_API int fh_helmholtz1d_spike_solve_peer_kdev_c128(const fh_problem1d *prob_host, const double *k_k2_dev,
                                              int64_t row_begin, int64_t n_rows, void *workspace,
                                              size_t workspace_bytes, void *const *peer_bufs_dev, int rank,
                                              int n_ranks, unsigned long long *seq_dev, double timeout_s,
                                              fh_c128 *x_iface, fh_c128 *u, int32_t *status2, fh_stream_t stream);

/* ---- residual / parity metrics on device -------------------------------------------------------
 * out[s] = { ||b - A u||_2^2, ||u||_2^2, ||b||_2^2, ||A||_inf } for each system s (device fp64[batch][4]). */
FH_API int fh_tridiag_residual_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                             const fh_c128 *u, int64_t n, int64_t batch, double *out4,
                             fh_stream_t stream);

/* The same four numbers for GENERATED rows (the fused / partitioned solvers never store A): rows [row_begin, row_begin +
 * n_rows) of the system of wavenumber k are re-generated (bit-identical to fh_assemble_1d_c128) and applied to the slab
 * u = device complex128[n_rows] of the solution; u_left / u_right point at the unknowns next to the slab (NULL at a mesh
 * end).  out4 (device fp64[4]) = { sum |b - A u|^2, sum |u|^2, sum |b|^2, max row sum |A| } over the slab: add (max for
 * the fourth) over the ranks, then ||r|| / (||A||_inf ||u|| + ||b||) is the normwise backward error of the whole solve --
 * the gate for meshes no dense reference can run (BASELINE configs 2 and 4; nb:204 is the solve it checks).
 * k_k2_dev != NULL: (k, k**2) are read from device memory (graph replays).  16 B per row of HBM traffic. */
FH_API size_t fh_helmholtz1d_residual_workspace_bytes(void);
FH_API int fh_helmholtz1d_residual_c128(const fh_problem1d *prob_host, double k, double k2, const double *k_k2_dev,
                                 int64_t row_begin, int64_t n_rows, const fh_c128 *u, const fh_c128 *u_left,
                                 const fh_c128 *u_right, double *out4, void *workspace, size_t workspace_bytes,
                                 fh_stream_t stream);

/* Pivoted fallback (LAPACK zgtsv's algorithm: Gaussian elimination with partial pivoting between adjacent rows, which
 * is what np.linalg.solve, nb:204, reaches through zgesv on a tridiagonal A): for the few systems on which the unpivoted
 * solvers break down.  One CTA per system; the serial recurrence runs from both ends of the system at once (two
 * sweeping lanes that meet in a pivoted 2 x 2 system), the rows stream through shared memory one tile ahead.  Any n.
 * u must not alias b.  status[s] = 1 only if the system is singular to working precision, else 0 (may be NULL).
 * workspace: 64 B per row for every system in flight (the query sizes it for min(batch, 256) of them). */
FH_API size_t fh_tridiag_solve_pivoted_workspace_bytes(int64_t n, int64_t batch);
FH_API int fh_tridiag_solve_pivoted_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                  fh_c128 *u, int64_t n, int64_t batch, int32_t *status, void *workspace,
                                  size_t workspace_bytes, fh_stream_t stream);

/* Repair of breakdowns, decided and carried out on the device (no host round trip): every system whose status carries
 * FH_STATUS_NONFINITE (1) is looked at once more -- if the solution that is there (e.g. after
 * fh_tridiag_refine_flagged_c128) is finite and its normwise backward error ||b - A u|| / (||A||_inf ||u|| + ||b||) is
 * <= accept_tol, it is kept and the bit cleared; otherwise the system is re-solved with partial pivoting
 * (fh_tridiag_solve_pivoted_c128's kernel) and status[s] becomes 0, or 1 if it is singular to working precision.
 * On return (stream-ordered) the first five int32 of the workspace hold {flagged, -, kept, re-solved, singular}.
 * max_concurrent: how many re-solves may be in flight (64 B per row each); every flagged system is handled either way. */
FH_API size_t fh_tridiag_repair_flagged_workspace_bytes(int64_t n, int64_t batch, int64_t max_concurrent);
FH_API int fh_tridiag_repair_flagged_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                   fh_c128 *u, int64_t n, int64_t batch, int32_t *status, double accept_tol,
                                   void *workspace, size_t workspace_bytes, fh_stream_t stream);

/* r = b - A u for every system (complex128[batch][n]): the right-hand side of one refinement step. */
FH_API int fh_tridiag_residual_vector_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                    const fh_c128 *u, int64_t n, int64_t batch, fh_c128 *r, fh_stream_t stream);

/* Selective iterative refinement on the device (keeps np.linalg.solve's accuracy, nb:204, on the unpivoted fast path):
 * every system whose status carries FH_STATUS_SMALL_PIVOT (2, set by fh_tridiag_solve_batched_c128) gets one step
 * r = b - A u, A e = r, u += e; the flag is cleared, a correction solve that breaks down sets FH_STATUS_NONFINITE (1).
 * No host synchronisation: the list of flagged systems is built and consumed on the device, and the step itself is ONE
 * pass over the flagged systems (the residual is formed while the rows are read, the correction is added to u by a TMA
 * reduce-add: 96 bytes of traffic per refined row).  The workspace holds the list only (8 + 4 * batch bytes, rounded);
 * max_systems is accepted for compatibility and ignored -- every flagged system is refined by one call.  On return the
 * first two int32 of the workspace hold {flagged, refined}.  n <= fh_tridiag_batched_max_n(). */
FH_API size_t fh_tridiag_refine_flagged_workspace_bytes(int64_t n, int64_t batch, int64_t max_systems);
FH_API int fh_tridiag_refine_flagged_c128(const fh_c128 *dl, const fh_c128 *d, const fh_c128 *du, const fh_c128 *b,
                                   fh_c128 *u, int64_t n, int64_t batch, int32_t *status, void *workspace,
                                   size_t workspace_bytes, fh_stream_t stream);

/* ---- L2 error against the continuum solution -exp(ikx) (replaces nb:259-281 compute_L2_error, quirks kept:
 * complex square without abs, weight h*w).  sum_out[s] = the complex sum BEFORE the final sqrt (nb:281). */
FH_API int fh_l2_error_1d_c128(const double *nodes, const fh_c128 *u, const double *ks, int64_t n_elem, int64_t batch,
                        const double *gauss_pts_host, const double *gauss_wts_host, int n_gauss, fh_c128 *sum_out,
                        fh_stream_t stream);

/* =====================================================================================================
 * 2D: bilinear quads on the annulus (src/solvers/base.py and the four fem2d*.py classes)
 * nodes: device fp64[n_nodes][2]; elements: device int32[n_elements][4] (0-based, counter-clockwise).
 * Gauss points / weights are HOST arrays taken from scipy.special.roots_legendre by the caller, so the
 * kernels use the reference's exact constants.
 * ===================================================================================================== */

/* base.py:65-117 get_element_matrices for every element: Ke = complex128[n_elements][4][4]. */
FH_API int fh_q4_element_matrices_c128(const double *nodes, const int32_t *elements, int64_t n_elements,
                                double k_squared, const double *gauss_pts_host, const double *gauss_wts_host,
                                fh_c128 *Ke, fh_stream_t stream);
/* For each listed element: which of its edges (0,1) (1,2) (2,3) (3,0) have both ends on the circle of `radius`
 * (np.isclose, fem2d.py:47-52) and the edge lengths.  on_edge: uint8[n_ids][4]; length: fp64[n_ids][4]. */
FH_API int fh_q4_edge_geometry(const double *nodes, const int32_t *elements, const int32_t *elem_ids, int64_t n_ids,
                        double radius, uint8_t *on_edge, double *length, fh_stream_t stream);
/* fem2d.py:29-80 / fem2d_dirichlet_sommerfeld.py:26-78: Se = complex128[n_ids][4][4], coef = get_abc_coeff(). */
FH_API int fh_q4_robin_edge_matrices_c128(const uint8_t *on_edge, const double *length, int64_t n_ids, double coef_re,
                                   double coef_im, const double *gauss_pts_host, const double *gauss_wts_host,
                                   int n_gauss, fh_c128 *Se, fh_stream_t stream);
/* fem2d.py:82-137: Ne = complex128[n_ids][4]; g_vals = complex128[n_ids][4][n_gauss] holds the incident normal
 * derivative (fem2d.py:12-27, Bessel functions, evaluated by the caller) at the edge Gauss points. */
FH_API int fh_q4_neumann_edge_vectors_c128(const uint8_t *on_edge, const double *length, int64_t n_ids,
                                    const fh_c128 *g_vals, const double *gauss_pts_host,
                                    const double *gauss_wts_host, int n_gauss, fh_c128 *Ne, fh_stream_t stream);

/* Deterministic scatter (the `+=` loops of *.assemble(), fem2d.py:139-171): keys of every contribution
 * (matrix: 16 per element, key = row * n_nodes + col; vector: 4 per element, key = node), a stable sort that
 * keeps element order inside equal keys, then one ordered sum per unique key. */
FH_API int fh_q4_scatter_keys(const int32_t *elements, const int32_t *elem_ids /* NULL = all */, int64_t count,
                       int64_t n_nodes, int matrix, int64_t *keys, fh_stream_t stream);
FH_API size_t fh_sorted_segments_workspace_bytes(int64_t n_entries);
FH_API int fh_sorted_segments(const int64_t *keys, int64_t n_entries, int32_t *perm, int64_t *ukeys, int64_t *seg,
                       int64_t *n_unique_dev, void *workspace, size_t workspace_bytes, fh_stream_t stream);
FH_API int fh_csr_from_keys(const int64_t *ukeys, const int64_t *n_unique_dev, int64_t n_rows, int64_t n_cols,
                     int64_t *rowptr, int32_t *colidx, fh_stream_t stream);
/* out[s] = (ordered sum of contributions with id < split) + (ordered sum of the rest): K += S, F += N. */
FH_API int fh_segment_sum_c128(const fh_c128 *contrib, const int32_t *perm, const int64_t *seg,
                        const int64_t *n_seg_dev, int64_t n_seg_max, int64_t split, fh_c128 *out,
                        fh_stream_t stream);

/* apply_boundary_conditions (fem2d_dirichlet.py:11-29 etc.): flags[i] |= bit where |r_i - radius| < tol;
 * then row replacement A[row]=0, A[row,row]=1, F[row]=g for rows whose flag intersects `mask`
 * (bit 1 = inner circle, bit 2 = outer circle). */
FH_API int fh_q4_boundary_flags(const double *nodes, int64_t n_nodes, double radius, double tol, int bit, uint8_t *flags,
                         fh_stream_t stream);
FH_API int fh_csr_apply_dirichlet_c128(const int64_t *rowptr, const int32_t *colidx, fh_c128 *vals, fh_c128 *F,
                                const uint8_t *flags, int64_t n, int mask, double g_inner_re, double g_inner_im,
                                double g_outer_re, double g_outer_im, fh_stream_t stream);

/* ---- what surrounds the 2D assembly in the reference's drivers, on the device ------------------------------
 * Boundary detection (helmholtz.py:95-118): node_flag[i] = np.isclose(np.linalg.norm(nodes[i]), radius, atol=atol) (rtol =
 * numpy's default 1e-5), elem_flag[e] = any node of element e flagged; int32 flags, fh_compact_flags_i32 turns either
 * into the ascending index list np.where(...)[0] (list: int32[n], count: device int32[1]). */
FH_API int fh_q4_boundary_lists(const double *nodes, int64_t n_nodes, const int32_t *elements, int64_t n_elements,
                         double radius, double atol, int32_t *node_flag, int32_t *elem_flag, fh_stream_t stream);
FH_API int fh_compact_flags_i32(const int32_t *flags, int64_t n, int32_t *list, int32_t *count, fh_stream_t stream);
/* get_normal_derivative (fem2d.py:12-27) at the Gauss points of the flagged edges (fem2d.py:113-127):
 * g_vals = complex128[n_ids][4][n_gauss] = sum_{m=-n_fourier}^{n_fourier} i^m k J_m'(k r) e^{i m theta}, the input of
 * fh_q4_neumann_edge_vectors_c128.  J_m from CUDA's jn(): ~1e-11 absolute, not scipy's last bit (the host path is). */
FH_API int fh_q4_neumann_load_c128(const double *nodes, const int32_t *elements, const int32_t *elem_ids, int64_t n_ids,
                            const uint8_t *on_edge, double k, int n_fourier, const double *gauss_pts_host, int n_gauss,
                            fh_c128 *g_vals, fh_stream_t stream);
/* compute_L2_error (base.py:119-177).  fh_q4_gauss_points: xy = fp64[n_elements][n_gauss][n_gauss][2], the (x, y) of
 * base.py:141-148 (where the caller evaluates the analytic solution -- Bessel closed forms, scipy, as the reference).
 * fh_q4_l2_error_c128: sum_out[0] = sum over elements and Gauss points of |u_h - u_exact|^2 det J w_i w_j with
 * u = complex128[n_nodes], u_exact = complex128[n_elements][n_gauss][n_gauss]; the caller takes the square root. */
FH_API int fh_q4_gauss_points(const double *nodes, const int32_t *elements, int64_t n_elements,
                       const double *gauss_pts_host, int n_gauss, double *xy, fh_stream_t stream);
FH_API size_t fh_q4_l2_error_workspace_bytes(void);
FH_API int fh_q4_l2_error_c128(const double *nodes, const int32_t *elements, int64_t n_elements, const fh_c128 *u,
                        const fh_c128 *u_exact, const double *gauss_pts_host, const double *gauss_wts_host, int n_gauss,
                        double *sum_out, void *workspace, size_t workspace_bytes, fh_stream_t stream);

/* *.solve(): np.linalg.solve(K, F) (fem2d.py:176) -> dense LU with partial pivoting on the device.
 * A (n x n, row-major) and b are overwritten; piv: int32[n]; status: int32[1], non-zero = singular. */
FH_API int fh_csr_to_dense_c128(const int64_t *rowptr, const int32_t *colidx, const fh_c128 *vals, int64_t n,
                         fh_c128 *A, fh_stream_t stream);
FH_API int fh_dense_lu_solve_c128(fh_c128 *A, fh_c128 *b, fh_c128 *x, int64_t n, int32_t *piv, int32_t *status,
                           fh_stream_t stream);

/* *.solve() without the dense matrix: the CSR system is scattered into LAPACK band storage under an
 * optional node permutation (perm[i] = new index of node i, NULL = identity), factorised with partial pivoting (the
 * pivot rule and elimination order of zgbtf2 / zgesv, i.e. np.linalg.solve's accuracy) and solved; x comes back in the
 * original numbering.  kl / ku = lower / upper bandwidth under the permutation, from fh_csr_bandwidth (klku: device
 * int32[2]).  status: device int32[1], non-zero = singular to working precision.  When the update window of a column
 * step holds 4,096 entries or more (kl (kl + ku)), the elimination is shared by a thread-block cluster of 16 (8) CTAs --
 * same pivots, same operations per entry; the workspace (fh_band_lu_workspace_bytes) holds the band, the permuted
 * right-hand side and the reciprocal pivots. */
FH_API int fh_csr_bandwidth(const int64_t *rowptr, const int32_t *colidx, const int32_t *perm, int64_t n,
                     int32_t *klku, fh_stream_t stream);
FH_API size_t fh_band_lu_workspace_bytes(int64_t n, int64_t kl, int64_t ku);
FH_API int fh_csr_band_lu_solve_c128(const int64_t *rowptr, const int32_t *colidx, const fh_c128 *vals,
                              const fh_c128 *b, const int32_t *perm, int64_t n, int64_t kl, int64_t ku, fh_c128 *x,
                              void *workspace, size_t workspace_bytes, int32_t *status, fh_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FEM_HELMHOLTZ_H */
