"""1D P1 Helmholtz FEM on the GPU behind the reference notebook's own function names.

Mirrors code cell 2 of the reference's ``notebooks/1D_soln.ipynb`` (``nb:N`` = raw JSON line):

    create_mesh            nb:33-47      shape_functions      nb:55-67
    element_matrices       nb:70-96      boundary_matrices    nb:99-131
    assembly               nb:134-187    solve_helmholtz_fem  nb:190-205
    analytical_solution    nb:208-219    compute_L2_error     nb:259-281

Same names, argument order, defaults and return shapes.  What differs is where the work
happens: assembly writes tridiagonal bands (the reference's dense ``A`` is tridiagonal) with a
CUDA kernel, and the solve is a CUDA batched tridiagonal solver; both are reached through the
C ABI of ``include/fem_helmholtz.h``.  There is no CPU fallback.

Additions (keyword-only, defaults keep the reference behaviour): Dirichlet row replacement
(``solvers/fem2d_dirichlet.py:20-27`` convention) via ``bc_left`` / ``bc_right``, a batched
wavenumber sweep ``solve_helmholtz_sweep`` and the fused (bands never stored) variant.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BC_CODES, FH_H_NODAL, FH_H_SCALAR, Problem1d, check, ptr, stream_ptr

# nb:52: gauss_point, gauss_weight = np.polynomial.legendre.leggauss(1)
gauss_point = np.array([0.0])
gauss_weight = np.array([2.0])


# ------------------------------------------------------------------------------------------------
# mesh (L1 in SURVEY.md: problem data, produced on the host exactly as the reference does)
# ------------------------------------------------------------------------------------------------
def create_mesh(N, L=1.0):
    """nb:33-47.  ``h`` is the scalar ``L / N``; connectivity is the canonical ``[i, i+1]``
    (the GPU kernels treat it as implicit and never read it)."""
    h = L / N
    nodes = np.linspace(0, L, N + 1)
    idx = np.arange(N)
    elements = np.stack([idx, idx + 1], axis=1)
    return h, nodes, elements


# ------------------------------------------------------------------------------------------------
# element routines (device, one thread: they exist for API parity and bit-parity tests)
# ------------------------------------------------------------------------------------------------
def shape_functions(xi):
    """nb:55-67 -> (basis_fns[2], dbasis_dxi[2]) as numpy arrays."""
    lib = _lib.require_cuda()
    out = torch.empty(4, dtype=torch.float64, device=_lib.device())
    check(lib.fh_shape_functions_1d(float(xi), ptr(out[:2]), ptr(out[2:]), stream_ptr()),
          "fh_shape_functions_1d")
    res = out.cpu().numpy()
    return res[:2].copy(), res[2:].copy()


def element_matrices(h, xi=gauss_point[0], w=gauss_weight[0]):
    """nb:70-96 -> (Ke[2,2], Me[2,2]) as numpy arrays, bit-identical to the reference."""
    lib = _lib.require_cuda()
    out = torch.empty(8, dtype=torch.float64, device=_lib.device())
    check(lib.fh_element_matrices_1d(float(h), float(xi), float(w), ptr(out[:4]), ptr(out[4:]),
                                     stream_ptr()), "fh_element_matrices_1d")
    res = out.cpu().numpy()
    return res[:4].reshape(2, 2).copy(), res[4:].reshape(2, 2).copy()


def boundary_matrices(h, k, theta=0.0):
    """nb:99-131 -> (Ne[2], Se[2,2]) complex numpy arrays (``h`` is unused, as in the reference)."""
    lib = _lib.require_cuda()
    out = torch.empty(6, dtype=torch.complex128, device=_lib.device())
    check(lib.fh_boundary_matrices_1d(float(k), float(np.cos(theta)), ptr(out[:2]), ptr(out[2:]),
                                      stream_ptr()), "fh_boundary_matrices_1d")
    res = out.cpu().numpy()
    return res[:2].copy(), res[2:].reshape(2, 2).copy()


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def squares_like_the_reference(ks):
    """``k**2`` for every entry of a float64 array, evaluated the way the reference evaluates it (nb:185: ``k`` is a Python
    or numpy SCALAR there, one wavenumber per call, so ``k**2`` is libm's ``pow(k, 2.0)``).  That is not always the
    correctly rounded square: ``ks**2`` / ``ks*ks`` of an array differ from it by one ulp for about one wavenumber in a
    thousand (64 of the 65,536 of BASELINE config 3), and with them the assembled diagonals -- found by the randomised
    parity test.  ``np.float_power`` goes through libm's pow element by element (0.9 ms for 65,536 values); a sample is
    checked against CPython's own ``**`` and the loop over scalars is the fallback."""
    ks = np.ascontiguousarray(ks, dtype=np.float64).ravel()
    for held, squares in _SQUARES_SEEN:  # a sweep driver hands over the same wavenumbers again and again (30 us to compare)
        if held.shape == ks.shape and np.array_equal(held, ks):
            return squares.copy()
    out = np.float_power(ks, 2.0)
    probe = np.unique(np.linspace(0, max(len(ks) - 1, 0), num=min(len(ks), 64)).astype(np.int64)) if len(ks) else []
    if any(float(ks[i]) ** 2 != out[i] for i in probe):
        out = np.array([float(v) ** 2 for v in ks], dtype=np.float64)
    if len(ks) >= 1024:
        _SQUARES_SEEN.insert(0, (ks.copy(), out.copy()))
        del _SQUARES_SEEN[2:]
    return out


_SQUARES_SEEN = []  # the last two large wavenumber arrays and their squares


def _host_k_and_k2(k):
    """Return (ks fp64[n_k], k2s fp64[n_k], scalar?) with k**2 evaluated on the host the way nb:185 does: entry by
    entry as a scalar power (:func:`squares_like_the_reference`), whatever container the wavenumbers come in."""
    if isinstance(k, torch.Tensor):
        k = k.detach().cpu().numpy()
    if np.isscalar(k) or (isinstance(k, np.ndarray) and k.ndim == 0):
        return np.array([float(k)]), np.array([float(k**2)]), True
    if isinstance(k, np.ndarray):
        ks = np.ascontiguousarray(k, dtype=np.float64).ravel()
        return ks, squares_like_the_reference(ks), False
    ks = list(k)
    return (np.array([float(v) for v in ks], dtype=np.float64),
            np.array([float(v**2) for v in ks], dtype=np.float64), False)


def _problem(N, h, theta, bc_left, bc_right, g_left, g_right, nodes_dev):
    if bc_left not in ("neumann", "dirichlet"):
        raise ValueError(f"bc_left must be 'neumann' or 'dirichlet', got {bc_left!r}")
    if bc_right not in ("sommerfeld", "dirichlet"):
        raise ValueError(f"bc_right must be 'sommerfeld' or 'dirichlet', got {bc_right!r}")
    gl, gr = complex(g_left), complex(g_right)
    return Problem1d(
        n_elem=int(N), h=float(h), x=(nodes_dev.data_ptr() if nodes_dev is not None else None),
        h_mode=FH_H_NODAL if nodes_dev is not None else FH_H_SCALAR,
        bc_left=BC_CODES[bc_left], bc_right=BC_CODES[bc_right],
        g_left_re=gl.real, g_left_im=gl.imag, g_right_re=gr.real, g_right_im=gr.imag,
        cos_theta=float(np.cos(theta)), gauss_xi=float(gauss_point[0]), gauss_w=float(gauss_weight[0]))


def _check_elements(N, elements):
    if elements is None:
        return
    shape = tuple(elements.shape)
    if shape != (N, 2):
        raise ValueError(f"elements must have shape ({N}, 2), got {shape}")
    first, last = np.asarray(elements[0]).tolist(), np.asarray(elements[-1]).tolist()
    if first != [0, 1] or last != [N - 1, N]:
        raise ValueError("only the reference's canonical connectivity [[i, i+1]] is supported")


STATUS_NONFINITE = 1    # breakdown of the unpivoted solve: NaN / Inf (zero pivot, singular system) or a defect > 1e-8
STATUS_SMALL_PIVOT = 2  # the interface-row probe found a defect > 5e-14; cleared again once the system has been refined


def _raise_if_singular(status):
    if status is not None and bool(torch.any((status & STATUS_NONFINITE) != 0).item()):
        bad = torch.nonzero(status & STATUS_NONFINITE).flatten()[:8].tolist()
        raise np.linalg.LinAlgError(f"Singular matrix (zero pivot / non-finite solution) in systems {bad}")


# ------------------------------------------------------------------------------------------------
# the assembled operator
# ------------------------------------------------------------------------------------------------
class TridiagonalSystem:
    """Device-resident ``A`` of ``assembly``: bands ``dl, d, du`` as complex128[batch, n] CUDA tensors
    (``dl[:, 0]`` and ``du[:, n-1]`` are zero).  ``np.asarray(A)`` / ``A.to_dense()`` materialise the
    reference's dense matrix, so ``np.linalg.solve(A, b)`` still works for small N."""

    def __init__(self, dl, d, du, rhs=None, scalar=True):
        self.dl, self.d, self.du, self.rhs = dl, d, du, rhs
        self.batch, self.n = d.shape
        self._scalar = scalar

    @property
    def shape(self):
        return (self.n, self.n) if self._scalar else (self.batch, self.n, self.n)

    dtype = np.dtype(np.complex128)

    def bands(self, i=0):
        """(dl, d, du) of system i as numpy arrays of length n (LAPACK gtsv style + padding)."""
        return tuple(t[i].cpu().numpy() for t in (self.dl, self.d, self.du))

    def to_dense(self, i=0):
        dl, d, du = self.bands(i)
        A = np.zeros((self.n, self.n), dtype=complex)
        idx = np.arange(self.n)
        A[idx, idx] = d
        A[idx[1:], idx[:-1]] = dl[1:]
        A[idx[:-1], idx[1:]] = du[:-1]
        return A

    def __array__(self, dtype=None, copy=None):
        if self.batch != 1:
            raise ValueError("np.asarray() of a batched system is ambiguous; use to_dense(i)")
        A = self.to_dense(0)
        return A.astype(dtype) if dtype is not None else A

    def solve(self, b=None, check_singular=True):
        """Solve every system of the batch on the GPU.  ``b``: None (use the assembled load),
        numpy or torch, shape [n] or [batch, n].  Returns a CUDA tensor complex128[batch, n]."""
        rhs = self.rhs if b is None else _lib.as_device(b, torch.complex128).reshape(self.batch, self.n)
        u, status = solve_tridiagonal(self.dl, self.d, self.du, rhs)
        if check_singular:
            _raise_if_singular(status)
        return u

    def residual(self, u, b=None):
        rhs = self.rhs if b is None else _lib.as_device(b, torch.complex128).reshape(self.batch, self.n)
        return tridiagonal_residual(self.dl, self.d, self.du, rhs, u)


# ------------------------------------------------------------------------------------------------
# assembly + solve
# ------------------------------------------------------------------------------------------------
def assemble_bands(N, k, h, theta=0.0, *, bc_left="neumann", bc_right="sommerfeld", g_left=1.0,
                   g_right=0.0, nodes=None, out=None):
    """Device-side assembly (nb:134-187) for one or many wavenumbers.

    Returns ``(dl, d, du, b)`` CUDA tensors complex128[n_k, N+1].  ``nodes`` (fp64[N+1]) switches to
    per-element sizes ``x[e+1]-x[e]``; the default reproduces the reference's scalar ``h``.
    ``out``: optional preallocated complex128[4, n_k, N+1] CUDA tensor."""
    lib = _lib.require_cuda()
    ks, k2s = k if isinstance(k, tuple) else _host_k_and_k2(k)[:2]  # tuple = (ks, k2s) already split
    n_k, n = len(ks), int(N) + 1
    dev = _lib.device()
    ks_d, k2s_d = _lib.as_device(ks, torch.float64), _lib.as_device(k2s, torch.float64)
    nodes_d = _lib.as_device(nodes, torch.float64) if nodes is not None else None
    if nodes_d is not None and nodes_d.numel() != n:
        raise ValueError(f"nodes must have {n} entries")
    if out is None:
        out = torch.empty((4, n_k, n), dtype=torch.complex128, device=dev)
    assert out.shape == (4, n_k, n) and out.dtype == torch.complex128 and out.is_contiguous()
    prob = _problem(N, h, theta, bc_left, bc_right, g_left, g_right, nodes_d)
    # non-uniform mesh, many wavenumbers: the row geometry is tabulated once in caller-owned scratch (48 B per row)
    ws_bytes = lib.fh_assemble_1d_workspace_bytes(int(N), prob.h_mode, n_k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    check(lib.fh_assemble_1d_ws_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), n_k, ptr(out[0]), ptr(out[1]),
                                     ptr(out[2]), ptr(out[3]), ptr(ws), ws_bytes, stream_ptr()), "fh_assemble_1d_ws_c128")
    return out[0], out[1], out[2], out[3]


def assembly(N: int, k, h: float, elements=None, theta=0.0, *, bc_left="neumann", bc_right="sommerfeld",
             g_left=1.0, g_right=0.0, nodes=None):
    """nb:134-187 -> ``(A, b)``.

    ``A`` is a :class:`TridiagonalSystem` (device bands; ``np.asarray(A)`` is the reference's dense
    matrix).  For a scalar ``k`` (the reference's call) ``b`` is a numpy complex128[N+1] array; for a
    sequence of wavenumbers ``b`` is a CUDA tensor complex128[n_k, N+1]."""
    _check_elements(int(N), elements)
    _, _, scalar = _host_k_and_k2(k)
    dl, d, du, b = assemble_bands(N, k, h, theta, bc_left=bc_left, bc_right=bc_right, g_left=g_left,
                                  g_right=g_right, nodes=nodes)
    A = TridiagonalSystem(dl, d, du, rhs=b, scalar=scalar)
    return (A, b[0].cpu().numpy()) if scalar else (A, b)


_refine_ws = {}   # (n, batch, device) -> (capacity in systems, buffer); a handful of shapes are kept
_WS_KEEP = 4


def _remember(cache, key, value):
    cache.pop(key, None)
    while len(cache) >= _WS_KEEP:
        cache.pop(next(iter(cache)))
    cache[key] = value
    return value


def _refine_workspace(n, batch, device, min_systems=0):
    """Workspace of fh_tridiag_refine_flagged_c128, kept between calls (allocating it per step would cost more than the
    refinement).  Sized for batch/16 systems per call to begin with (a fine sweep flags 2-3 % of its wavenumbers) and
    grown when a batch turns out to flag more (``min_systems``)."""
    lib = _lib.load()
    key = (n, batch, device)
    ent = _refine_ws.get(key)
    cap = min(batch, max(256, batch // 16, int(min_systems)))
    if ent is None or ent[0] < cap:
        nbytes = lib.fh_tridiag_refine_flagged_workspace_bytes(n, batch, cap)
        ent = _remember(_refine_ws, key, (cap, torch.empty(nbytes, dtype=torch.uint8, device=device)))
    return ent[1]


def refine_flagged(dl, d, du, b, u, status, sync=True, ws=None):
    """One step of iterative refinement for exactly the systems whose elimination met a small pivot
    (``status & STATUS_SMALL_PIVOT``): r = b - A u, solve A e = r, u += e -- all on the device, in one pass over the
    flagged systems, which are read in place through an index list that never visits the host
    (``fh_tridiag_refine_flagged_c128``).  The solver is
    unpivoted, so a pivot that cancels to 1e-5 of its scale costs five digits of backward error (seen for ~1 in 10^4
    wavenumbers of a fine sweep); one refinement step restores it.  Returns the number of refined systems; reading it
    is the only host synchronisation (``sync=False`` skips it and returns None; ``ws``: a caller-owned workspace --
    since round 2 it holds the index list only, so every flagged system is refined by one call)."""
    lib = _lib.require_cuda()
    batch, n = d.shape
    if n > lib.fh_tridiag_batched_max_n():
        return _refine_large(dl, d, du, b, u, status)
    own = ws is None
    if own:
        ws = _refine_workspace(n, batch, d.device)
    total = 0
    while True:
        check(lib.fh_tridiag_refine_flagged_c128(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(u), n, batch, ptr(status),
                                                 ptr(ws), ws.numel(), stream_ptr()), "fh_tridiag_refine_flagged_c128")
        if not sync:
            return None
        flagged, refined = (int(v) for v in ws[:8].view(torch.int32).tolist())
        total += refined
        if flagged <= refined:
            return total
        # more flagged systems than the workspace holds: finish them now, and size the workspace for the next batch
        # of this shape (so that a later sync=False call covers it in one go)
        if own:
            ws = _refine_workspace(n, batch, d.device, min_systems=flagged + flagged // 4)


def _refine_large(dl, d, du, b, u, status):
    """:func:`refine_flagged` for systems too long for the batched kernel (n > 4112: the recursive-partition solver,
    whose back-substitution carries the same interface-row probe).  These come one or a few at a time, so the loop is
    driven from the host: r = b - A u, A e = r by the same solver, u += e.  The small-pivot bit is cleared; a breakdown
    of the correction solve (non-finite, or again a severe defect) sets the breakdown bit, which sends the system to
    :func:`repair_flagged` (verify, else pivoted re-solve)."""
    lib = _lib.require_cuda()
    batch, n = d.shape
    st_host = status.cpu()
    idx = [i for i in range(batch) if int(st_host[i]) & STATUS_SMALL_PIVOT]
    if not idx:
        return 0
    ws_bytes = lib.fh_tridiag_solve_large_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=d.device)
    r, e = torch.empty(n, dtype=torch.complex128, device=d.device), torch.empty(n, dtype=torch.complex128, device=d.device)
    st_e = torch.zeros(1, dtype=torch.int32, device=d.device)
    for i in idx:
        check(lib.fh_tridiag_residual_vector_c128(ptr(dl[i]), ptr(d[i]), ptr(du[i]), ptr(b[i]), ptr(u[i]), n, 1, ptr(r),
                                                  stream_ptr()), "fh_tridiag_residual_vector_c128")
        check(lib.fh_tridiag_solve_large_c128(ptr(dl[i]), ptr(d[i]), ptr(du[i]), ptr(r), ptr(e), n, ptr(st_e), ptr(ws),
                                              ws_bytes, stream_ptr()), "fh_tridiag_solve_large_c128")
        u[i] += e
        status[i:i + 1] = (status[i:i + 1] & ~STATUS_SMALL_PIVOT) | (st_e & STATUS_NONFINITE)
    return len(idx)


# What counts as repaired: three digits inside the 1e-10 contract (nb:204 / BASELINE north_star).  The unpivoted fast
# path sits at 2e-16 (median) and LAPACK's zgtsv at 2e-17, but a refined solution at a resonance of the partition can
# land at a few 1e-15 (k = 806.8412 in the 8-rank sweep: 2.4e-15) -- re-solving that with the serial pivoted kernel
# buys nothing the contract asks for.
ACCEPT_BACKWARD_ERROR = 1e-13

_repair_ws = {}


def _repair_workspace(n, batch, device, concurrent=16):
    """Workspace of fh_tridiag_repair_flagged_c128, kept between calls (64 B per row for each of ``concurrent`` re-solves
    in flight + 4 B per system of the batch)."""
    lib = _lib.load()
    key = (n, batch, device)
    ent = _repair_ws.get(key)
    if ent is None:
        nbytes = lib.fh_tridiag_repair_flagged_workspace_bytes(n, batch, concurrent)
        ent = _remember(_repair_ws, key, torch.empty(nbytes, dtype=torch.uint8, device=device))
    return ent


def repair_flagged(dl, d, du, b, u, status, sync=True, accept_tol=ACCEPT_BACKWARD_ERROR, ws=None):
    """Deal with the systems on which the unpivoted kernels reported a breakdown (``status & STATUS_NONFINITE``): a
    non-finite value -- a vanishing leading minor is harmless to LAPACK's zgesv, which the reference uses (nb:204), so
    it must be harmless here -- or a probe defect beyond what one refinement step is trusted to repair.  Entirely on the
    device (``fh_tridiag_repair_flagged_c128``: compact the flags, verify, re-solve), stream-ordered, no host round
    trip: a flagged system whose present solution (run :func:`refine_flagged` first) is finite with a normwise backward
    error <= ``accept_tol`` is kept; any other is re-solved with partial pivoting by a one-CTA-per-system kernel
    (~0.2 ms for 4,097 rows).  Afterwards the flag stays set only for systems that are singular to working precision.

    ``sync=True`` reads the counters back and returns ``{"flagged", "kept", "resolved", "singular"}``; ``sync=False``
    returns None (the counters stay in the first 32 bytes of the workspace: :func:`repair_counts`)."""
    lib = _lib.require_cuda()
    batch, n = d.shape
    if ws is None:
        ws = _repair_workspace(n, batch, d.device)
    check(lib.fh_tridiag_repair_flagged_c128(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(u), n, batch, ptr(status),
                                             float(accept_tol), ptr(ws), ws.numel(), stream_ptr()),
          "fh_tridiag_repair_flagged_c128")
    return repair_counts(ws) if sync else None


def repair_counts(ws):
    """Counters left in a repair workspace by the last :func:`repair_flagged` call (synchronises)."""
    h = ws[:32].view(torch.int32).tolist()
    return {"flagged": h[0], "kept": h[2], "resolved": h[3], "singular": h[4]}


def resolve_nonfinite(dl, d, du, b, u, status):
    """:func:`repair_flagged`, returning the number of systems that went through the pivoted re-solve."""
    return repair_flagged(dl, d, du, b, u, status)["resolved"]


def solve_tridiagonal(dl, d, du, b, out=None, refine=True):
    """Batched tridiagonal solve on the GPU (replaces nb:204).  All arguments complex128[batch, n]
    CUDA tensors (numpy accepted and uploaded).  Returns ``(u, status)``; ``status[i] & STATUS_NONFINITE`` flags a
    breakdown (the caller decides whether to raise).  ``refine=True`` runs :func:`refine_flagged` afterwards."""
    lib = _lib.require_cuda()
    dl, d, du, b = (_lib.as_device(t, torch.complex128) for t in (dl, d, du, b))
    if d.dim() == 1:
        dl, d, du, b = (t.reshape(1, -1) for t in (dl, d, du, b))
    batch, n = d.shape
    for t in (dl, du, b):
        if tuple(t.shape) != (batch, n):
            raise ValueError("dl, d, du, b must share the shape [batch, n]")
    u = out if out is not None else torch.empty_like(d)
    status = torch.empty(batch, dtype=torch.int32, device=d.device)
    ws_bytes = lib.fh_tridiag_solve_batched_workspace_bytes(n, batch)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=d.device)
    check(lib.fh_tridiag_solve_batched_c128(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(u), n, batch, ptr(status),
                                            ptr(ws), ws_bytes, stream_ptr()), "fh_tridiag_solve_batched_c128")
    if refine:
        # both steps are stream-ordered device work: the refinement first (it repairs most severe defects too), then
        # the verify-or-pivot pass over whatever still carries the breakdown bit
        refine_flagged(dl, d, du, b, u, status)
        repair_flagged(dl, d, du, b, u, status, sync=False)
    return u, status


def tridiagonal_residual(dl, d, du, b, u):
    """Device-side parity metrics per system: dict of numpy arrays ``backward`` (normwise backward
    error ||b-Au|| / (||A||_inf ||u|| + ||b||)), ``relres`` (||b-Au||/||b||), ``r2, u2, b2, anorm``."""
    lib = _lib.require_cuda()
    dl, d, du, b, u = (_lib.as_device(t, torch.complex128) for t in (dl, d, du, b, u))
    if d.dim() == 1:
        dl, d, du, b, u = (t.reshape(1, -1) for t in (dl, d, du, b, u))
    batch, n = d.shape
    out = torch.empty((batch, 4), dtype=torch.float64, device=d.device)
    check(lib.fh_tridiag_residual_c128(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(u), n, batch, ptr(out),
                                       stream_ptr()), "fh_tridiag_residual_c128")
    o = out.cpu().numpy()
    r, un, bn, an = np.sqrt(o[:, 0]), np.sqrt(o[:, 1]), np.sqrt(o[:, 2]), o[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        return {"backward": r / (an * un + bn), "relres": r / bn, "r2": o[:, 0], "u2": o[:, 1],
                "b2": o[:, 2], "anorm": an}


PIPELINES = ("staged", "one_pass", "fused")


def _pick_pipeline(pipeline, fused, N, nodal_h):
    """``pipeline=None`` -> "fused" if the legacy ``fused=True`` switch is set, else the default: the one-pass kernel
    (A written once, solved from registers: same bits as the staged pair, 80 instead of 144 B/DOF of traffic; meshes
    beyond fh_tridiag_batched_max_n() nodes: A written by the first reduction level of the recursive-partition solver,
    162 instead of 226 B/DOF) wherever it applies -- uniform meshes of 3 nodes or more --, the staged pair elsewhere."""
    if pipeline is None:
        pipeline = "fused" if fused else "auto"
    if pipeline == "auto":
        pipeline = "one_pass" if (not nodal_h) and int(N) >= 2 else "staged"
    if pipeline not in PIPELINES:
        raise ValueError(f"pipeline must be one of {PIPELINES} (or None / 'auto'), got {pipeline!r}")
    return pipeline


def solve_helmholtz_sweep(N, ks, L=1.0, theta=0.0, *, bc_left="neumann", bc_right="sommerfeld",
                          g_left=1.0, g_right=0.0, fused=False, pipeline=None, nodal_h=False, out=None,
                          check_singular=True, return_status=False):
    """Assemble and solve the 1D problem for every wavenumber in ``ks`` (the shape of the reference's
    sweep loop nb:300-317 / nb:469-471, batched).  Returns ``(nodes, U)`` with ``U`` a CUDA tensor
    complex128[n_k, N+1].

    ``pipeline``: "staged" -- ``assembly`` writes the bands, the batched solver reads them back (the reference's
    structure, two kernels); "one_pass" -- one kernel writes the same bands and solves from its registers (bit-identical
    results; the default where it applies); "fused" (or ``fused=True``) -- the rows are generated in registers and only
    ``U`` is written.  In every variant flagged systems are refined / repaired to ``np.linalg.solve`` quality."""
    lib = _lib.require_cuda()
    h, nodes = L / N, np.linspace(0, L, N + 1)  # create_mesh without the (unused, 16 B/element) connectivity
    ks_h, k2s_h, _ = _host_k_and_k2(ks)
    n_k, n = len(ks_h), int(N) + 1
    dev = _lib.device()
    pipeline = _pick_pipeline(pipeline, fused, N, nodal_h)
    fused = pipeline == "fused"
    nodes_d = _lib.as_device(nodes, torch.float64) if nodal_h else None
    if pipeline == "one_pass" and (nodal_h or N < 2):
        pipeline = "staged"
    if pipeline == "one_pass":
        _, U, status = assemble_and_solve(N, (ks_h, k2s_h), L, theta, bc_left=bc_left, bc_right=bc_right, g_left=g_left,
                                          g_right=g_right, out=out)
    elif not fused:
        dl, d, du, b = assemble_bands(N, (ks_h, k2s_h), h, theta, bc_left=bc_left, bc_right=bc_right, g_left=g_left, g_right=g_right,
                                      nodes=nodes if nodal_h else None)
        U, status = solve_tridiagonal(dl, d, du, b, out=out)
    else:
        ks_d, k2s_d = _lib.as_device(ks_h, torch.float64), _lib.as_device(k2s_h, torch.float64)
        U = out if out is not None else torch.empty((n_k, n), dtype=torch.complex128, device=dev)
        status = torch.zeros(n_k, dtype=torch.int32, device=dev)
        prob = _problem(N, h, theta, bc_left, bc_right, g_left, g_right, nodes_d)
        if n <= lib.fh_tridiag_batched_max_n():
            check(lib.fh_helmholtz1d_sweep_fused_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), n_k, ptr(U),
                                                      ptr(status), None, 0, stream_ptr()),
                  "fh_helmholtz1d_sweep_fused_c128")
        else:  # long meshes: recursive partition with generated level-0 rows, one wavenumber at a time
            ws_bytes = lib.fh_tridiag_solve_large_workspace_bytes(n)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            for i in range(n_k):
                check(lib.fh_helmholtz1d_solve_large_fused_c128(
                    C.byref(prob), float(ks_h[i]), float(k2s_h[i]), ptr(U[i]), ptr(status[i:i + 1]), ptr(ws),
                    ws_bytes, stream_ptr()), "fh_helmholtz1d_solve_large_fused_c128")
    if fused:
        idx = torch.nonzero(status != 0).flatten()
        if idx.numel() > 0:
            # the fused kernels store no bands: assemble them for the few flagged wavenumbers (small-pivot probe or
            # breakdown; short and long meshes alike) and solve those again through the staged path (refinement,
            # verification, pivoted re-solve)
            sel = idx.cpu().numpy()
            bands = assemble_bands(N, (ks_h[sel], k2s_h[sel]), h, theta, bc_left=bc_left, bc_right=bc_right,
                                   g_left=g_left, g_right=g_right, nodes=nodes if nodal_h else None)
            u_sub, st_sub = solve_tridiagonal(*bands)
            U.index_copy_(0, idx, u_sub)
            status.index_copy_(0, idx, st_sub)
    if check_singular:
        _raise_if_singular(status)
    return (nodes, U, status) if return_status else (nodes, U)


def assemble_and_solve(N, ks, L=1.0, theta=0.0, *, bc_left="neumann", bc_right="sommerfeld", g_left=1.0, g_right=0.0,
                       out_bands=None, out=None, refine=True):
    """``A, b = assembly(...)`` followed by ``u = np.linalg.solve(A, b)`` (nb:203-204) for every wavenumber in ``ks``, in
    ONE pass: the kernel generates the rows, stores them as the bands ``assemble_bands`` would have stored (bit for
    bit) and solves from its registers, so ``A`` is written once and never read back.  Returns
    ``((dl, d, du, b), U, status)`` -- CUDA tensors complex128[n_k, N+1].  Meshes beyond the batched solver's size are
    taken one system at a time through ``fh_helmholtz1d_assemble_solve_large_c128`` (the solver's first reduction level
    writes the bands); meshes of fewer than 2 elements take the two-kernel path."""
    lib = _lib.require_cuda()
    ks_h, k2s_h = ks if isinstance(ks, tuple) else _host_k_and_k2(ks)[:2]  # tuple = (ks, k2s) already split
    n_k, n = len(ks_h), int(N) + 1
    dev = _lib.device()
    if N < 2:
        bands = assemble_bands(N, (ks_h, k2s_h), L / N, theta, bc_left=bc_left, bc_right=bc_right, g_left=g_left,
                               g_right=g_right, out=out_bands)
        U, status = solve_tridiagonal(*bands, out=out, refine=refine)
        return bands, U, status
    B = out_bands if out_bands is not None else torch.empty((4, n_k, n), dtype=torch.complex128, device=dev)
    assert B.shape == (4, n_k, n) and B.dtype == torch.complex128 and B.is_contiguous()
    U = out if out is not None else torch.empty((n_k, n), dtype=torch.complex128, device=dev)
    status = torch.empty(n_k, dtype=torch.int32, device=dev)
    prob = _problem(N, L / N, theta, bc_left, bc_right, g_left, g_right, None)
    if n > lib.fh_tridiag_batched_max_n():
        # long meshes, one system at a time: the first reduction level of the recursive-partition solver writes the
        # bands (fh_helmholtz1d_assemble_solve_large_c128: same bits as assembly + solve, 162 instead of 226 B per row)
        nbytes = lib.fh_tridiag_solve_large_workspace_bytes(n)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        for i in range(n_k):
            check(lib.fh_helmholtz1d_assemble_solve_large_c128(
                C.byref(prob), float(ks_h[i]), float(k2s_h[i]), None, ptr(B[0][i]), ptr(B[1][i]), ptr(B[2][i]),
                ptr(B[3][i]), ptr(U[i]), ptr(status[i:i + 1]), ptr(ws), nbytes, stream_ptr()),
                "fh_helmholtz1d_assemble_solve_large_c128")
        bands = (B[0], B[1], B[2], B[3])
        if refine:
            refine_flagged(*bands, U, status)
            repair_flagged(*bands, U, status, sync=False)
        return bands, U, status
    ks_d, k2s_d = _lib.as_device(ks_h, torch.float64), _lib.as_device(k2s_h, torch.float64)
    check(lib.fh_helmholtz1d_sweep_assemble_solve_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), n_k, ptr(B[0]), ptr(B[1]),
                                                       ptr(B[2]), ptr(B[3]), ptr(U), ptr(status), None, 0, stream_ptr()),
          "fh_helmholtz1d_sweep_assemble_solve_c128")
    bands = (B[0], B[1], B[2], B[3])
    if refine:
        refine_flagged(*bands, U, status)
        repair_flagged(*bands, U, status, sync=False)
    return bands, U, status


class _WavenumberFeed:
    """Hands (k, k**2) to device memory without blocking the host: a small ring of pinned host slots, each copied with
    ``non_blocking=True`` and guarded by an event, so back-to-back ``solve(k)`` calls queue behind each other on the
    stream instead of waiting for the previous replay (a pageable ``copy_`` is a cudaMemcpyAsync + stream
    synchronise).  nb:185: ``k**2`` is evaluated on the host."""

    def __init__(self, k_dev, depth=8):
        self.k_dev = k_dev
        self.slots = torch.empty((depth, 2), dtype=torch.float64).pin_memory()
        self.events = [None] * depth
        self.i = 0

    def push(self, k):
        i = self.i
        self.i = (i + 1) % len(self.events)
        if self.events[i] is not None:
            self.events[i].synchronize()  # (only when the ring has wrapped onto a copy that is still queued)
        self.slots[i, 0] = float(k)
        self.slots[i, 1] = float(k**2)
        self.k_dev.copy_(self.slots[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


class LargeMeshPlan:
    """One long mesh, many solves (BASELINE configs 2 and 5: N = 10^6, a new wavenumber per call).  At that size the
    work is 0.16 GB -- tens of microseconds of HBM time -- and the path is bound by its eight dependent kernel launches
    (three reduction levels, the first of which also writes the bands of A; the coarse cluster solve; three
    back-substitution levels).  The plan owns the
    buffers and a captured CUDA graph of the whole sequence; ``solve(k)`` updates (k, k**2) in device memory and
    replays it.  Results are bit-identical to ``solve_helmholtz_sweep(N, [k])``.

    ``fused=True``: A is never stored -- the rows are generated (``solve_helmholtz_sweep(..., fused=True)``); on the
    reference's uniform mesh that is the periodic recursion of fh_tridiag_large.cu: one single-CTA kernel plus one
    streaming launch per large level, 16 B written per node."""

    def __init__(self, N, L=1.0, theta=0.0, *, bc_left="neumann", bc_right="sommerfeld", g_left=1.0, g_right=0.0,
                 use_graph=True, fused=False):
        lib = _lib.require_cuda()
        self.N, self.n, self.L, self.fused = int(N), int(N) + 1, float(L), bool(fused)
        dev = _lib.device()
        self.k_dev = torch.zeros(2, dtype=torch.float64, device=dev)
        self.bands = None if fused else torch.empty((4, 1, self.n), dtype=torch.complex128, device=dev)
        self.U = torch.empty((1, self.n), dtype=torch.complex128, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws_bytes = (lib.fh_tridiag_solve_large_workspace_bytes(self.n) if fused
                    else lib.fh_tridiag_solve_batched_workspace_bytes(self.n, 1))
        self.ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        self._prob = _problem(N, L / N, theta, bc_left, bc_right, g_left, g_right, None)
        self._kw = dict(theta=theta, bc_left=bc_left, bc_right=bc_right, g_left=g_left, g_right=g_right)
        self.graph = None
        self._feed = _WavenumberFeed(self.k_dev)
        self._set_k(1.0)
        self._launch()                      # warm-up: module load, function attributes, driver entry points
        torch.cuda.synchronize()
        if use_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._launch()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._launch()

    def _set_k(self, k):
        self._feed.push(k)

    def _launch(self):
        lib, B = _lib.load(), self.bands
        if self.fused:
            check(lib.fh_helmholtz1d_solve_large_fused_kdev_c128(C.byref(self._prob), ptr(self.k_dev), ptr(self.U),
                                                                 ptr(self.status), ptr(self.ws), self.ws.numel(),
                                                                 stream_ptr()),
                  "fh_helmholtz1d_solve_large_fused_kdev_c128")
            return
        if self.n > lib.fh_tridiag_batched_max_n():  # A written by the first reduction level of the solver
            check(lib.fh_helmholtz1d_assemble_solve_large_c128(
                C.byref(self._prob), 0.0, 0.0, ptr(self.k_dev), ptr(B[0]), ptr(B[1]), ptr(B[2]), ptr(B[3]), ptr(self.U),
                ptr(self.status), ptr(self.ws), self.ws.numel(), stream_ptr()), "fh_helmholtz1d_assemble_solve_large_c128")
            return
        check(lib.fh_assemble_1d_c128(C.byref(self._prob), ptr(self.k_dev[0:1]), ptr(self.k_dev[1:2]), 1, ptr(B[0]),
                                      ptr(B[1]), ptr(B[2]), ptr(B[3]), stream_ptr()), "fh_assemble_1d_c128")
        check(lib.fh_tridiag_solve_batched_c128(ptr(B[0]), ptr(B[1]), ptr(B[2]), ptr(B[3]), ptr(self.U), self.n, 1,
                                                ptr(self.status), ptr(self.ws), self.ws.numel(), stream_ptr()),
              "fh_tridiag_solve_batched_c128")

    def solve(self, k):
        """Assemble and solve for wavenumber ``k``; returns the device tensor complex128[N+1] (valid until the next
        call).  Stream-ordered on the current stream; ``status`` is checked by :meth:`solution`."""
        self._set_k(k)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch()
        return self.U[0]

    def backward_error(self):
        """Normwise backward error ||b - A u|| / (||A||_inf ||u|| + ||b||) of the last solve, measured on the device:
        the rows are re-generated (bit-identical to the assembled bands), 16 B per node read.  Synchronises."""
        lib = _lib.load()
        if getattr(self, "_res", None) is None:
            dev = self.U.device
            self._res = (torch.empty(lib.fh_helmholtz1d_residual_workspace_bytes(), dtype=torch.uint8, device=dev),
                         torch.empty(4, dtype=torch.float64, device=dev))
        ws, out = self._res
        check(lib.fh_helmholtz1d_residual_c128(C.byref(self._prob), 0.0, 0.0, ptr(self.k_dev), 0, self.n, ptr(self.U),
                                               None, None, ptr(out), ptr(ws), ws.numel(), stream_ptr()),
              "fh_helmholtz1d_residual_c128")
        r2, u2, b2, an = out.tolist()
        return float(np.sqrt(r2) / (an * np.sqrt(u2) + np.sqrt(b2)))

    def solution(self, k):
        """``solve_helmholtz_fem(N, k)`` for this mesh: ``(nodes, u)`` as numpy arrays."""
        u = self.solve(k)
        out = u.cpu().numpy()
        if int(self.status.item()) != 0:
            # rare: the probe of the unpivoted solve found a defect, or it broke down -- take the general staged path
            # (refinement, verification against the 1e-13 bar, pivoted re-solve; LinAlgError only if singular)
            return solve_helmholtz_fem(self.N, k, L=self.L, **self._kw)
        return np.linspace(0, self.L, self.N + 1), out


def solve_helmholtz_fem(N, k, order=0, L=1.0, *, theta=0.0, bc_left="neumann", bc_right="sommerfeld",
                        g_left=1.0, g_right=0.0, fused=False):
    """nb:190-205 -> ``(nodes, u)`` as numpy arrays (``order`` is accepted and ignored, as in the
    reference).  create_mesh -> device assembly -> device solve -> one D2H copy of ``u``."""
    nodes, U = solve_helmholtz_sweep(N, [k], L, theta, bc_left=bc_left, bc_right=bc_right,
                                     g_left=g_left, g_right=g_right, fused=fused)
    return nodes, U[0].cpu().numpy()


def analytical_solution(x, k):
    """nb:208-219: the continuum solution ``-exp(i k x)`` (verification helper, evaluated where ``x``
    lives: numpy in -> numpy out, torch in -> torch out)."""
    if isinstance(x, torch.Tensor):
        return -torch.exp(1j * k * x.to(torch.complex128))
    return -np.exp(1j * k * x)


class HostSweepPlan:
    """Reusable buffers for :func:`solve_helmholtz_sweep_host`: a pinned host array for the result, and two
    sets of device buffers so that chunk i+1 is assembled+solved while chunk i is copied back.  (Pinning
    gigabytes of host memory costs about a second, so a caller that sweeps repeatedly keeps the plan.)

    ``reduce="l2"``: the sweep returns, per wavenumber, the 16-byte sum inside ``compute_L2_error`` (nb:259-281) instead
    of the solution -- what the reference's own sweep driver ``convergence_test`` (nb:284-362) consumes; no solution ever
    crosses PCIe and the plan holds no n_k x (N+1) host buffer.  ``pipeline``: see :func:`solve_helmholtz_sweep`."""

    def __init__(self, N, n_k, chunk=None, numa_local=True, reduce=None, pipeline=None, fused=False):
        lib = _lib.require_cuda()
        if reduce not in (None, "l2"):
            raise ValueError("reduce must be None or 'l2'")
        # pinned buffers on the GPU's own NUMA node (matters with one process per GPU on a two-socket host)
        self.cpu_cores = _lib.bind_host_thread_near_device() if numa_local else None
        self.N, self.n, self.n_k, self.reduce = int(N), int(N) + 1, int(n_k), reduce
        self.pipeline = _pick_pipeline(pipeline, fused, N, False)
        if chunk is None:  # ~256 MB of solution per chunk keeps both the copy engine and the SMs busy
            chunk = max(1, min(self.n_k, (256 << 20) // (16 * self.n)))
        self.chunk = int(chunk)
        dev = _lib.device()
        self.k_host = torch.empty((2, self.n_k), dtype=torch.float64).pin_memory()
        self.k_dev = torch.empty((2, self.n_k), dtype=torch.float64, device=dev)
        slots = 2
        self.bands = ([torch.empty((4, self.chunk, self.n), dtype=torch.complex128, device=dev) for _ in range(slots)]
                      if self.pipeline != "fused" else None)
        self.U_dev = [torch.empty((self.chunk, self.n), dtype=torch.complex128, device=dev) for _ in range(slots)]
        self.status = torch.zeros(self.n_k, dtype=torch.int32, device=dev)
        self.status_host = torch.empty(self.n_k, dtype=torch.int32).pin_memory()
        self.copy_stream = torch.cuda.Stream()
        self.h2d_bytes = self.k_host.numel() * 8
        if reduce is None:
            self.U_host = torch.empty((self.n_k, self.n), dtype=torch.complex128).pin_memory()
            self.d2h_bytes = self.U_host.numel() * 16 + self.n_k * 4
        else:
            self.U_host = None
            self.nodes_dev = None  # set per call (depends on L)
            self.sums_dev = torch.empty(self.n_k, dtype=torch.complex128, device=dev)
            self.sums_host = torch.empty(self.n_k, dtype=torch.complex128).pin_memory()
            self.d2h_bytes = self.n_k * 16 + self.n_k * 4
        if self.pipeline != "fused" and self.n <= lib.fh_tridiag_batched_max_n():
            cap = max(256, self.chunk // 4)  # systems one stream-ordered refinement call can take
            nb = lib.fh_tridiag_refine_flagged_workspace_bytes(self.n, self.chunk, cap)
            self.refine_ws = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(slots)]
            nb = lib.fh_tridiag_repair_flagged_workspace_bytes(self.n, self.chunk, 8)
            self.repair_ws = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(slots)]
        else:
            self.refine_ws = self.repair_ws = None


def solve_helmholtz_sweep_host(N, ks, L=1.0, theta=0.0, *, plan=None, fused=False, pipeline=None, reduce=None,
                               check_singular=True, **bc):
    """Host-to-host wavenumber sweep: ``ks`` (numpy / list) in, ``(nodes, U)`` out with ``U`` a numpy
    complex128[n_k, N+1] array (a view of the plan's pinned buffer).  Per call: one pinned H2D copy of
    (k, k**2), chunked assemble+solve on the compute stream, and D2H copies of the finished chunks on a
    second stream, overlapped with the next chunk's kernels.  Flagged systems are refined and breakdowns repaired on
    the device inside the pipeline (no host round trip); the status array is read once at the end.

    ``reduce="l2"``: returns ``(nodes, sums)`` instead, ``sums`` = numpy complex128[n_k], the sum inside
    ``compute_L2_error`` (nb:259-281; ``np.sqrt`` of it is the reference's return value) of every solution against
    ``analytical_solution`` -- the only thing the reference's ``convergence_test`` (nb:284-362) does with ``u``.  The
    solutions stay on the device; 20 bytes per wavenumber cross PCIe."""
    lib = _lib.require_cuda()
    h, nodes = L / N, np.linspace(0, L, N + 1)
    ks_h, k2s_h, _ = _host_k_and_k2(ks)
    n_k = len(ks_h)
    if plan is None:
        plan = HostSweepPlan(N, n_k, reduce=reduce, pipeline=pipeline, fused=fused)
    assert plan.N == int(N) and plan.n_k == n_k, "plan was built for another sweep shape"
    reduce, pipeline = plan.reduce, plan.pipeline
    bcs = dict(bc_left=bc.get("bc_left", "neumann"), bc_right=bc.get("bc_right", "sommerfeld"),
               g_left=bc.get("g_left", 1.0), g_right=bc.get("g_right", 0.0))
    prob = _problem(N, h, theta, bcs["bc_left"], bcs["bc_right"], bcs["g_left"], bcs["g_right"], None)
    main = torch.cuda.current_stream()
    plan.k_host[0].copy_(torch.from_numpy(ks_h))
    plan.k_host[1].copy_(torch.from_numpy(k2s_h))
    plan.k_dev.copy_(plan.k_host, non_blocking=True)
    if reduce == "l2":
        plan.nodes_dev = torch.from_numpy(nodes).to(plan.k_dev.device)
        gp = np.ascontiguousarray(gauss_point, dtype=np.float64)
        gw = np.ascontiguousarray(gauss_weight, dtype=np.float64)
    done = [None, None]
    for c, lo in enumerate(range(0, n_k, plan.chunk)):
        hi = min(n_k, lo + plan.chunk)
        m, slot = hi - lo, c & 1
        if done[slot] is not None:
            main.wait_event(done[slot])  # the copy engine has drained this slot's previous chunk
        ks_d, k2s_d = plan.k_dev[0, lo:hi], plan.k_dev[1, lo:hi]
        U, st = plan.U_dev[slot][:m], plan.status[lo:hi]
        if pipeline == "fused":
            check(lib.fh_helmholtz1d_sweep_fused_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), m, ptr(U), ptr(st),
                                                      None, 0, stream_ptr()), "fh_helmholtz1d_sweep_fused_c128")
        else:
            # (a chunk smaller than the plan's takes the head of the slot's buffers: same pointers every call)
            B = plan.bands[slot].view(-1)[:4 * m * plan.n].view(4, m, plan.n)
            dl, d, du, b = (B[i] for i in range(4))
            if pipeline == "one_pass":
                check(lib.fh_helmholtz1d_sweep_assemble_solve_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), m, ptr(dl),
                                                                   ptr(d), ptr(du), ptr(b), ptr(U), ptr(st), None, 0,
                                                                   stream_ptr()),
                      "fh_helmholtz1d_sweep_assemble_solve_c128")
            else:
                check(lib.fh_assemble_1d_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), m, ptr(dl), ptr(d), ptr(du),
                                              ptr(b), stream_ptr()), "fh_assemble_1d_c128")
                check(lib.fh_tridiag_solve_batched_c128(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(U), plan.n, m, ptr(st),
                                                        None, 0, stream_ptr()), "fh_tridiag_solve_batched_c128")
            if plan.refine_ws is not None:
                # stream-ordered, no host round trip; whatever one call cannot take is finished below
                refine_flagged(dl, d, du, b, U, st, sync=False, ws=plan.refine_ws[slot])
                repair_flagged(dl, d, du, b, U, st, sync=False, ws=plan.repair_ws[slot])
        if reduce == "l2":
            check(lib.fh_l2_error_1d_c128(ptr(plan.nodes_dev), ptr(U), ptr(ks_d), int(N), m,
                                          gp.ctypes.data_as(C.c_void_p), gw.ctypes.data_as(C.c_void_p), len(gp),
                                          ptr(plan.sums_dev[lo:hi]), stream_ptr()), "fh_l2_error_1d_c128")
            continue
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(plan.copy_stream):
            plan.copy_stream.wait_event(ready)
            plan.U_host[lo:hi].copy_(U, non_blocking=True)
            done[slot] = torch.cuda.Event()
            done[slot].record(plan.copy_stream)
    if reduce == "l2":
        plan.sums_host.copy_(plan.sums_dev, non_blocking=True)
    main.wait_stream(plan.copy_stream)
    plan.status_host.copy_(plan.status, non_blocking=True)
    main.synchronize()  # everything above has finished
    status_host = plan.status_host
    left = torch.nonzero(status_host != 0).flatten()
    if left.numel() > 0:
        # flagged systems of the fused variant (which keeps no bands), more flagged systems in a chunk than its
        # refinement workspace holds, or systems singular to working precision: solve those wavenumbers again through
        # the staged path (refinement + pivoted repair) and patch their part of the result
        sel = left.numpy()
        bands = assemble_bands(N, (ks_h[sel], k2s_h[sel]), h, theta, **bcs)
        u_fix, st_fix = solve_tridiagonal(*bands)
        if reduce == "l2":
            plan.sums_host[left] = torch.from_numpy(l2_error_sums(ks_h[sel], N, nodes, u_fix))
        else:
            plan.U_host[left] = u_fix.cpu()
        status_host[left] = st_fix.cpu()
        plan.status.copy_(status_host)
    if check_singular and bool(torch.any((status_host & STATUS_NONFINITE) != 0)):
        bad = torch.nonzero(status_host & STATUS_NONFINITE).flatten()[:8].tolist()
        raise np.linalg.LinAlgError(f"Singular matrix (zero pivot / non-finite solution) in systems {bad}")
    return nodes, (plan.sums_host.numpy() if reduce == "l2" else plan.U_host.numpy())


def l2_error_sums(k, N, nodes, u_num, gauss_points=gauss_point, gauss_weights=gauss_weight):
    """Device evaluation of the sum inside ``compute_L2_error`` (nb:259-281) for one or many wavenumbers:
    ``k`` scalar or [n_k], ``u_num`` [N+1] or [n_k, N+1] (numpy or CUDA tensor).  Returns numpy complex128[n_k]."""
    lib = _lib.require_cuda()
    ks = np.atleast_1d(np.asarray(k.detach().cpu().numpy() if isinstance(k, torch.Tensor) else k, dtype=np.float64))
    u_d = _lib.as_device(u_num, torch.complex128).reshape(len(ks), int(N) + 1)
    x_d = _lib.as_device(np.asarray(nodes, dtype=np.float64), torch.float64)
    ks_d = _lib.as_device(ks, torch.float64)
    gp = np.ascontiguousarray(gauss_points, dtype=np.float64)
    gw = np.ascontiguousarray(gauss_weights, dtype=np.float64)
    out = torch.empty(len(ks), dtype=torch.complex128, device=u_d.device)
    check(lib.fh_l2_error_1d_c128(ptr(x_d), ptr(u_d), ptr(ks_d), int(N), len(ks), gp.ctypes.data_as(C.c_void_p),
                                  gw.ctypes.data_as(C.c_void_p), len(gp), ptr(out), stream_ptr()), "fh_l2_error_1d_c128")
    return out.cpu().numpy()


def compute_L2_error(k, N, elements, nodes, u_num, gauss_points=gauss_point, gauss_weights=gauss_weight):
    """nb:259-281 on the device, quirks included (complex square without abs, weight h*w, complex sqrt).
    ``elements`` is accepted for signature parity (the connectivity is the canonical one)."""
    _check_elements(int(N), elements)
    return np.sqrt(l2_error_sums(k, N, nodes, u_num, gauss_points, gauss_weights)[0])


def convergence_test(k_values, N_values, L=1.0, verbose=True):
    """nb:284-331 without the plotting: for every N one batched device sweep over ``k_values``, the device L2
    error, then the reference's log-log fit (it takes sqrt twice, nb:281 and nb:314 -- kept).  Prints the
    reference's lines and returns ``{k: slope}``."""
    errs = {k: [] for k in k_values}
    hs = []
    for N in N_values:
        h, nodes, _ = create_mesh(N)            # nb:300: note the reference ignores L here
        hs.append(h)
        _, U = solve_helmholtz_sweep(N, list(k_values))
        sums = l2_error_sums(np.array([float(k) for k in k_values]), N, nodes, U)
        for k, s in zip(k_values, sums):
            errs[k].append(np.sqrt(np.sqrt(s)))
    rates = {}
    for k in k_values:
        slope, _ = np.polyfit(np.log(hs), np.log(errs[k]), 1)
        rates[k] = slope
        if verbose:
            print(f"Convergence rate for k = {k}: {slope:.2f}")
    return rates
