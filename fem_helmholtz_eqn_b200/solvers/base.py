"""``BaseSolver`` of the 2D annulus solvers on the GPU.

Mirrors ``src/solvers/base.py`` (constructor kwargs, method names) and hosts the machinery the four concrete
classes share: element matrices, boundary-edge terms, the deterministic CSR scatter, Dirichlet row replacement
and the dense device solve -- each a call through the C ABI (``include/fem_helmholtz.h``).  The concrete
classes only say which boundary terms they use, exactly like the reference's copy-paste variants do.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC
from typing import Tuple

import numpy as np
import torch
from scipy.special import jvp, roots_legendre

from .. import _lib
from .._lib import check, ptr, stream_ptr

__all__ = ["BaseSolver", "CsrMatrix"]

_EDGES = ((0, 1), (1, 2), (2, 3), (3, 0))  # fem2d.py:37-42


def _host_f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


class CsrMatrix:
    """Device CSR matrix (``rowptr`` int64[n+1], ``colidx`` int32[nnz], ``vals`` complex128[nnz]) that still looks
    like the reference's dense ``K`` to numpy: ``np.asarray(K)``, ``K.shape``, ``K[i, j]`` (via the dense copy)."""

    def __init__(self, rowptr, colidx, vals, n):
        self.rowptr, self.colidx, self.vals, self.n = rowptr, colidx, vals, int(n)
        self.shape = (self.n, self.n)
        self.dtype = np.dtype(np.complex128)

    @property
    def nnz(self):
        return int(self.vals.numel())

    def to_dense(self):
        rp, ci, v = (t.cpu().numpy() for t in (self.rowptr, self.colidx, self.vals))
        A = np.zeros(self.shape, dtype=complex)
        rows = np.repeat(np.arange(self.n), np.diff(rp))
        A[rows, ci] = v
        return A

    def __array__(self, dtype=None, copy=None):
        A = self.to_dense()
        return A.astype(dtype) if dtype is not None else A

    def __getitem__(self, idx):
        return self.to_dense()[idx]

    def coo(self):
        rp, ci, v = (t.cpu().numpy() for t in (self.rowptr, self.colidx, self.vals))
        return np.repeat(np.arange(self.n), np.diff(rp)), ci.astype(np.int64), v


class BaseSolver(ABC):
    # which boundary terms a concrete class assembles (set by the subclasses)
    _neumann_inner = False
    _robin_outer = False
    _robin_gauss = 2
    _dirichlet_inner = False
    _dirichlet_outer = False

    def __init__(self, type: str, eqn, k_squared: float = 5.0, n_fourier: int = 1, abc_order: int = 3,
                 inner_radius: float = 1.0, outer_radius: float = 1.5):
        self.type = type
        self.eqn = eqn
        self.k_squared = k_squared
        self.n_fourier = n_fourier
        self.abc_order = abc_order
        self.inner_radius = inner_radius
        self.outer_radius = outer_radius
        self._dev = None
        self._pattern = {}
        self._F_dev = self._F_host = self._N_dev = self._N_host = None

    # ``F`` / ``N_global`` are numpy arrays in the reference; here they live on the device and come to the host when
    # somebody looks at them (a 10^6-node load vector is 6 ms of pageable D2H copy that solve() never needs)
    @property
    def F(self):
        if self._F_host is None and self._F_dev is not None:
            self._F_host = self._F_dev.cpu().numpy()
        return self._F_host

    @F.setter
    def F(self, value):
        self._F_host = value

    @property
    def N_global(self):
        if self._N_host is None:
            if self._N_dev is None:
                raise AttributeError("N_global is set by assemble() of the solver classes with a Neumann boundary")
            self._N_host = self._N_dev.cpu().numpy()
        return self._N_host

    # ------------------------------------------------------------------ reference helpers (host, 4 numbers)
    def get_shape_functions(self, xi, eta):
        """base.py:32-55 (the kernels carry the same formulas; this host copy exists for API parity)."""
        N = 0.25 * np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)])
        dN_dxi = np.array([-0.25 * (1 - eta), 0.25 * (1 - eta), 0.25 * (1 + eta), -0.25 * (1 + eta)])
        dN_deta = np.array([-0.25 * (1 - xi), -0.25 * (1 + xi), 0.25 * (1 + xi), 0.25 * (1 - xi)])
        return N, dN_dxi, dN_deta

    def get_shape_fuctions_1d(self, xi):
        """base.py:57-63 (sic)."""
        return np.array([0.5 * (1 - xi), 0.5 * (1 + xi)])

    # ------------------------------------------------------------------ device state
    def _device_mesh(self):
        if self._dev is None:
            _lib.require_cuda()
            elements = np.asarray(self.eqn.elements)
            if elements.ndim != 2 or elements.shape[1] != 4:
                raise ValueError("the GPU path assembles 4-node quads; elements must have shape [ne, 4]")
            self._dev = {
                "nodes": _lib.as_device(np.asarray(self.eqn.nodes, dtype=np.float64), torch.float64),
                "elements": _lib.as_device(elements.astype(np.int32), torch.int32),
                "n": int(self.eqn.n_nodes), "ne": int(self.eqn.n_elements),
            }
        return self._dev

    def _element_matrices_device(self, ids=None):
        """base.py:65-117 for all elements (or the listed ones): complex128[count, 4, 4] on the device."""
        lib, m = _lib.require_cuda(), self._device_mesh()
        gp, gw = roots_legendre(2)
        (_, gp_p), (_, gw_p) = _host_f64(gp), _host_f64(gw)
        elements = m["elements"] if ids is None else m["elements"][torch.as_tensor(ids, device=m["elements"].device)]
        elements = elements.contiguous()
        count = elements.shape[0]
        Ke = torch.empty((count, 4, 4), dtype=torch.complex128, device=elements.device)
        check(lib.fh_q4_element_matrices_c128(ptr(m["nodes"]), ptr(elements), count, float(self.k_squared), gp_p, gw_p,
                                              ptr(Ke), stream_ptr()), "fh_q4_element_matrices_c128")
        return Ke

    def get_element_matrices(self, idx) -> Tuple[np.ndarray, np.ndarray]:
        """base.py:65-117 -> ``(K_e[4,4], F_e[4])`` for one element (F_e is identically zero in the reference)."""
        Ke = self._element_matrices_device([int(idx)])[0].cpu().numpy()
        return Ke, np.zeros(4, dtype=complex)

    def _edge_geometry(self, ids, radius):
        lib, m = _lib.require_cuda(), self._device_mesh()
        ids_d = _lib.as_device(np.asarray(ids, dtype=np.int32), torch.int32)
        nb = ids_d.numel()
        on_edge = torch.zeros((nb, 4), dtype=torch.uint8, device=ids_d.device)
        length = torch.zeros((nb, 4), dtype=torch.float64, device=ids_d.device)
        check(lib.fh_q4_edge_geometry(ptr(m["nodes"]), ptr(m["elements"]), ptr(ids_d), nb, float(radius), ptr(on_edge),
                                      ptr(length), stream_ptr()), "fh_q4_edge_geometry")
        return ids_d, on_edge, length

    def _robin_matrices_device(self, ids, n_gauss):
        """fem2d.py:29-80 / fem2d_dirichlet_sommerfeld.py:26-78 for the listed elements: complex128[nb, 4, 4]."""
        lib = _lib.require_cuda()
        _, on_edge, length = self._edge_geometry(ids, self.eqn.outer_radius)
        coef = complex(self.eqn.get_abc_coeff(k_squared=self.k_squared, order=self.abc_order))
        gp, gw = roots_legendre(n_gauss)
        (_, gp_p), (_, gw_p) = _host_f64(gp), _host_f64(gw)
        nb = on_edge.shape[0]
        Se = torch.zeros((nb, 4, 4), dtype=torch.complex128, device=on_edge.device)
        check(lib.fh_q4_robin_edge_matrices_c128(ptr(on_edge), ptr(length), nb, coef.real, coef.imag, gp_p, gw_p,
                                                 int(n_gauss), ptr(Se), stream_ptr()), "fh_q4_robin_edge_matrices_c128")
        return Se

    def get_normal_derivative(self, x, y) -> complex:
        """fem2d.py:12-27: incident-wave normal derivative (Bessel sum; host, verification-grade special functions)."""
        r = np.sqrt(x**2 + y**2)
        theta = np.arctan2(y, x)
        k = np.sqrt(self.k_squared)
        total = complex(0.0, 0.0)
        for m in range(-self.n_fourier, self.n_fourier + 1):
            total += (1j**m) * k * jvp(m, k * r) * np.exp(1j * m * theta)
        return total

    # where the incident-wave load g of the Neumann edges is evaluated: "host" (scipy's Bessel functions, as the
    # reference: the load vector is bit-identical to its) or "device" (fh_q4_neumann_load_c128, CUDA's jn(): ~1e-11)
    neumann_load = "host"

    def _neumann_vectors_device(self, ids):
        """fem2d.py:82-137 for the listed elements: complex128[nb, 4].  The Bessel-function load g is evaluated at the
        Gauss points of the flagged edges -- on the host and uploaded, or by the device kernel
        (``neumann_load = "device"``) --; the edge kernel integrates."""
        lib, m = _lib.require_cuda(), self._device_mesh()
        ids_d, on_edge, length = self._edge_geometry(ids, self.eqn.inner_radius)
        gp, gw = roots_legendre(3)
        (_, gp_p), (_, gw_p) = _host_f64(gp), _host_f64(gw)
        if self.neumann_load not in ("host", "device"):
            raise ValueError("neumann_load must be 'host' or 'device'")
        if self.neumann_load == "device":
            g_d = torch.empty((len(ids), 4, 3), dtype=torch.complex128, device=on_edge.device)
            check(lib.fh_q4_neumann_load_c128(ptr(m["nodes"]), ptr(m["elements"]), ptr(ids_d), len(ids), ptr(on_edge),
                                              float(np.sqrt(self.k_squared)), int(self.n_fourier), gp_p, 3, ptr(g_d),
                                              stream_ptr()), "fh_q4_neumann_load_c128")
        else:
            flags = on_edge.cpu().numpy()
            nodes, elements = np.asarray(self.eqn.nodes), np.asarray(self.eqn.elements)
            g = np.zeros((len(ids), 4, 3), dtype=complex)
            # The Gauss points of the flagged edges, with the reference's scalar arithmetic (fem2d.py:119-123 and the
            # r, theta of fem2d.py:12-16: `x**2` of a NumPy scalar is libm's pow, not the square an array would take) ...
            where, rs, thetas = [], [], []
            shape = [self.get_shape_fuctions_1d(xi) for xi in gp]
            for q, e in enumerate(ids):
                coords = nodes[elements[e]]
                for ed, (p0, p1) in enumerate(_EDGES):
                    if not flags[q, ed]:
                        continue
                    for gi, N in enumerate(shape):
                        x = 0.0 + N[0] * coords[p0, 0] + N[1] * coords[p1, 0]
                        y = 0.0 + N[0] * coords[p0, 1] + N[1] * coords[p1, 1]
                        where.append((q, ed, gi))
                        rs.append(np.sqrt(x**2 + y**2))
                        thetas.append(np.arctan2(y, x))
            # ... and the Bessel sum of get_normal_derivative over all of them at once: the same operations in the same
            # order per point (scipy's jvp and NumPy's complex exp are element-wise loops over the scalar routines), a
            # few array calls instead of (2 n_fourier + 1) scipy calls per point
            if where:
                r_a, th_a = np.array(rs, dtype=np.float64), np.array(thetas, dtype=np.float64)
                k = np.sqrt(self.k_squared)
                total = np.zeros(len(where), dtype=complex)
                for m in range(-self.n_fourier, self.n_fourier + 1):
                    total += (1j**m) * k * jvp(m, k * r_a) * np.exp(1j * m * th_a)
                idx = np.array(where)
                g[idx[:, 0], idx[:, 1], idx[:, 2]] = total
            g_d = _lib.as_device(g, torch.complex128)
        Ne = torch.zeros((len(ids), 4), dtype=torch.complex128, device=g_d.device)
        check(lib.fh_q4_neumann_edge_vectors_c128(ptr(on_edge), ptr(length), len(ids), ptr(g_d), gp_p, gw_p, 3, ptr(Ne),
                                                  stream_ptr()), "fh_q4_neumann_edge_vectors_c128")
        return Ne

    # ------------------------------------------------------------------ symbolic + numeric scatter
    def _segments(self, tag, id_lists, matrix):
        """Stable grouping of every contribution by its (row, col) / node key; cached per pattern."""
        key = (tag, matrix)
        if key in self._pattern:
            return self._pattern[key]
        lib, m = _lib.require_cuda(), self._device_mesh()
        per = 16 if matrix else 4
        dev = m["elements"].device
        counts = [m["ne"] if ids is None else len(ids) for ids in id_lists]
        total = per * sum(counts)
        keys = torch.empty(total, dtype=torch.int64, device=dev)
        off = 0
        for ids, cnt in zip(id_lists, counts):
            if cnt == 0:
                continue
            ids_d = None if ids is None else _lib.as_device(np.asarray(ids, dtype=np.int32), torch.int32)
            check(lib.fh_q4_scatter_keys(ptr(m["elements"]), ptr(ids_d), cnt, m["n"], int(matrix), ptr(keys[off:]),
                                         stream_ptr()), "fh_q4_scatter_keys")
            off += per * cnt
        perm = torch.empty(total, dtype=torch.int32, device=dev)
        ukeys = torch.empty(total, dtype=torch.int64, device=dev)
        seg = torch.empty(total + 1, dtype=torch.int64, device=dev)
        n_unique = torch.zeros(1, dtype=torch.int64, device=dev)
        ws_bytes = lib.fh_sorted_segments_workspace_bytes(total)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib.fh_sorted_segments(ptr(keys), total, ptr(perm), ptr(ukeys), ptr(seg), ptr(n_unique), ptr(ws), ws_bytes,
                                     stream_ptr()), "fh_sorted_segments")
        nu = int(n_unique.item())
        out = {"perm": perm, "ukeys": ukeys[:nu], "seg": seg[:nu + 1], "n_unique": n_unique, "nu": nu,
               "split": per * counts[0], "total": total}
        if matrix:
            rowptr = torch.empty(m["n"] + 1, dtype=torch.int64, device=dev)
            colidx = torch.empty(nu, dtype=torch.int32, device=dev)
            check(lib.fh_csr_from_keys(ptr(ukeys), ptr(n_unique), m["n"], m["n"], ptr(rowptr), ptr(colidx), stream_ptr()),
                  "fh_csr_from_keys")
            out["rowptr"], out["colidx"] = rowptr, colidx
        self._pattern[key] = out
        return out

    def _segment_sum(self, contrib, pat):
        lib = _lib.require_cuda()
        out = torch.zeros(pat["nu"], dtype=torch.complex128, device=contrib.device)
        check(lib.fh_segment_sum_c128(ptr(contrib.contiguous()), ptr(pat["perm"]), ptr(pat["seg"]), ptr(pat["n_unique"]),
                                      pat["nu"], pat["split"], ptr(out), stream_ptr()), "fh_segment_sum_c128")
        return out

    def assemble(self) -> None:
        """``*.assemble()`` (fem2d.py:139-171, fem2d_dirichlet.py:31-45, fem2d_dirichlet_sommerfeld.py:80-103,
        fem2d_neumann_dirichlet.py:99-121): sets ``self.K`` (CsrMatrix), ``self.F`` (numpy) and, where the reference
        has them, ``self.S`` / ``self.N_global``.  Values equal the reference's dense entries bit for bit."""
        m = self._device_mesh()
        Ke = self._element_matrices_device().reshape(-1)
        outer = list(self.eqn.outer_boundary_element_indices) if self._robin_outer else []
        inner = list(self.eqn.inner_boundary_element_indices) if self._neumann_inner else []
        if self._robin_outer:
            Se = self._robin_matrices_device(outer, self._robin_gauss).reshape(-1)
            pat = self._segments("K+S", [None, outer], True)
            vals = self._segment_sum(torch.cat([Ke, Se]), pat)
            # the reference also keeps S on its own (fem2d.py:141): same pattern, K part zeroed -- summed on first
            # access (``self.S``), not on every assemble(): it is a second full scatter that solve() never needs
            self._S_lazy = (Ke.numel(), Se, pat, m["n"])
            self._S = None
        else:
            pat = self._segments("K", [None], True)
            vals = self._segment_sum(Ke, pat)
        self.K = CsrMatrix(pat["rowptr"], pat["colidx"], vals, m["n"])
        F = torch.zeros(m["n"], dtype=torch.complex128, device=vals.device)
        if self._neumann_inner:
            Ne = self._neumann_vectors_device(inner).reshape(-1)
            vpat = self._segments("N", [inner], False)
            sums = self._segment_sum(Ne, vpat)
            F[vpat["ukeys"]] = sums                       # F += N_global  (0 + x == x)
            self._N_dev, self._N_host = F.clone(), None
        self._F_dev, self._F_host = F, None

    @property
    def S(self):
        """Robin boundary matrix of the last ``assemble()`` (fem2d.py:141 keeps it as ``self.S``), on K's pattern."""
        if getattr(self, "_S", None) is None:
            lazy = getattr(self, "_S_lazy", None)
            if lazy is None:
                raise AttributeError("S is set by assemble() of the solver classes with a Robin boundary")
            n_ke, Se, pat, n = lazy
            zeros = torch.zeros(n_ke, dtype=Se.dtype, device=Se.device)
            self._S = CsrMatrix(pat["rowptr"], pat["colidx"], self._segment_sum(torch.cat([zeros, Se]), pat), n)
        return self._S

    def _dirichlet_rows_device(self, tol: float = 1e-10):
        """apply_boundary_conditions(self.K, self.F) without bringing F to the host (what solve() needs)."""
        lib, m = _lib.require_cuda(), self._device_mesh()
        flags = torch.zeros(m["n"], dtype=torch.uint8, device=m["nodes"].device)
        mask = 0
        for on, radius, bit in ((self._dirichlet_inner, self.inner_radius, 1), (self._dirichlet_outer, self.outer_radius, 2)):
            if on:
                check(lib.fh_q4_boundary_flags(ptr(m["nodes"]), m["n"], float(radius), float(tol), bit, ptr(flags),
                                               stream_ptr()), "fh_q4_boundary_flags")
                mask |= bit
        if mask:
            check(lib.fh_csr_apply_dirichlet_c128(ptr(self.K.rowptr), ptr(self.K.colidx), ptr(self.K.vals), ptr(self._F_dev),
                                                  ptr(flags), m["n"], mask, 1.0, 0.0, 0.0, 0.0, stream_ptr()),
                  "fh_csr_apply_dirichlet_c128")
            self._F_host = None

    def apply_boundary_conditions(self, A=None, b=None, tol: float = 1e-10):
        """Dirichlet row replacement (fem2d_dirichlet.py:11-29, fem2d_dirichlet_sommerfeld.py:12-24,
        fem2d_neumann_dirichlet.py:85-97): ``A[row] = 0; A[row, row] = 1; b[row] = g`` for the nodes within ``tol`` of a
        Dirichlet circle (g = 1 inside, 0 outside), applied to the ``(A, b)`` passed in -- in place -- and returned, as the
        reference does.  ``A``: a :class:`CsrMatrix` (device; the default, ``self.K``) or a dense numpy matrix;
        ``b``: numpy / torch vector (default ``self.F``)."""
        lib, m = _lib.require_cuda(), self._device_mesh()
        dev = m["nodes"].device
        flags = torch.zeros(m["n"], dtype=torch.uint8, device=dev)
        mask = 0
        if self._dirichlet_inner:
            check(lib.fh_q4_boundary_flags(ptr(m["nodes"]), m["n"], float(self.inner_radius), float(tol), 1, ptr(flags),
                                           stream_ptr()), "fh_q4_boundary_flags")
            mask |= 1
        if self._dirichlet_outer:
            check(lib.fh_q4_boundary_flags(ptr(m["nodes"]), m["n"], float(self.outer_radius), float(tol), 2, ptr(flags),
                                           stream_ptr()), "fh_q4_boundary_flags")
            mask |= 2
        own = A is None or A is getattr(self, "K", None)
        A = self.K if A is None else A
        if b is None or (own and b is self._F_host):
            b_dev, b_own = self._F_dev, True
        else:
            b_dev, b_own = _lib.as_device(b, torch.complex128).clone(), False
        if not mask:
            return A, (self.F if b_own else b)
        if isinstance(A, CsrMatrix):
            check(lib.fh_csr_apply_dirichlet_c128(ptr(A.rowptr), ptr(A.colidx), ptr(A.vals), ptr(b_dev), ptr(flags),
                                                  m["n"], mask, 1.0, 0.0, 0.0, 0.0, stream_ptr()),
                  "fh_csr_apply_dirichlet_c128")
        else:  # a dense matrix from the caller (the reference's type): same rule, rows picked by the device flags
            b = self.F if b is None else b
            f = flags.cpu().numpy()
            rows = np.nonzero(f & mask)[0]
            A[rows] = 0
            A[rows, rows] = 1
            g = np.where(f[rows] & 1 & mask, 1.0, 0.0)  # inner circle: 1, outer: 0 (an inner flag wins, as in the kernel)
            b_host = np.array(b, dtype=complex) if not isinstance(b, np.ndarray) else b
            b_host[rows] = g
            return A, b_host
        if b_own:
            self._F_host = None        # (the device copy changed; the host view is rebuilt on access)
            return A, self.F
        out = b_dev.cpu().numpy()
        if isinstance(b, np.ndarray):
            b[...] = out
            out = b
        return A, out

    # which device solver np.linalg.solve maps to: "band" (sparse, default) or "dense"; "auto" = band unless the band is
    # no narrower than the matrix
    linear_solver = "auto"

    def _band_plan(self):
        """Bandwidth of K under the mesh's own numbering and, if that is wide (an unstructured gmsh mesh), under a
        reverse Cuthill-McKee ordering of its pattern (host, scipy; the pattern is a few kilobytes).  Cached per
        pattern.  Returns ``(perm_or_None, kl, ku)``."""
        key = ("band", self.K.nnz, self.K.n)
        if key in self._pattern:
            return self._pattern[key]
        lib, n, dev = _lib.require_cuda(), self.K.n, self.K.vals.device
        klku = torch.zeros(2, dtype=torch.int32, device=dev)

        def width(perm):
            check(lib.fh_csr_bandwidth(ptr(self.K.rowptr), ptr(self.K.colidx), ptr(perm), n, ptr(klku), stream_ptr()),
                  "fh_csr_bandwidth")
            kl, ku = (int(v) for v in klku.tolist())
            return kl, ku

        perm, (kl, ku) = None, width(None)
        if 2 * kl + ku + 1 > max(96, n // 6):
            from scipy.sparse import csr_matrix
            from scipy.sparse.csgraph import reverse_cuthill_mckee
            rp, ci = self.K.rowptr.cpu().numpy(), self.K.colidx.cpu().numpy()
            pat = csr_matrix((np.ones(len(ci), dtype=np.int8), ci, rp), shape=(n, n))
            order = reverse_cuthill_mckee((pat + pat.T).tocsr(), symmetric_mode=True)   # order[new] = old
            new_of_old = np.empty(n, dtype=np.int32)
            new_of_old[order] = np.arange(n, dtype=np.int32)
            cand = _lib.as_device(new_of_old, torch.int32)
            kl2, ku2 = width(cand)
            if 2 * kl2 + ku2 < 2 * kl + ku:
                perm, kl, ku = cand, kl2, ku2
        self._pattern[key] = (perm, kl, ku)
        return self._pattern[key]

    def _solve_device(self):
        """np.linalg.solve(K, F) (fem2d.py:176) on the device: banded LU with partial pivoting on the CSR matrix
        (``fh_csr_band_lu_solve_c128``), or -- when the band would be as wide as the matrix -- the dense LU."""
        lib, n = _lib.require_cuda(), self.K.n
        dev = self.K.vals.device
        x = torch.empty(n, dtype=torch.complex128, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        mode = self.linear_solver
        if mode not in ("auto", "band", "dense"):
            raise ValueError("linear_solver must be 'auto', 'band' or 'dense'")
        if mode != "dense":
            perm, kl, ku = self._band_plan()
            if mode == "auto" and 2 * kl + ku + 1 >= n:
                mode = "dense"
        if mode != "dense":
            ws_bytes = lib.fh_band_lu_workspace_bytes(n, kl, ku)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib.fh_csr_band_lu_solve_c128(ptr(self.K.rowptr), ptr(self.K.colidx), ptr(self.K.vals), ptr(self._F_dev),
                                                ptr(perm), n, kl, ku, ptr(x), ptr(ws), ws_bytes, ptr(status), stream_ptr()),
                  "fh_csr_band_lu_solve_c128")
        else:
            A = torch.empty((n, n), dtype=torch.complex128, device=dev)
            check(lib.fh_csr_to_dense_c128(ptr(self.K.rowptr), ptr(self.K.colidx), ptr(self.K.vals), n, ptr(A),
                                           stream_ptr()), "fh_csr_to_dense_c128")
            b = self._F_dev.clone()
            piv = torch.empty(n, dtype=torch.int32, device=dev)
            check(lib.fh_dense_lu_solve_c128(ptr(A), ptr(b), ptr(x), n, ptr(piv), ptr(status), stream_ptr()),
                  "fh_dense_lu_solve_c128")
        if int(status.item()) != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return x

    def solve(self):
        """``*.solve()``: assemble -> (Dirichlet rows) -> solve -> ``(u_real, u_imag)`` numpy arrays."""
        self.assemble()
        if self._dirichlet_inner or self._dirichlet_outer:
            self._dirichlet_rows_device()
        u = self._solve_device().cpu().numpy()
        return np.real(u), np.imag(u)

    def get_analytical_solution(self, x, y):
        raise NotImplementedError("Analytical solution not implemented")

    def get_analytical_solution_ordered(self, x, y, order):
        raise NotImplementedError("Analytical solution not implemented")

    def compute_L2_error(self, u_num_real, u_num_imag):
        """base.py:119-177 on the device: ``fh_q4_gauss_points`` gives the (x, y) of every element's 2 x 2 Gauss points,
        the class's analytic solution -- Bessel closed forms, host scipy, vectorised where the formula allows -- is
        evaluated there, and ``fh_q4_l2_error_c128`` integrates |u_h - u_exact|^2 det J w_i w_j in the reference's
        per-element order (the sum over elements is a fixed tree: agrees with the sequential sum to rounding)."""
        lib, m = _lib.require_cuda(), self._device_mesh()
        gp, gw = roots_legendre(2)
        (_, gp_p), (_, gw_p) = _host_f64(gp), _host_f64(gw)
        dev = m["nodes"].device
        xy = torch.empty((m["ne"], 2, 2, 2), dtype=torch.float64, device=dev)
        check(lib.fh_q4_gauss_points(ptr(m["nodes"]), ptr(m["elements"]), m["ne"], gp_p, 2, ptr(xy), stream_ptr()),
              "fh_q4_gauss_points")
        pts = xy.cpu().numpy().reshape(-1, 2)
        try:
            exact = np.asarray(self.get_analytical_solution(pts[:, 0], pts[:, 1]), dtype=complex)
            if exact.shape != (len(pts),):
                raise ValueError
        except (ValueError, TypeError):  # a formula that only takes scalars
            exact = np.array([self.get_analytical_solution(a, b) for a, b in pts], dtype=complex)
        u = np.asarray(u_num_real) + 1j * np.asarray(u_num_imag)
        u_d, ex_d = _lib.as_device(u, torch.complex128), _lib.as_device(exact, torch.complex128)
        ws = torch.empty(lib.fh_q4_l2_error_workspace_bytes(), dtype=torch.uint8, device=dev)
        out = torch.empty(1, dtype=torch.float64, device=dev)
        check(lib.fh_q4_l2_error_c128(ptr(m["nodes"]), ptr(m["elements"]), m["ne"], ptr(u_d), ptr(ex_d), gp_p, gw_p, 2,
                                      ptr(out), ptr(ws), ws.numel(), stream_ptr()), "fh_q4_l2_error_c128")
        return np.sqrt(float(out.item()))
