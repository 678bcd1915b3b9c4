"""ORACLE (test infrastructure, never imported by the product).

Literal CPU restatement of the reference's 1D P1 Helmholtz solver, i.e. code cell 2
of ``notebooks/1D_soln.ipynb`` (``nb:N`` = raw JSON line N of that file).  It keeps
the reference's *cost model* (four dense (N+1)^2 complex arrays, a Python scatter
loop, three dense passes, LAPACK zgesv) so that it can stand in for the reference
when the reference's CPU path is timed, and the reference's *operation order* so
that every assembled value is bit-identical.

Pinned by tests/test_oracle_golden.py against vectors produced by executing the
notebook cell itself (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

# nb:52 -- np.polynomial.legendre.leggauss(1) == ([0.], [2.]); kept as literals so
# the oracle does not depend on numpy's quadrature tables.
GAUSS_POINT_1 = 0.0
GAUSS_WEIGHT_1 = 2.0


def create_mesh(N, L=1.0):
    """nb:33-47: uniform mesh; h is the *scalar* L/N (not a nodal difference)."""
    h = L / N
    nodes = np.linspace(0, L, N + 1)
    elements = np.stack([np.arange(N), np.arange(1, N + 1)], axis=1)
    return h, nodes, elements


def shape_functions(xi):
    """nb:55-67: P1 basis on [-1, 1] and its xi-derivatives."""
    return np.array([(1 - xi) / 2, (1 + xi) / 2]), np.array([-0.5, 0.5])


def element_matrices(h, xi=GAUSS_POINT_1, w=GAUSS_WEIGHT_1):
    """nb:70-96: 2x2 stiffness / mass by one Gauss point.

    Operation order matters for bit parity: ``((g_i * g_j) * jac) * w`` with
    ``g = dphi/dxi / jac`` and ``jac = h / 2``.
    """
    phi, dphi = shape_functions(xi)
    jac = h / 2
    g = dphi / jac
    Ke = np.empty((2, 2))
    Me = np.empty((2, 2))
    for i in (0, 1):
        for j in (0, 1):
            Ke[i, j] = g[i] * g[j] * jac * w
            Me[i, j] = phi[i] * phi[j] * jac * w
    return Ke, Me


def boundary_matrices(h, k, theta=0.0):
    """nb:99-131: Neumann load at xi=-1 and Sommerfeld 2x2 at xi=+1 (h unused)."""
    flux = -1j * k * np.cos(theta)  # nb:113
    left, _ = shape_functions(-1.0)
    right, _ = shape_functions(1.0)
    Ne = np.zeros(2, dtype=complex)
    Se = np.zeros((2, 2), dtype=complex)
    for i in (0, 1):
        Ne[i] = left[i] * flux
        for j in (0, 1):
            Se[i, j] = 1j * k * right[i] * right[j]
    return Ne, Se


def assembly(N, k, h, elements, theta=0.0):
    """nb:134-187: dense global scatter, then ``A = -K + (k**2)*M + S``."""
    n = N + 1
    K = np.zeros((n, n), dtype=complex)
    M = np.zeros((n, n), dtype=complex)
    S = np.zeros((n, n), dtype=complex)
    b = np.zeros(n, dtype=complex)
    Ke, Me = element_matrices(h)  # nb:158: once, h is uniform
    for e in range(N):  # nb:161-168
        conn = elements[e]
        for i in (0, 1):
            for j in (0, 1):
                K[conn[i], conn[j]] += Ke[i, j]
                M[conn[i], conn[j]] += Me[i, j]
    Ne, Se = boundary_matrices(h, k, theta)
    first, last = elements[0], elements[-1]
    for i in (0, 1):
        b[first[i]] += Ne[i]  # nb:174-176
        for j in (0, 1):
            S[last[i], last[j]] += Se[i, j]  # nb:179-182
    A = -K + (k**2) * M + S  # nb:185
    return A, b


def apply_dirichlet_rows(A, b, left=None, right=None):
    """Row replacement exactly as fem2d_dirichlet.py:20-27 does it in 2D:
    ``A[row] = 0; A[row, row] = 1; b[row] = g``.  Columns are not eliminated.
    ``left`` / ``right`` are the prescribed values at x=0 / x=L (None = keep row).
    """
    if left is not None:
        A[0] = 0
        A[0, 0] = 1
        b[0] = left
    if right is not None:
        A[-1] = 0
        A[-1, -1] = 1
        b[-1] = right
    return A, b


def solve_helmholtz_fem(N, k, order=0, L=1.0, dirichlet=(None, None)):
    """nb:190-205 (``order`` accepted and ignored, as in the reference).

    ``dirichlet=(g0, gL)`` adds the 2D classes' row-replacement rule, which is how
    BASELINE config 1 ("Dirichlet BCs") is defined.
    """
    h, nodes, elements = create_mesh(N, L)
    A, b = assembly(N, k, h, elements)
    A, b = apply_dirichlet_rows(A, b, *dirichlet)
    return nodes, np.linalg.solve(A, b)  # nb:204 (LAPACK zgesv)


def analytical_solution(x, k):
    """nb:208-219: continuum solution -exp(ikx)."""
    return -np.exp(1j * k * x)


def compute_L2_error(k, N, elements, nodes, u_num, gauss_points=(GAUSS_POINT_1,),
                     gauss_weights=(GAUSS_WEIGHT_1,)):
    """nb:259-281, quirks included: the *complex* error is squared without abs()
    (nb:277), weighted by h*w (not h/2*w), and the result is a complex sqrt."""
    total = 0.0
    for e in range(N):
        conn = elements[e, :]
        xe = nodes[conn]
        h = xe[-1] - xe[0]
        ue = u_num[conn]
        acc = 0.0
        for xi, wi in zip(gauss_points, gauss_weights):
            phi, _ = shape_functions(xi)
            x = np.dot(phi, xe)
            uh = np.dot(phi, ue)
            acc += (analytical_solution(x, k) - uh) ** 2 * h * wi
        total += acc
    return np.sqrt(total)


def convergence_rates(k_values, N_values, L=1.0):
    """nb:284-331 without the plotting: slope of log(err) vs log(h) per k.
    The reference takes sqrt twice (nb:281 and nb:314); reproduced."""
    errs = {k: [] for k in k_values}
    hs = []
    for N in N_values:
        h, nodes, elements = create_mesh(N)
        hs.append(h)
        for k in k_values:
            _, u = solve_helmholtz_fem(N, k, order=0)
            errs[k].append(np.sqrt(compute_L2_error(k, N, elements, nodes, u)))
    return {k: np.polyfit(np.log(hs), np.log(errs[k]), 1)[0] for k in k_values}
