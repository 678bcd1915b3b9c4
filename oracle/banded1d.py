"""ORACLE (test infrastructure, never imported by the product).

Banded CPU restatement of the reference's 1D system for sizes the literal dense
reference (oracle/ref1d.py) cannot reach (N >= ~2e4 needs 16*N^2 bytes per matrix).

* ``assemble_bands`` reproduces, row by row and in the reference's rounding order,
  exactly the numbers that ``np.diag(A, -1/0/+1)`` of nb:134-187 holds; the test
  suite checks ``==`` against the dense reference for N <= 4096.
* ``solve_bands`` is LAPACK ``zgtsv`` (Gaussian elimination with partial pivoting
  on the tridiagonal) -- the banded equivalent of nb:204's ``zgesv``.
* ``discrete_exact`` is the closed-form solution of the *discrete* scheme (derived
  in SURVEY.md section 8c); it is the only oracle that exists at N = 1e6 .. 2^27.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack

from . import ref1d

NEUMANN, SOMMERFELD, DIRICHLET = "neumann", "sommerfeld", "dirichlet"


def assemble_bands(N, k, h, theta=0.0, bc_left=NEUMANN, bc_right=SOMMERFELD,
                   g_left=1.0, g_right=0.0, x=None):
    """Tridiagonal bands of nb:134-187 (+ fem2d_dirichlet.py:20-27 row rule).

    Returns ``dl, d, du, b`` as complex128[N+1]; ``dl[0]`` and ``du[N]`` are 0.
    ``x`` (fp64[N+1]) switches to per-element ``h_e = x[e+1]-x[e]`` ("nodal-h" mode,
    an extension; the reference only has the scalar-h mode).
    """
    n = N + 1
    k2 = k**2  # nb:185 evaluates k**2 on the host; keep its exact value
    if x is None:
        Ke, Me = ref1d.element_matrices(h)
        ke = np.broadcast_to(Ke, (N, 2, 2))
        me = np.broadcast_to(Me, (N, 2, 2))
    else:
        pairs = [ref1d.element_matrices(float(x[e + 1] - x[e])) for e in range(N)]
        ke = np.array([p[0] for p in pairs])
        me = np.array([p[1] for p in pairs])
    # nb:161-168: K[j,j] = 0 + Ke(e=j-1)[1,1] + Ke(e=j)[0,0]  (element order)
    Kd = np.zeros(n)
    Md = np.zeros(n)
    Kd[1:] += ke[:, 1, 1]
    Md[1:] += me[:, 1, 1]
    Kd[:-1] += ke[:, 0, 0]
    Md[:-1] += me[:, 0, 0]
    d = (-Kd + k2 * Md).astype(complex)          # nb:185
    du = np.zeros(n, dtype=complex)
    dl = np.zeros(n, dtype=complex)
    du[:-1] = -ke[:, 0, 1] + k2 * me[:, 0, 1]
    dl[1:] = -ke[:, 1, 0] + k2 * me[:, 1, 0]
    Ne, Se = ref1d.boundary_matrices(h, k, theta)
    b = np.zeros(n, dtype=complex)
    b[0] += Ne[0]                                  # nb:174-176 (Ne[1] == 0)
    b[1] += Ne[1]
    d[-1] += Se[1, 1]                              # nb:179-182 (only Se[1,1] != 0)
    if bc_left == DIRICHLET:                       # fem2d_dirichlet.py:20-22
        d[0], du[0], b[0] = 1.0, 0.0, g_left
    elif bc_left != NEUMANN:
        raise ValueError(bc_left)
    if bc_right == DIRICHLET:                      # fem2d_dirichlet.py:25-27
        d[-1], dl[-1], b[-1] = 1.0, 0.0, g_right
    elif bc_right != SOMMERFELD:
        raise ValueError(bc_right)
    return dl, d, du, b


def to_dense(dl, d, du):
    return np.diag(d) + np.diag(dl[1:], -1) + np.diag(du[:-1], 1)


def solve_bands(dl, d, du, b):
    """LAPACK zgtsv on copies; raises LinAlgError like np.linalg.solve would."""
    if len(d) == 1:
        if d[0] == 0:
            raise np.linalg.LinAlgError("singular 1x1 system")
        return np.asarray(b, dtype=complex) / d[0]
    _, _, _, x, info = lapack.zgtsv(dl[1:].copy(), d.copy(), du[:-1].copy(), b.copy())
    if info != 0:
        raise np.linalg.LinAlgError(f"zgtsv info={info}")
    return x


def solve_helmholtz_banded(N, k, L=1.0, theta=0.0, **bc):
    h = L / N
    nodes = np.linspace(0, L, N + 1)
    return nodes, solve_bands(*assemble_bands(N, k, h, theta, **bc))


def residual_metrics(dl, d, du, b, u):
    """Normwise backward error ||b-Au||_2 / (||A||_inf ||u||_2 + ||b||_2) and
    plain relative residual ||b-Au||_2/||b||_2 (gate G3)."""
    r = b - d * u
    r[1:] -= dl[1:] * u[:-1]
    r[:-1] -= du[:-1] * u[1:]
    rows = np.abs(d).copy()
    rows[1:] += np.abs(dl[1:])
    rows[:-1] += np.abs(du[:-1])
    rn, un, bn = np.linalg.norm(r), np.linalg.norm(u), np.linalg.norm(b)
    return rn / (rows.max() * un + bn), rn / bn if bn > 0 else np.inf


def discrete_exact(N, k, L=1.0, bc_left=NEUMANN, bc_right=SOMMERFELD):
    """Exact solution of the reference's discrete scheme (theta_inc = 0, g_left=1,
    g_right=0).  With a = 1/h + k^2 h/4, the interior row a*u[j-1] + (-2/h + k^2 h/2)
    *u[j] + a*u[j+1] = 0 has solutions exp(+-i j t), t = 2*atan(k h / 2)."""
    h = L / N
    t = 2.0 * np.arctan(k * h / 2.0)
    j = np.arange(N + 1, dtype=np.float64)
    if bc_left == NEUMANN and bc_right == SOMMERFELD:
        return -np.exp(1j * j * t)
    if bc_left == DIRICHLET and bc_right == SOMMERFELD:
        return np.exp(1j * j * t)
    if bc_left == DIRICHLET and bc_right == DIRICHLET:
        return (np.sin(t * (N - j)) / np.sin(t * N)).astype(complex)
    raise ValueError((bc_left, bc_right))
