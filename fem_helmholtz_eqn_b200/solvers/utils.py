"""Closed-form annulus solutions the 2D classes measure their error against (``src/solvers/utils.py``):

    get_solution                         :5-73    Neumann inner circle + ABC of order 0..3 on the outer one ("default")
    get_analytical_solution_sommerfeld   :76-150  u = 1 on the inner circle + ABC on the outer one ("dirichlet_sommerfeld")

Verification-only special functions, evaluated on the host with scipy exactly as the reference does (per Fourier mode a
2 x 2 system in Bessel functions).  Same names, argument order and -- on purpose -- the same order handling: the
Sommerfeld variant knows the orders 0, 1, 2 and **4** (its last branch tests ``order == 4`` where the third-order
coefficient is meant, utils.py:127), so ``abc_order=3``, the classes' default, ends in ``UnboundLocalError`` there as it
does in the reference; ``get_solution`` does the same for an order outside 0..3.  Both accept scalars or arrays.
"""
import numpy as np
from scipy.special import j0, jv, jvp, y0, yv, yvp


def _abc(order_terms, k, R):
    """ik, ik - 1/(2R), + i/(8kR^2), - 1/(8kR^2) + i/(8k^2R^3): the outer-boundary coefficient with ``order_terms``
    of its terms (1..4), written term by term as utils.py:29-62 / :95-140 spell it out."""
    c = 1j * k
    if order_terms >= 2:
        c = 1j * k - 1 / (2 * R)
    if order_terms == 3:
        c = 1j * k - (1.0 / (2.0 * R)) + (1j / (8 * k * R**2))
    if order_terms == 4:
        c = 1j * k - (1.0 / (2.0 * R)) - (1 / (8 * k * R**2)) + (1j / (8 * k**2 * R**3))
    return c


def get_solution(x, y, order, k_squared, inner_radius, outer_radius, n_fourier) -> complex:
    """utils.py:5-73: sum over the Fourier modes -n_fourier..n_fourier of (A_n J_n(kr) + B_n Y_n(kr)) e^{in theta} with
    A_n, B_n from  d/dr = -i^n J_n'(ka) at r = a  and the ABC row at r = b."""
    r = np.sqrt(x**2 + y**2)
    theta = np.arctan2(y, x)
    k = np.sqrt(k_squared)
    a, b = inner_radius, outer_radius
    u_ex = 0.0 + 0.0j
    for n in range(-n_fourier, n_fourier + 1):
        rhs = -(1j**n) * jvp(n, k * a)
        if order == 0:      # (utils.py:19-28: this branch is written without the factor k)
            outer = [jvp(n, k * b) - 1j * jv(n, k * b), yvp(n, k * b) - 1j * yv(n, k * b)]
        elif order in (1, 2, 3):
            c = _abc(order + 1, k, b)
            outer = [k * jvp(n, k * b) - c * jv(n, k * b), k * yvp(n, k * b) - c * yv(n, k * b)]
        else:
            raise UnboundLocalError("cannot access local variable 'M' where it is not associated with a value")
        M = np.array([[jvp(n, k * a), yvp(n, k * a)], outer])
        An, Bn = np.linalg.solve(M, [rhs, 0])
        u_ex = u_ex + (An * jv(n, k * r) + Bn * yv(n, k * r)) * np.exp(1j * n * theta)
    return u_ex


def get_analytical_solution_sommerfeld(x, y, order, k_squared, inner_radius, outer_radius):
    """utils.py:76-150: A J0(kr) + B Y0(kr) with u = 1 on the inner circle -- which the reference hard-wires at r = 1
    (``j0(k)``, ``y0(k)``, utils.py:86, whatever ``inner_radius`` is) -- and the ABC row at r = outer_radius."""
    r = np.sqrt(x**2 + y**2)
    k = np.sqrt(k_squared)
    R = outer_radius
    if order == 0:
        outer = [jvp(0, k * R) - 1j * j0(k * R), yvp(0, k * R) - 1j * y0(k * R)]
    elif order in (1, 2, 4):
        c = _abc({1: 2, 2: 3, 4: 4}[order], k, R)
        outer = [k * jvp(0, k * R) - c * j0(k * R), k * yvp(0, k * R) - c * y0(k * R)]
    else:
        raise UnboundLocalError("cannot access local variable 'matrix' where it is not associated with a value")
    matrix = np.array([[j0(k), y0(k)], outer])
    A, B = np.linalg.solve(matrix, np.array([1, 0]))
    return A * j0(k * r) + B * y0(k * r)
