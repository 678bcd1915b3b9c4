"""``"dirichlet"`` solver: u = 1 on the inner circle, u = 0 on the outer one, both by row replacement.
Mirrors ``src/solvers/fem2d_dirichlet.py``."""
import numpy as np
from scipy.special import j0, y0

from .base import BaseSolver


class FEM2DDirichletSolver(BaseSolver):
    _dirichlet_inner = True
    _dirichlet_outer = True

    def get_analytical_solution(self, x, y):
        """fem2d_dirichlet.py:57-85: A J0(kr) + B Y0(kr) with u(r_in) = 1, u(r_out) = 0."""
        r = np.sqrt(x**2 + y**2)
        k = np.sqrt(self.k_squared)
        M = np.array([[j0(k * self.inner_radius), y0(k * self.inner_radius)],
                      [j0(k * self.outer_radius), y0(k * self.outer_radius)]])
        A, B = np.linalg.solve(M, np.array([1, 0]))
        return A * j0(k * r) + B * y0(k * r)
