"""``"neumann_dirichlet"`` solver: Neumann load on the inner circle, u = 0 on the outer one (row replacement).
Mirrors ``src/solvers/fem2d_neumann_dirichlet.py``."""
import numpy as np
from scipy.special import jv, jvp, yv, yvp

from .base import BaseSolver


class FEM2DNeumannDirichletSolver(BaseSolver):
    _neumann_inner = True
    _dirichlet_outer = True

    def neumann_element_matrices(self, idx):
        """fem2d_neumann_dirichlet.py:27-83 for one element -> N_e[4]."""
        return self._neumann_vectors_device([int(idx)])[0].cpu().numpy()

    def get_analytical_solution(self, x, y) -> complex:
        """fem2d_neumann_dirichlet.py:133-156."""
        r = np.sqrt(x**2 + y**2)
        theta = np.arctan2(y, x)
        k = np.sqrt(self.k_squared)
        R = self.outer_radius
        u = 0 + 0j
        for m in range(-self.n_fourier, self.n_fourier + 1):
            num = jv(m, R * k) * yv(m, k * r) - yv(m, R * k) * jv(m, k * r)
            den = jvp(m, k) * yv(m, R * k) - yvp(m, k) * jv(m, R * k)
            u += 1j**m * np.exp(1j * m * theta) * jvp(m, k) * num / den
        return u
