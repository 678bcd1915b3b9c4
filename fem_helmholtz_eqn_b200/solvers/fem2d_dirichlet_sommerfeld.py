"""``"dirichlet_sommerfeld"`` solver: u = 1 on the inner circle (row replacement), ABC on the outer one.
Mirrors ``src/solvers/fem2d_dirichlet_sommerfeld.py`` (note: 4-point Gauss on the Robin edges here)."""
from .base import BaseSolver
from .utils import get_analytical_solution_sommerfeld


class FEM2DDirichletSommerfeldSolver(BaseSolver):
    _robin_outer = True
    _robin_gauss = 4
    _dirichlet_inner = True

    def sommerfeld_element_matrices(self, idx):
        """fem2d_dirichlet_sommerfeld.py:26-78 for one element -> S_e[4,4]."""
        return self._robin_matrices_device([int(idx)], self._robin_gauss)[0].cpu().numpy()

    def get_analytical_solution(self, x, y) -> complex:
        """fem2d_dirichlet_sommerfeld.py:115-131 -> ``solvers/utils.py:get_analytical_solution_sommerfeld`` (which knows
        the orders 0, 1, 2 and 4: with the default ``abc_order=3`` the reference raises UnboundLocalError, and so does
        this)."""
        return get_analytical_solution_sommerfeld(x, y, self.abc_order, self.k_squared, self.inner_radius,
                                                  self.outer_radius)
