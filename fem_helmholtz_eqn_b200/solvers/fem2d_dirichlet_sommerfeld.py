"""``"dirichlet_sommerfeld"`` solver: u = 1 on the inner circle (row replacement), ABC on the outer one.
Mirrors ``src/solvers/fem2d_dirichlet_sommerfeld.py`` (note: 4-point Gauss on the Robin edges here)."""
from .base import BaseSolver


class FEM2DDirichletSommerfeldSolver(BaseSolver):
    _robin_outer = True
    _robin_gauss = 4
    _dirichlet_inner = True

    def sommerfeld_element_matrices(self, idx):
        """fem2d_dirichlet_sommerfeld.py:26-78 for one element -> S_e[4,4]."""
        return self._robin_matrices_device([int(idx)], self._robin_gauss)[0].cpu().numpy()
