"""``"default"`` solver: Neumann load on the inner circle, Sommerfeld / ABC (Robin) condition on the outer one.
Mirrors ``src/solvers/fem2d.py`` (2-point Gauss on the Robin edges, 3-point on the Neumann edges)."""
from .base import BaseSolver
from .utils import get_solution


class FEM2DSolver(BaseSolver):
    _neumann_inner = True
    _robin_outer = True
    _robin_gauss = 2

    def sommerfeld_element_matrices(self, idx):
        """fem2d.py:29-80 for one element -> S_e[4,4]."""
        return self._robin_matrices_device([int(idx)], self._robin_gauss)[0].cpu().numpy()

    def neumann_element_matrices(self, idx):
        """fem2d.py:82-137 for one element -> N_e[4]."""
        return self._neumann_vectors_device([int(idx)])[0].cpu().numpy()

    def get_analytical_solution(self, x, y) -> complex:
        """fem2d.py:183-224: the closed form of ``solvers/utils.py:get_solution`` for this class's boundary conditions
        and ``abc_order`` (the Hankel-function sum the reference also evaluates there is discarded by it, so it is not
        computed here)."""
        return get_solution(x, y, self.abc_order, self.k_squared, self.inner_radius, self.outer_radius, self.n_fourier)
