"""``"default"`` solver: Neumann load on the inner circle, Sommerfeld / ABC (Robin) condition on the outer one.
Mirrors ``src/solvers/fem2d.py`` (2-point Gauss on the Robin edges, 3-point on the Neumann edges)."""
import numpy as np

from .base import BaseSolver


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
