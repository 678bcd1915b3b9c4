"""fem_helmholtz_eqn_b200 -- B200-native (sm_100a) drop-in for the finite-element Helmholtz
assembly + solve path of anishasahu/fem-helmholtz-eqn.

1D (the notebook's function names):   ``from fem_helmholtz_eqn_b200.fem1d import *``
2D (``get_solver`` and the solver classes): ``from fem_helmholtz_eqn_b200.utils import get_solver``

All numerical work happens in hand-written CUDA kernels reached through the C ABI declared in
``include/fem_helmholtz.h``; importing the package does not need a GPU, calling it does.
"""
from . import _lib  # noqa: F401
from .fem1d import (TridiagonalSystem, analytical_solution, assemble_and_solve, assemble_bands, assembly,  # noqa: F401
                    boundary_matrices, compute_L2_error, convergence_test, create_mesh, element_matrices, gauss_point, gauss_weight,
                    shape_functions, solve_helmholtz_fem, solve_helmholtz_sweep, solve_tridiagonal,
                    tridiagonal_residual)

__version__ = "0.1.0"
