"""Solver registry and factory -- the reference's dispatch boundary (``src/utils.py:7-22``)."""
from .solvers.fem2d import FEM2DSolver
from .solvers.fem2d_dirichlet import FEM2DDirichletSolver
from .solvers.fem2d_dirichlet_sommerfeld import FEM2DDirichletSommerfeldSolver
from .solvers.fem2d_neumann_dirichlet import FEM2DNeumannDirichletSolver

solvers = ["default", "dirichlet", "dirichlet_sommerfeld", "neumann_dirichlet"]

_REGISTRY = {
    "default": FEM2DSolver,
    "dirichlet": FEM2DDirichletSolver,
    "neumann_dirichlet": FEM2DNeumannDirichletSolver,
    "dirichlet_sommerfeld": FEM2DDirichletSommerfeldSolver,
}


def get_solver(type: str, eqn, **kwargs):
    assert type in solvers, f"Invalid solver type: {type} expected one of: {solvers}"
    return _REGISTRY[type](type, eqn, **kwargs)
