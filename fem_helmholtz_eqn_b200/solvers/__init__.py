from .base import BaseSolver, CsrMatrix  # noqa: F401
from .fem2d import FEM2DSolver  # noqa: F401
from .fem2d_dirichlet import FEM2DDirichletSolver  # noqa: F401
from .fem2d_dirichlet_sommerfeld import FEM2DDirichletSommerfeldSolver  # noqa: F401
from .fem2d_neumann_dirichlet import FEM2DNeumannDirichletSolver  # noqa: F401
