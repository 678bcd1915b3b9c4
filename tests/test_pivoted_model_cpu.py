"""CPU check of the index bookkeeping of the two-sided pivoted solve (cta_pivoted_solve in csrc/fh_residual.cu) on its
numpy model: against np.linalg.solve (the reference's solver, nb:204) on matrices that NEED pivoting."""
import numpy as np
import pytest

from models import pivoted_babe_model as M
from oracle import banded1d


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 100, 257, 513])
def test_two_sided_pivoted_model_matches_dense_solve(n):
    rng = np.random.default_rng(n)
    dl, d, du, b = (rng.normal(size=n) + 1j * rng.normal(size=n) for _ in range(4))
    if n > 3:
        d[::3] = 0  # exact zeros on the diagonal: unpivoted elimination breaks down at the first one
    A = np.diag(d) + np.diag(dl[1:], -1) + np.diag(du[:-1], 1)
    ref = np.linalg.solve(A, b)
    x = M.solve(dl, d, du, b)
    assert np.linalg.norm(x - ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("N,k", [(1, 2.0), (2, 4.0), (64, 128.0), (1000, 2000.0), (4096, 806.8412377190356)])
def test_two_sided_pivoted_model_on_helmholtz_rows(N, k):
    """k*h = 2 (first pivot exactly zero) and the resonance of the batched kernel's partition seen in the 8-rank sweep."""
    bands = banded1d.assemble_bands(N, k, 1.0 / N)
    x = M.solve(*bands)
    assert banded1d.residual_metrics(*bands, x)[0] < 1e-15


def test_two_sided_pivoted_model_singular_system_gives_nonfinite():
    n = 9
    dl, d, du, b = np.ones(n, complex), np.zeros(n, complex), np.ones(n, complex), np.ones(n, complex)
    # odd-sized "hollow" tridiagonal matrix: structurally singular
    x = M.solve(dl, d, du, b)
    assert not np.all(np.isfinite(x))
