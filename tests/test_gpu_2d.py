"""GPU parity tests of the 2D annulus path (get_solver(...).assemble()/.solve()) against K / F / u produced by
the reference's own classes (tests/golden/ref2d.npz).  Assembled values: bit-exact.  Solutions: 1e-10."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

fh = pytest.importorskip("fem_helmholtz_eqn_b200")
from fem_helmholtz_eqn_b200.helmholtz import HelmHoltz  # noqa: E402
from fem_helmholtz_eqn_b200.utils import get_solver, solvers  # noqa: E402

RUNS = [("default", 0), ("default", 1), ("default", 3), ("dirichlet", 3), ("dirichlet_sommerfeld", 0),
        ("dirichlet_sommerfeld", 2), ("neumann_dirichlet", 3)]


def _eqn(g, mname):
    ri, ro = g[f"{mname}_radii"]
    return HelmHoltz.from_arrays(g[f"{mname}_nodes"], g[f"{mname}_elements"], float(ri), float(ro)), float(ri), float(ro)


def _nonzeros(K):
    r, c, v = K.coo()
    keep = v != 0          # np.nonzero(dense) of the reference drops structural zeros (e.g. replaced rows)
    order = np.lexsort((c[keep], r[keep]))
    return r[keep][order], c[keep][order], v[keep][order]


@pytest.mark.parametrize("mname", ["m0", "m1"])
def test_element_and_edge_routines_bit_exact(golden2d, mname):
    g = golden2d
    eqn, ri, ro = _eqn(g, mname)
    s = get_solver("default", eqn, k_squared=5.0, abc_order=3, inner_radius=ri, outer_radius=ro)
    Ke, Fe = s.get_element_matrices(0)
    assert np.array_equal(Ke, g[f"{mname}_Ke0"]) and not Fe.any()
    assert np.array_equal(s.sommerfeld_element_matrices(eqn.outer_boundary_element_indices[0]), g[f"{mname}_Se_outer0"])
    assert np.array_equal(s.neumann_element_matrices(eqn.inner_boundary_element_indices[0]), g[f"{mname}_Ne_inner0"])
    assert list(eqn.inner_boundary_element_indices) == g[f"{mname}_inner_el"].tolist()
    assert list(eqn.outer_boundary_element_indices) == g[f"{mname}_outer_el"].tolist()


@pytest.mark.parametrize("mname", ["m0", "m1"])
@pytest.mark.parametrize("kind,order", RUNS)
def test_assemble_and_solve_match_reference(golden2d, mname, kind, order):
    g = golden2d
    eqn, ri, ro = _eqn(g, mname)
    for k2 in (5.0, 30.25):
        tag = f"{mname}_{kind}_o{order}_k{k2}"
        s = get_solver(kind, eqn, k_squared=k2, abc_order=order, inner_radius=ri, outer_radius=ro)
        s.assemble()
        r, c, v = _nonzeros(s.K)
        assert np.array_equal(r, g[f"{tag}_Kpre_rows"]) and np.array_equal(c, g[f"{tag}_Kpre_cols"]), tag
        assert np.array_equal(v, g[f"{tag}_Kpre_vals"]), tag
        assert np.array_equal(s.F, g[f"{tag}_Fpre"]), tag
        ur, ui = s.solve()
        r, c, v = _nonzeros(s.K)
        assert np.array_equal(r, g[f"{tag}_K_rows"]) and np.array_equal(c, g[f"{tag}_K_cols"]), tag
        assert np.array_equal(v, g[f"{tag}_K_vals"]) and np.array_equal(s.F, g[f"{tag}_F"]), tag
        assert ur.dtype == np.float64 and ui.dtype == np.float64
        assert rel_err(ur + 1j * ui, g[f"{tag}_u"]) < 1e-10, (tag, rel_err(ur + 1j * ui, g[f"{tag}_u"]))
        # the dense view is the reference's K: its own solve statement works on it
        u2 = np.linalg.solve(np.asarray(s.K), s.F)
        assert rel_err(u2, g[f"{tag}_u"]) < 1e-10


def test_factory_and_errors(golden2d):
    eqn, ri, ro = _eqn(golden2d, "m0")
    assert solvers == ["default", "dirichlet", "dirichlet_sommerfeld", "neumann_dirichlet"]
    with pytest.raises(AssertionError):
        get_solver("robin", eqn)
    with pytest.raises(ValueError):
        get_solver("default", eqn, abc_order=7).assemble()
    with pytest.raises(NotImplementedError):
        get_solver("default", eqn).get_analytical_solution(1.0, 0.2)


def test_dirichlet_against_analytic_bessel_solution():
    """fem2d_dirichlet.py:57-85: J0/Y0 closed form; bilinear quads on a 24 x 96 annulus give ~1e-3."""
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=96, n_r=24, mesher="structured")
    s = get_solver("dirichlet", eqn, k_squared=5.0)
    ur, ui = s.solve()
    exact = s.get_analytical_solution(eqn.nodes[:, 0], eqn.nodes[:, 1])
    assert rel_err(ur + 1j * ui, exact) < 2e-3
    assert s.compute_L2_error(ur, ui) < 5e-3
