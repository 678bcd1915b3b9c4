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
    with pytest.raises(NotImplementedError):       # mesh generation with gmsh is out of scope
        HelmHoltz(mesher="gmsh")


def test_dirichlet_against_analytic_bessel_solution():
    """fem2d_dirichlet.py:57-85: J0/Y0 closed form; bilinear quads on a 24 x 96 annulus give ~1e-3."""
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=96, n_r=24, mesher="structured")
    s = get_solver("dirichlet", eqn, k_squared=5.0)
    ur, ui = s.solve()
    exact = s.get_analytical_solution(eqn.nodes[:, 0], eqn.nodes[:, 1])
    assert rel_err(ur + 1j * ui, exact) < 2e-3
    assert s.compute_L2_error(ur, ui) < 5e-3


@pytest.mark.parametrize("kind", ["default", "dirichlet", "dirichlet_sommerfeld", "neumann_dirichlet"])
def test_sparse_band_solver_matches_dense_and_numpy(kind):
    """*.solve() runs a banded LU with partial pivoting on the CSR matrix.  Same answer as the dense
    device LU and as np.linalg.solve on the dense view (the reference's statement), to 1e-12."""
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=40, n_r=10, mesher="structured")
    s = get_solver(kind, eqn, k_squared=30.25)
    assert s.linear_solver == "auto"
    ur, ui = s.solve()
    perm, kl, ku = s._band_plan()
    assert 2 * kl + ku + 1 < eqn.n_nodes // 3                         # a narrow band (natural or reordered)
    u_np = np.linalg.solve(np.asarray(s.K), s.F)
    assert rel_err(ur + 1j * ui, u_np) < 1e-12
    d = get_solver(kind, eqn, k_squared=30.25)
    d.linear_solver = "dense"
    dr, di = d.solve()
    assert rel_err(ur + 1j * ui, dr + 1j * di) < 1e-12


def test_sparse_band_solver_on_a_cluster_of_ctas_at_4920_nodes():
    """The banded LU shares its update window over a cluster of CTAs when the window is large (kl (kl + ku) >= 4096):
    a 120 x 40 annulus (4,920 nodes, bandwidth 239) against the residual of the reference's dense K."""
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=120, n_r=40, mesher="structured")
    s = get_solver("dirichlet_sommerfeld", eqn, k_squared=30.25)
    ur, ui = s.solve()
    _, kl, ku = s._band_plan()
    assert kl * (kl + ku) >= 4096
    K, u = np.asarray(s.K), ur + 1j * ui
    r = K @ u - s.F
    assert np.linalg.norm(r) / (np.linalg.norm(K, np.inf) * np.linalg.norm(u) + np.linalg.norm(s.F)) < 1e-15


def test_sparse_band_solver_reorders_a_scrambled_mesh():
    """An unstructured numbering (what gmsh hands out) has a band as wide as the matrix: the solver falls back on a
    reverse Cuthill-McKee ordering of the pattern and still returns the solution in the caller's numbering."""
    base = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=48, n_r=12, mesher="structured")
    rng = np.random.default_rng(7)
    new_of_old = rng.permutation(base.n_nodes)
    nodes = np.empty_like(base.nodes)
    nodes[new_of_old] = base.nodes
    eqn = HelmHoltz.from_arrays(nodes, new_of_old[np.asarray(base.elements)])
    s = get_solver("dirichlet_sommerfeld", eqn, k_squared=30.25)
    ur, ui = s.solve()
    perm, kl, ku = s._band_plan()
    assert perm is not None and 2 * kl + ku + 1 < eqn.n_nodes // 3
    ref = get_solver("dirichlet_sommerfeld", base, k_squared=30.25)
    rr, ri = ref.solve()
    assert rel_err((ur + 1j * ui)[new_of_old], rr + 1j * ri) < 1e-10
    assert rel_err(ur + 1j * ui, np.linalg.solve(np.asarray(s.K), s.F)) < 1e-12


def test_sparse_band_solver_reports_singular_matrices():
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=12, n_r=4, mesher="structured")
    s = get_solver("default", eqn, k_squared=5.0)
    s.assemble()
    s.K.vals.zero_()                                               # an exactly singular matrix
    with pytest.raises(np.linalg.LinAlgError):
        s._solve_device()


def test_robin_matrix_S_is_available_on_demand(golden2d):
    """fem2d.py:141 keeps the boundary matrix S next to K; here it is summed on first access.  K - S must be the
    interior stiffness (no entry of S outside the outer-boundary nodes' rows and columns)."""
    eqn, ri, ro = _eqn(golden2d, "m0")
    s = get_solver("default", eqn, k_squared=5.0, abc_order=1, inner_radius=ri, outer_radius=ro)
    s.assemble()
    S, K = np.asarray(s.S), np.asarray(s.K)
    outer = np.zeros(eqn.n_nodes, bool)
    outer[np.asarray(eqn.outer_boundary_node_indices)] = True
    assert np.abs(S[~outer]).max() == 0.0 and np.abs(S[:, ~outer]).max() == 0.0 and np.abs(S).max() > 0.0
    plain = get_solver("dirichlet", eqn, k_squared=5.0, inner_radius=ri, outer_radius=ro)
    plain.assemble()                                                   # same element matrices, no Robin term
    assert np.allclose(K - S, np.asarray(plain.K), rtol=1e-14, atol=1e-14)
    with pytest.raises(AttributeError):
        plain.S


RUNS_POST = [(kind, order, k2, nf) for kind, order in RUNS for k2, nf in ((5.0, 1), (30.25, 2))]


@pytest.mark.parametrize("mname", ["m0", "m1"])
def test_compute_l2_error_on_device_matches_reference(golden2d, golden2d_post, mname):
    """f2: BaseSolver.compute_L2_error (base.py:119-177) -- Gauss points and the quadrature on the device, the analytic
    solution on the host as in the reference -- against the reference's own value for the reference's own solution,
    for all four classes (1e-12: the sum over elements is a tree here, sequential there)."""
    g, gp = golden2d, golden2d_post
    eqn, ri, ro = _eqn(g, mname)
    for kind, order, k2, nf in RUNS_POST:
        tag = f"{mname}_{kind}_o{order}_k{k2}"
        s = get_solver(kind, eqn, k_squared=k2, abc_order=order, inner_radius=ri, outer_radius=ro, n_fourier=nf)
        u = gp[f"{tag}_u"]
        l2 = s.compute_L2_error(u.real, u.imag)
        assert abs(l2 - gp[f"{tag}_l2"]) <= 1e-12 * gp[f"{tag}_l2"], (tag, l2, gp[f"{tag}_l2"])
        ur, ui = s.solve()                         # ... and of this repo's own solution: the same number to 1e-8
        assert abs(s.compute_L2_error(ur, ui) - gp[f"{tag}_l2"]) <= 1e-8 * gp[f"{tag}_l2"], tag


@pytest.mark.parametrize("mname", ["m0", "m1"])
def test_neumann_load_generated_on_device(golden2d, mname):
    """f4: get_normal_derivative (fem2d.py:12-27) at the edge Gauss points by fh_q4_neumann_load_c128 (CUDA jn()):
    the load vector agrees with the reference's (scipy Bessel functions) to 1e-10; the default host path is bit-exact."""
    g = golden2d
    eqn, ri, ro = _eqn(g, mname)
    for kind in ("default", "neumann_dirichlet"):
        for k2 in (5.0, 30.25):
            tag = f"{mname}_{kind}_o3_k{k2}"
            s = get_solver(kind, eqn, k_squared=k2, abc_order=3, inner_radius=ri, outer_radius=ro)
            s.neumann_load = "device"
            s.assemble()
            want = g[f"{tag}_Fpre"]
            assert np.abs(s.F - want).max() <= 1e-10 * np.abs(want).max(), (tag, np.abs(s.F - want).max())
            ur, ui = s.solve()
            assert rel_err(ur + 1j * ui, g[f"{tag}_u"]) < 1e-9
    s = get_solver("default", eqn, k_squared=5.0, n_fourier=3, inner_radius=ri, outer_radius=ro)
    host = s._neumann_vectors_device(list(eqn.inner_boundary_element_indices)).cpu().numpy()
    s.neumann_load = "device"
    dev = s._neumann_vectors_device(list(eqn.inner_boundary_element_indices)).cpu().numpy()
    assert np.abs(dev - host).max() <= 1e-10 * np.abs(host).max()


def test_boundary_lists_on_device_equal_the_host_lists(golden2d):
    """f3: helmholtz.py:95-118 (boundary nodes by np.isclose on the node radius, elements touching them) by
    fh_q4_boundary_lists + fh_compact_flags_i32: the same ascending lists as numpy gives, on the golden meshes, a fine
    structured mesh and a scrambled numbering."""
    g = golden2d
    for mname in ("m0", "m1"):
        ri, ro = (float(v) for v in g[f"{mname}_radii"])
        eqn = HelmHoltz.from_arrays(g[f"{mname}_nodes"], g[f"{mname}_elements"], ri, ro, topology="device")
        assert eqn.inner_boundary_element_indices == g[f"{mname}_inner_el"].tolist()
        assert eqn.outer_boundary_element_indices == g[f"{mname}_outer_el"].tolist()
    host = HelmHoltz(inner_radius=0.7, outer_radius=2.1, n_theta=300, n_r=41)
    rng = np.random.default_rng(3)
    new_of_old = rng.permutation(host.n_nodes)
    nodes = np.empty_like(host.nodes)
    nodes[new_of_old] = host.nodes
    nodes += rng.uniform(-4e-6, 4e-6, nodes.shape)       # inside and outside the 1e-5 + 1e-5 R window
    elements = new_of_old[np.asarray(host.elements)]
    a = HelmHoltz.from_arrays(nodes, elements, 0.7, 2.1)
    b = HelmHoltz.from_arrays(nodes, elements, 0.7, 2.1, topology="device")
    for name in ("inner_boundary_node_indices", "outer_boundary_node_indices"):
        assert np.array_equal(getattr(a, name), getattr(b, name)) and len(getattr(a, name)) > 0
    assert a.inner_boundary_element_indices == b.inner_boundary_element_indices
    assert a.outer_boundary_element_indices == b.outer_boundary_element_indices


def test_apply_boundary_conditions_acts_on_the_matrix_it_is_given(golden2d):
    """apply_boundary_conditions(A, b) modifies and returns ITS arguments (fem2d_dirichlet.py:11-29), a
    device CSR matrix or the reference's dense type alike."""
    g = golden2d
    eqn, ri, ro = _eqn(g, "m0")
    tag = "m0_dirichlet_o3_k5.0"
    s = get_solver("dirichlet", eqn, k_squared=5.0, inner_radius=ri, outer_radius=ro)
    s.assemble()
    K_dense, F = np.asarray(s.K).copy(), s.F.copy()
    A2, b2 = s.apply_boundary_conditions(K_dense, F)
    assert A2 is K_dense and b2 is F
    r, c = np.nonzero(A2)
    assert np.array_equal(r, g[f"{tag}_K_rows"]) and np.array_equal(c, g[f"{tag}_K_cols"])
    assert np.array_equal(A2[r, c], g[f"{tag}_K_vals"]) and np.array_equal(b2, g[f"{tag}_F"])
    assert np.array_equal(_nonzeros(s.K)[2], g[f"{tag}_Kpre_vals"])         # the solver's own K is untouched
    K3, F3 = s.apply_boundary_conditions()                                     # default: the solver's own pair
    assert K3 is s.K and np.array_equal(F3, g[f"{tag}_F"]) and np.array_equal(_nonzeros(s.K)[2], g[f"{tag}_K_vals"])
