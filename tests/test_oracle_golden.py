"""Pins the oracle (oracle/*.py) to vectors produced by the reference itself
(tests/golden/*.npz, made by oracle/make_golden.py) and to the notebook's stored
stdout (nb:439-444).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, rel_err
from oracle import banded1d, ref1d, ref2d


def test_element_boundary_shape_bit_exact(golden1d):
    g = golden1d
    for h, Ke, Me in zip(g["em_h"], g["em_Ke"], g["em_Me"]):
        ke, me = ref1d.element_matrices(float(h))
        assert np.array_equal(ke, Ke) and np.array_equal(me, Me)
    for (k, th), Ne, Se in zip(g["bm_k_theta"], g["bm_Ne"], g["bm_Se"]):
        ne, se = ref1d.boundary_matrices(0.1, float(k), float(th))
        assert np.array_equal(ne, Ne) and np.array_equal(se, Se)
    for xi, phi, dphi in zip(g["sf_xi"], g["sf_phi"], g["sf_dphi"]):
        p, dp = ref1d.shape_functions(float(xi))
        assert np.array_equal(p, phi) and np.array_equal(dp, dphi)


def test_dense_oracle_matches_reference(golden1d):
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        if N > 1000:
            continue  # dense path is O(N^2); the banded oracle covers the larger cases
        h, nodes, elements = ref1d.create_mesh(N, L)
        assert h == float(g[f"c{idx}_h"]) and np.array_equal(nodes, g[f"c{idx}_nodes"])
        assert np.array_equal(elements[:4], g[f"c{idx}_elements_head"])
        A, b = ref1d.assembly(N, k, h, elements)
        assert np.array_equal(np.diag(A), g[f"c{idx}_d"])
        assert np.array_equal(np.diag(A, 1), g[f"c{idx}_du"][:-1])
        assert np.array_equal(np.diag(A, -1), g[f"c{idx}_dl"][1:])
        assert np.array_equal(b, g[f"c{idx}_b"])
        _, u = ref1d.solve_helmholtz_fem(N, k, L=L)
        assert rel_err(u, g[f"c{idx}_u"]) < 1e-11
        _, u_dd = ref1d.solve_helmholtz_fem(N, k, L=L, dirichlet=(1, 0))
        assert rel_err(u_dd, g[f"c{idx}_u_dd"]) < 1e-10
        _, u_ds = ref1d.solve_helmholtz_fem(N, k, L=L, dirichlet=(1, None))
        assert rel_err(u_ds, g[f"c{idx}_u_ds"]) < 1e-10
        l2 = ref1d.compute_L2_error(k, N, elements, nodes, g[f"c{idx}_u"])
        assert abs(l2 - g[f"c{idx}_l2"]) <= 1e-13 * abs(g[f"c{idx}_l2"])


def test_banded_oracle_bit_exact_bands_and_solution(golden1d):
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        h = L / N
        dl, d, du, b = banded1d.assemble_bands(N, k, h)
        for name, arr in (("dl", dl), ("d", d), ("du", du), ("b", b)):
            assert np.array_equal(arr, g[f"c{idx}_{name}"]), (idx, name)
        u = banded1d.solve_bands(dl, d, du, b)
        # zgtsv vs the reference's dense zgesv: both backward stable; the forward gap is
        # bounded by cond(A)*eps, ~1e-9 at N=4096, k=1.
        tol = 1e-10 if N <= 2049 else 5e-9
        assert rel_err(u, g[f"c{idx}_u"]) < tol, (idx, rel_err(u, g[f"c{idx}_u"]))
        be, _ = banded1d.residual_metrics(dl, d, du, b, g[f"c{idx}_u"])
        assert be < 1e-15
        if N <= 999:
            for key, bc in (("u_dd", dict(bc_left="dirichlet", bc_right="dirichlet")),
                            ("u_ds", dict(bc_left="dirichlet", bc_right="sommerfeld"))):
                ub = banded1d.solve_bands(*banded1d.assemble_bands(N, k, h, **bc))
                assert rel_err(ub, g[f"c{idx}_{key}"]) < 1e-10


def test_nodal_h_mode_equals_scalar_h_on_exact_mesh():
    # on a mesh whose differences are exact (dyadic h) the two modes coincide bit for bit
    N, k, L = 64, 7.0, 1.0
    x = np.linspace(0, L, N + 1)
    a = banded1d.assemble_bands(N, k, L / N)
    b = banded1d.assemble_bands(N, k, L / N, x=x)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))


def test_discrete_closed_form(golden1d):
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        ex = banded1d.discrete_exact(N, k, L)
        tol = 1e-10 if N <= 2049 else 5e-9
        assert rel_err(g[f"c{idx}_u"], ex) < tol
        if N <= 999:
            assert rel_err(g[f"c{idx}_u_ds"], banded1d.discrete_exact(N, k, L, "dirichlet", "sommerfeld")) < 1e-9
            assert rel_err(g[f"c{idx}_u_dd"], banded1d.discrete_exact(N, k, L, "dirichlet", "dirichlet")) < 1e-8


def test_continuum_cross_check(golden1d):
    # nb:219: -exp(ikx) differs from the FEM solution by the dispersion error ~k^3 h^2 L/12
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        bound = 2 * k**3 * (L / N) ** 2 * L / 12
        if bound < 0.5:
            err = np.max(np.abs(g[f"c{idx}_u"] - ref1d.analytical_solution(g[f"c{idx}_nodes"], k)))
            assert err <= bound + 1e-9


def test_convergence_rates_match_notebook_stdout(golden1d):
    stored = json.load(open(os.path.join(GOLDEN, "versions.json")))["notebook_stored_stdout"]
    rates = ref1d.convergence_rates([1, 5, 10, 15, 30, 100], [10, 20, 40, 80, 160])
    lines = [f"Convergence rate for k = {k}: {r:.2f}" for k, r in rates.items()]
    assert "\n".join(lines) + "\n" == stored          # nb:439-444, verbatim
    assert np.allclose(np.array(list(rates.values())), golden1d["conv_rates"], rtol=1e-9)


@pytest.mark.parametrize("mname,n_r,n_t", [("m0", 3, 12), ("m1", 6, 24)])
def test_2d_oracle_matches_reference_classes(golden2d, mname, n_r, n_t):
    g = golden2d
    ri, ro = g[f"{mname}_radii"]
    m = ref2d.structured_annulus(n_r, n_t, ri, ro)
    assert np.array_equal(m.nodes, g[f"{mname}_nodes"]) and np.array_equal(m.elements, g[f"{mname}_elements"])
    assert np.array_equal(ref2d.q4_element_matrix(m.nodes[m.elements[0]], 5.0), g[f"{mname}_Ke0"])
    for order in range(4):
        assert ref2d.abc_coeff(5.0, order, ro) == g[f"{mname}_abc{order}"]
    with pytest.raises(ValueError):
        ref2d.abc_coeff(5.0, 7, ro)
    runs = [("default", 0), ("default", 1), ("default", 3), ("dirichlet", 3),
            ("dirichlet_sommerfeld", 0), ("dirichlet_sommerfeld", 2), ("neumann_dirichlet", 3)]
    for kind, order in runs:
        for k2 in (5.0, 30.25):
            tag = f"{mname}_{kind}_o{order}_k{k2}"
            Kpre, Fpre = ref2d.assemble(kind, m, k2, 1, order)
            r, c = np.nonzero(Kpre)
            assert np.array_equal(r, g[f"{tag}_Kpre_rows"]) and np.array_equal(c, g[f"{tag}_Kpre_cols"])
            assert np.array_equal(Kpre[r, c], g[f"{tag}_Kpre_vals"])
            assert np.array_equal(Fpre, g[f"{tag}_Fpre"])
            ur, ui, K, F = ref2d.solve(kind, m, k2, 1, order, ri, ro)
            r, c = np.nonzero(K)
            assert np.array_equal(r, g[f"{tag}_K_rows"]) and np.array_equal(c, g[f"{tag}_K_cols"])
            assert np.array_equal(K[r, c], g[f"{tag}_K_vals"]) and np.array_equal(F, g[f"{tag}_F"])
            assert rel_err(ur + 1j * ui, g[f"{tag}_u"]) < 1e-10
