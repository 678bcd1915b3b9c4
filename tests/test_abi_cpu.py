"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/fem_helmholtz.h declares, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

fh = pytest.importorskip("fem_helmholtz_eqn_b200")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fem_helmholtz.h")).read()
    return sorted(set(re.findall(r"FH_API[^;(]*?\b(fh_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == fh._lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(fh._lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert fh._lib.load().fh_abi_version() == 1
    assert fh._lib.load().fh_tridiag_batched_max_n() == 4112


def test_invalid_arguments_are_reported_without_gpu():
    lib = fh._lib.load()
    rc = lib.fh_assemble_1d_c128(None, None, None, 1, None, None, None, None, None)
    assert rc == -1 and b"NULL" in lib.fh_last_error()
    p = fh._lib.Problem1d(n_elem=0, h=0.1, h_mode=0, bc_left=0, bc_right=1)
    rc = lib.fh_assemble_1d_c128(ctypes.byref(p), None, None, 1, None, None, None, None, None)
    assert rc == -1 and b"n_elem" in lib.fh_last_error()


def test_round2_entry_points_validate_their_arguments_without_gpu():
    """Host-side halves of the entry points added in round 2: workspace queries and argument checks (no kernel runs)."""
    lib = fh._lib.load()
    n, batch = 4097, 1000
    # workspaces: 64 B per row per re-solve in flight (+ the per-batch bookkeeping); queries are monotone
    piv1, piv8 = lib.fh_tridiag_solve_pivoted_workspace_bytes(n, 1), lib.fh_tridiag_solve_pivoted_workspace_bytes(n, 8)
    assert piv1 == 64 * n and piv8 == 8 * piv1
    assert lib.fh_tridiag_solve_pivoted_workspace_bytes(n, 10**6) == 256 * piv1            # capped concurrency
    rep = lib.fh_tridiag_repair_flagged_workspace_bytes(n, batch, 4)
    assert rep > 4 * piv1 and rep > lib.fh_tridiag_repair_flagged_workspace_bytes(n, batch, 2)
    assert lib.fh_helmholtz1d_residual_workspace_bytes() > 0 and lib.fh_spike_peer_buffer_bytes() >= 2 * 16 * 128
    assert lib.fh_q4_l2_error_workspace_bytes() > 0
    # the geometry table of the non-uniform assembly: 48 B per mesh row, only where it pays
    assert lib.fh_assemble_1d_workspace_bytes(4096, 1, 8) == 48 * 4097
    assert lib.fh_assemble_1d_workspace_bytes(4096, 1, 7) == 0 and lib.fh_assemble_1d_workspace_bytes(4096, 0, 100) == 0
    p = fh._lib.Problem1d(n_elem=100, h=0.01, h_mode=0, bc_left=0, bc_right=1)
    rc = lib.fh_assemble_1d_slab_c128(ctypes.byref(p), 1.0, 1.0, None, 50, 60, None, None, None, None, None)
    assert rc == -1 and b"outside the mesh" in lib.fh_last_error()
    rc = lib.fh_helmholtz1d_residual_c128(ctypes.byref(p), 1.0, 1.0, None, 0, 101, None, None, None, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.fh_last_error()
    rc = lib.fh_tridiag_repair_flagged_c128(None, None, None, None, None, 10, 10, None, 1e-13, None, 0, None)
    assert rc == -1 and b"NULL" in lib.fh_last_error()
    rc = lib.fh_helmholtz1d_spike_solve_peer_kdev_c128(ctypes.byref(p), None, 0, 101, None, 0, None, 0, 2, None, 2.0, None,
                                                       None, None, None)
    assert rc == -1 and b"NULL" in lib.fh_last_error()
    rc = lib.fh_compact_flags_i32(None, 10, None, None, None)
    assert rc == -1


def test_pipeline_choice_is_host_logic():
    from fem_helmholtz_eqn_b200 import fem1d
    assert fem1d._pick_pipeline(None, False, 4096, False) == "one_pass"      # uniform mesh within the batched size
    assert fem1d._pick_pipeline(None, False, 4112, False) == "one_pass"      # 4113 nodes: the long-mesh one-pass solver
    assert fem1d._pick_pipeline(None, False, 1, False) == "staged"           # a single element: the two-kernel path
    assert fem1d._pick_pipeline(None, False, 4096, True) == "staged"         # non-uniform mesh
    assert fem1d._pick_pipeline(None, True, 4096, False) == "fused"          # the legacy switch
    assert fem1d._pick_pipeline("staged", True, 4096, False) == "staged"     # an explicit choice wins
    with pytest.raises(ValueError):
        fem1d._pick_pipeline("fastest", False, 4096, False)


def test_wavenumber_squares_follow_the_reference_scalar_power():
    """nb:185 evaluates ``k**2`` on a scalar: libm's pow, which is not always the correctly rounded square.  The sweep
    entry points take arrays of wavenumbers and must square them the same way (one ulp of k**2 is one ulp of every
    assembled diagonal): k = 1252.1780850687405 is a wavenumber where the two differ."""
    import torch
    from fem_helmholtz_eqn_b200 import fem1d
    k = 1252.1780850687405
    assert (k ** 2).hex() == "0x1.7eccdf4ec05c5p+20" and (k * k).hex() == "0x1.7eccdf4ec05c4p+20"
    rng = np.random.default_rng(5)
    ks = np.concatenate([[k], rng.uniform(0.1, 2000.0, 20000), np.linspace(1.0, 1000.0, 65536)])
    want = np.array([float(v) ** 2 for v in ks])
    assert np.array_equal(fem1d.squares_like_the_reference(ks), want)
    assert int((want != ks * ks).sum()) > 0                   # (the distinction is real on this platform)
    for container in (ks[:50], list(ks[:50]), torch.from_numpy(ks[:50].copy())):
        assert np.array_equal(fem1d._host_k_and_k2(container)[1], want[:50])
    assert fem1d._host_k_and_k2(k)[1][0] == k ** 2 and fem1d._host_k_and_k2(np.float64(k))[1][0] == k ** 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fh.solve_helmholtz_fem(10, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fh.element_matrices(0.1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fem_helmholtz_eqn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f


def test_host_mesh_and_k2_logic():
    h, nodes, elements = fh.create_mesh(100, 2.0)
    assert h == 2.0 / 100 and nodes.shape == (101,) and elements.shape == (100, 2)
    assert elements.dtype == np.int64 and elements[-1].tolist() == [99, 100]
    from fem_helmholtz_eqn_b200.fem1d import _host_k_and_k2
    ks, k2s, scalar = _host_k_and_k2(5)
    assert scalar and ks[0] == 5.0 and k2s[0] == 25.0
    ks, k2s, scalar = _host_k_and_k2([1.5, 2])
    assert not scalar and k2s.tolist() == [1.5**2, 4.0]
    ks, k2s, scalar = _host_k_and_k2(np.linspace(1, 1000, 7))
    assert np.array_equal(k2s, np.linspace(1, 1000, 7) ** 2)


def test_kernel_algorithm_model_matches_lapack():
    """The numpy model of the CUDA kernel's algorithm (tests/models) agrees with zgtsv on the edge sizes
    of the slot map; keeps the index maps / cluster coupling honest on machines without a GPU."""
    from oracle import banded1d
    from models.partition_pcr_model import solve
    rng = np.random.default_rng(1)
    for n in (1, 2, 9, 17, 23, 2048, 2050, 2056, 2057, 2058, 3001, 4096, 4097, 4111, 4112):
        dl = rng.normal(size=n) + 1j * rng.normal(size=n)
        du = rng.normal(size=n) + 1j * rng.normal(size=n)
        d = 4 + rng.normal(size=n) + 1j * rng.normal(size=n)
        b = rng.normal(size=n) + 1j * rng.normal(size=n)
        dl[0] = 0
        du[-1] = 0
        u = solve(dl, d, du, b)
        ref = banded1d.solve_bands(dl, d, du, b)
        assert np.linalg.norm(u - ref) / np.linalg.norm(ref) < 1e-13, n
    for N, k in ((999, 10.0), (4096, 1000.0)):
        bands = banded1d.assemble_bands(N, k, 1.0 / N)
        be, _ = banded1d.residual_metrics(*bands, solve(*bands))
        assert be < 1e-14


def test_2d_host_objects_match_reference(golden2d):
    """Mesh stand-in, boundary lists and ABC coefficients are host logic: check them against the golden file."""
    from fem_helmholtz_eqn_b200.helmholtz import HelmHoltz, structured_annulus
    g = golden2d
    for mname, (n_r, n_t) in (("m0", (3, 12)), ("m1", (6, 24))):
        ri, ro = g[f"{mname}_radii"]
        nodes, elements = structured_annulus(n_r, n_t, ri, ro)
        assert np.array_equal(nodes, g[f"{mname}_nodes"]) and np.array_equal(elements, g[f"{mname}_elements"])
        eqn = HelmHoltz(inner_radius=float(ri), outer_radius=float(ro), n_theta=n_t, n_r=n_r, mesher="structured")
        assert list(eqn.inner_boundary_element_indices) == g[f"{mname}_inner_el"].tolist()
        assert list(eqn.outer_boundary_element_indices) == g[f"{mname}_outer_el"].tolist()
        for order in range(4):
            assert eqn.get_abc_coeff(5.0, order) == g[f"{mname}_abc{order}"]
        with pytest.raises(ValueError):
            eqn.get_abc_coeff(5.0, 4)
    from fem_helmholtz_eqn_b200.utils import get_solver, solvers
    assert list(g["solver_names"]) == solvers
    with pytest.raises(AssertionError):
        get_solver("nope", eqn)


def test_analytic_annulus_solutions_match_reference(golden2d, golden2d_post):
    """solvers/utils.py (host scipy, like the reference's src/solvers/utils.py) and the four classes'
    get_analytical_solution against the values the reference's own classes return at the mesh nodes -- bit for bit
    when called point by point, to rounding when vectorised; the reference's order quirk (UnboundLocalError for the
    Sommerfeld variant at abc_order=3, its default) is kept."""
    from fem_helmholtz_eqn_b200.helmholtz import HelmHoltz
    from fem_helmholtz_eqn_b200.utils import get_solver
    g, gp = golden2d, golden2d_post
    runs = [("default", 0), ("default", 1), ("default", 3), ("dirichlet", 3), ("dirichlet_sommerfeld", 0),
            ("dirichlet_sommerfeld", 2), ("neumann_dirichlet", 3)]
    for mname in ("m0", "m1"):
        ri, ro = (float(v) for v in g[f"{mname}_radii"])
        eqn = HelmHoltz.from_arrays(g[f"{mname}_nodes"], g[f"{mname}_elements"], ri, ro)
        for kind, order in runs:
            for k2, nf in ((5.0, 1), (30.25, 2)):
                s = get_solver(kind, eqn, k_squared=k2, abc_order=order, inner_radius=ri, outer_radius=ro, n_fourier=nf)
                want = gp[f"{mname}_{kind}_o{order}_k{k2}_exact_nodes"]
                got = np.array([s.get_analytical_solution(x, y) for x, y in eqn.nodes])
                assert np.array_equal(got, want), (mname, kind, order, k2)
                vec = s.get_analytical_solution(eqn.nodes[:, 0], eqn.nodes[:, 1])
                assert np.allclose(vec, want, rtol=1e-13, atol=1e-15)
    s = get_solver("dirichlet_sommerfeld", eqn)          # abc_order=3: the reference's unhandled branch
    with pytest.raises(UnboundLocalError):
        s.get_analytical_solution(1.2, 0.1)
