"""GPU parity tests of the 1D path: every call goes through the C ABI (ctypes) into the CUDA
kernels and is compared with the CPU oracle / the reference's golden vectors.

Tolerances (fp64 / complex128):
  * assembled values (bands, load, element matrices): bit-exact (==);
  * solutions: forward error <= 1e-10 where cond(A)*eps allows (N <= 2049), else the bound
    documented in SURVEY.md 0.4 (5e-9 at N = 4096, k = 1); normwise backward error <= 1e-13 always
    (the north-star bar is 1e-10).
"""
import numpy as np
import pytest
import torch

from conftest import golden_cases, rel_err
from oracle import banded1d

pytestmark = pytest.mark.gpu

fh = pytest.importorskip("fem_helmholtz_eqn_b200")


def _np(t):
    return t.detach().cpu().numpy()


def test_library_loaded_and_device_is_blackwell():
    info = fh._lib.device_info()
    assert info["cc"][0] == 10, info
    assert info["sm_count"] >= 100 and info["smem_optin"] >= 200 * 1024


def test_element_routines_bit_exact(golden1d):
    g = golden1d
    for h, Ke, Me in zip(g["em_h"], g["em_Ke"], g["em_Me"]):
        ke, me = fh.element_matrices(float(h))
        assert np.array_equal(ke, Ke) and np.array_equal(me, Me)
    for (k, th), Ne, Se in zip(g["bm_k_theta"], g["bm_Ne"], g["bm_Se"]):
        ne, se = fh.boundary_matrices(0.1, float(k), float(th))
        assert np.array_equal(ne, Ne) and np.array_equal(se, Se)
    for xi, phi, dphi in zip(g["sf_xi"], g["sf_phi"], g["sf_dphi"]):
        p, dp = fh.shape_functions(float(xi))
        assert np.array_equal(p, phi) and np.array_equal(dp, dphi)


def test_create_mesh_matches_reference(golden1d):
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        h, nodes, elements = fh.create_mesh(N, L)
        assert h == float(g[f"c{idx}_h"]) and np.array_equal(nodes, g[f"c{idx}_nodes"])
        assert np.array_equal(elements[:4], g[f"c{idx}_elements_head"]) and elements.shape == (N, 2)


def test_assembly_bit_exact_vs_reference(golden1d):
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        h, nodes, elements = fh.create_mesh(N, L)
        A, b = fh.assembly(N, k, h, elements)
        dl, d, du = A.bands()
        assert np.array_equal(dl, g[f"c{idx}_dl"]), idx
        assert np.array_equal(d, g[f"c{idx}_d"]), idx
        assert np.array_equal(du, g[f"c{idx}_du"]), idx
        assert np.array_equal(b, g[f"c{idx}_b"]), idx
        assert A.shape == (N + 1, N + 1)


def test_assembly_dense_view_is_drop_in(golden1d):
    g = golden1d
    idx, N, k, L = golden_cases(g)[0]  # notebook demo: N=100, k=5, L=2
    h, nodes, elements = fh.create_mesh(N, L)
    A, b = fh.assembly(N, k, h, elements)
    u = np.linalg.solve(A, b)  # the reference's own statement (nb:204) on our A, b
    assert rel_err(u, g[f"c{idx}_u"]) < 1e-12
    dense = np.asarray(A)
    assert dense.shape == (N + 1, N + 1) and np.count_nonzero(dense) == 3 * N + 1


def test_assembly_batched_and_bc_variants_bit_exact():
    N, L = 257, 1.0
    h = L / N
    ks = [1.0, 7.5, 40.0, 333.25]
    for bc in (dict(), dict(bc_left="dirichlet", bc_right="dirichlet", g_left=1.0, g_right=0.0),
               dict(bc_left="dirichlet", bc_right="sommerfeld", g_left=2.0 - 1.0j),
               dict(bc_left="neumann", bc_right="dirichlet", g_right=0.5j)):
        for theta in (0.0, 0.4):
            dl, d, du, b = (_np(t) for t in fh.assemble_bands(N, ks, h, theta, **bc))
            for i, k in enumerate(ks):
                ref = banded1d.assemble_bands(N, k, h, theta, **bc)
                for got, want in zip((dl[i], d[i], du[i], b[i]), ref):
                    assert np.array_equal(got, want), (bc, theta, k)


def test_assembly_nodal_h_mode_bit_exact():
    rng = np.random.default_rng(0)
    N = 1000
    x = np.linspace(0.0, 1.0, N + 1)
    x[1:-1] += rng.uniform(-0.25, 0.25, N - 1) / N
    for k in (3.0, 250.0):
        got = [_np(t)[0] for t in fh.assemble_bands(N, k, 1.0 / N, nodes=x)]
        want = banded1d.assemble_bands(N, k, 1.0 / N, x=x)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_assembly_nodal_h_many_wavenumbers_bit_exact():
    """n_k >= 8 takes the geometry-table kernel (slabs cut across system boundaries); same bits as the oracle."""
    rng = np.random.default_rng(1)
    for N, n_k in ((1000, 8), (37, 300), (4096, 11)):
        x = np.linspace(0.0, 1.0, N + 1)
        x[1:-1] += rng.uniform(-0.25, 0.25, N - 1) / N
        ks = np.linspace(1.0, 400.0, n_k)
        for bc in (dict(), dict(bc_left="dirichlet", bc_right="dirichlet", g_left=1.0, g_right=0.5j)):
            got = [_np(t) for t in fh.assemble_bands(N, ks, 1.0 / N, 0.3, nodes=x, **bc)]
            for i in (0, 1, n_k // 2, n_k - 1):
                want = banded1d.assemble_bands(N, float(ks[i]), 1.0 / N, 0.3, x=x, **bc)
                assert all(np.array_equal(a[i], b) for a, b in zip(got, want)), (N, n_k, i, bc)


def test_solve_matches_reference_solution(golden1d):
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        nodes, u = fh.solve_helmholtz_fem(N, k, L=L)
        assert np.array_equal(nodes, g[f"c{idx}_nodes"]) and u.dtype == np.complex128
        tol = 1e-10 if N <= 2049 else 5e-9
        assert rel_err(u, g[f"c{idx}_u"]) < tol, (idx, rel_err(u, g[f"c{idx}_u"]))
        bands = banded1d.assemble_bands(N, k, L / N)
        be, rr = banded1d.residual_metrics(*bands, u)
        assert be < 1e-13, (idx, be, rr)  # LAPACK itself: be ~2e-17, plain ||r||/||b|| up to 2e-11 (k=1)
        assert rel_err(u, banded1d.discrete_exact(N, k, L)) < tol
        if N <= 999:  # BASELINE config 1 flavours: Dirichlet rows
            _, udd = fh.solve_helmholtz_fem(N, k, L=L, bc_left="dirichlet", bc_right="dirichlet")
            assert rel_err(udd, g[f"c{idx}_u_dd"]) < 1e-10
            _, uds = fh.solve_helmholtz_fem(N, k, L=L, bc_left="dirichlet")
            assert rel_err(uds, g[f"c{idx}_u_ds"]) < 1e-10


def test_fused_equals_staged(golden1d):
    for N in (8, 999, 2055, 3000, 4096):
        ks = np.array([1.0, 10.0, 123.456, 1000.0])
        _, Us = fh.solve_helmholtz_sweep(N, ks)
        _, Uf = fh.solve_helmholtz_sweep(N, ks, fused=True)
        assert torch.equal(Us, Uf), N  # same coefficients, same arithmetic -> identical bits
        _, Un = fh.solve_helmholtz_sweep(N, ks, fused=True, nodal_h=True)
        assert rel_err(_np(Un), _np(Us)) < 1e-9


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 17, 23, 63, 1000, 2047, 2048, 2049, 2055, 2056, 2057, 3001, 4097, 4112])
def test_random_systems_all_kernel_shapes(n):
    """Edge sizes of the kernel's slot map: < 1 chunk, chunk-aligned, head rows (2049..2056), the
    1-CTA/2-CTA switch (2056|2057), the flagship 4097 and the cluster maximum 4112."""
    rng = np.random.default_rng(n)
    batch = 5 if n < 3000 else 3
    sh = (batch, n)
    dl = rng.normal(size=sh) + 1j * rng.normal(size=sh)
    du = rng.normal(size=sh) + 1j * rng.normal(size=sh)
    d = 4.0 + rng.normal(size=sh) + 1j * rng.normal(size=sh)
    b = rng.normal(size=sh) + 1j * rng.normal(size=sh)
    u, status = fh.solve_tridiagonal(dl, d, du, b)
    assert int(status.abs().sum()) == 0
    u = _np(u)
    for i in range(batch):
        lo, hi = dl[i].copy(), du[i].copy()
        lo[0] = 0.0
        hi[-1] = 0.0  # the ABI ignores dl[0], du[n-1]
        ref = banded1d.solve_bands(lo, d[i], hi, b[i])
        assert rel_err(u[i], ref) < 1e-13, (n, i)
    res = fh.tridiagonal_residual(dl * np.r_[0.0, np.ones(n - 1)], d, du * np.r_[np.ones(n - 1), 0.0], b, u)
    assert res["backward"].max() < 1e-15


@pytest.mark.parametrize("n,batch", [(1, 5), (2, 10), (5, 30), (8, 30), (9, 40), (1000, 37), (2056, 11), (2057, 21), (3001, 12), (4096, 9),
                                     (4097, 9), (4112, 7), (300, 600)])
def test_device_side_refinement_of_flagged_subset(n, batch):
    """fh_tridiag_refine_flagged_c128: exactly the systems flagged FH_STATUS_SMALL_PIVOT get one refinement step (their
    matrices are read in place through a device-built index list); the others are not touched (bit-for-bit).  Covers
    systems of head rows only (n < 8), the single-CTA kernel, even / odd halves (cut rows), head rows (2056, 4112) and
    a batch in which every system is flagged."""
    rng = np.random.default_rng(1000 * n + batch)
    sh = (batch, n)
    dl = rng.normal(size=sh) + 1j * rng.normal(size=sh)
    du = rng.normal(size=sh) + 1j * rng.normal(size=sh)
    d = 4.0 + rng.normal(size=sh) + 1j * rng.normal(size=sh)
    b = rng.normal(size=sh) + 1j * rng.normal(size=sh)
    dl[:, 0] = 0.0
    du[:, -1] = 0.0
    tens = [torch.as_tensor(a, device="cuda") for a in (dl, d, du, b)]
    u, status = fh.solve_tridiagonal(*tens, refine=False)
    assert int(status.abs().sum()) == 0
    flagged = np.ones(batch, bool) if batch == 600 else rng.random(batch) < 0.4
    flagged[-1] = True
    noise = 1e-6 * (rng.normal(size=sh) + 1j * rng.normal(size=sh))
    u_bad = u.clone() + torch.as_tensor(noise * flagged[:, None], device="cuda")   # spoil the flagged solutions
    u_ref = u_bad.clone()
    status = torch.as_tensor(np.where(flagged, 2, 0).astype(np.int32), device="cuda")
    m = fh.fem1d.refine_flagged(*tens, u_bad, status)
    assert m == int(flagged.sum())
    assert int(status.abs().sum()) == 0                       # flags cleared, no new ones
    keep = torch.as_tensor(~flagged, device="cuda")
    assert torch.equal(u_bad[keep], u_ref[keep])              # unflagged systems untouched
    res = fh.tridiagonal_residual(*tens, u_bad)["backward"]
    assert res.max() < 1e-15, res.max()                       # one step repairs a 1e-6 perturbation
    assert rel_err(_np(u_bad), _np(u)) < 1e-12
    # the one-pass refinement kernel (residual formed while the rows are read, correction added by a TMA reduce-add)
    # gives the bits of the step spelled out through the C ABI: r = b - A u, A e = r, u + e
    from fem_helmholtz_eqn_b200 import _lib
    from fem_helmholtz_eqn_b200._lib import ptr, stream_ptr
    lib = _lib.require_cuda()
    r = torch.empty_like(u_ref)
    assert lib.fh_tridiag_residual_vector_c128(*(ptr(t) for t in tens), ptr(u_ref), n, batch, ptr(r), stream_ptr()) == 0
    e, st_e = fh.solve_tridiagonal(tens[0], tens[1], tens[2], r, refine=False)
    assert int((st_e & 1).sum()) == 0
    fl = torch.as_tensor(flagged, device="cuda")
    assert torch.equal(u_bad[fl], (u_ref + e)[fl])


def test_solver_is_deterministic_run_to_run():
    """The kernels use no floating-point atomics and every shared-memory hand-over is fenced by a barrier: repeated
    solves of the same 8,192-wavenumber sweep must agree bit for bit (a missing barrier shows up here as a
    difference between runs), staged and fused alike."""
    N = 4096
    ks = np.linspace(1.0, 1000.0, 8192)
    dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
    first, st0 = fh.solve_tridiagonal(dl, d, du, b, refine=False)
    fused0 = fh.solve_helmholtz_sweep(N, ks, fused=True, check_singular=False)[1].clone()
    for _ in range(5):
        again, st = fh.solve_tridiagonal(dl, d, du, b, refine=False)
        assert torch.equal(again, first) and torch.equal(st, st0)
        assert torch.equal(fh.solve_helmholtz_sweep(N, ks, fused=True, check_singular=False)[1], fused0)


def test_singular_system_raises_linalgerror():
    n = 64
    z = np.zeros((1, n), dtype=complex)
    A = fh.TridiagonalSystem(*(fh._lib.as_device(t, torch.complex128) for t in (z, z, z)), rhs=None)
    with pytest.raises(np.linalg.LinAlgError):
        A.solve(np.ones(n, dtype=complex))
    u, status = fh.solve_tridiagonal(z, z, z, z + 1.0)
    assert int(status[0]) != 0


def test_bad_arguments():
    with pytest.raises(ValueError):
        fh.assembly(10, 1.0, 0.1, np.zeros((9, 2), dtype=int))
    with pytest.raises(ValueError):
        fh.solve_helmholtz_fem(10, 1.0, bc_left="robin")
    with pytest.raises(fh._lib.FHError):
        fh.assemble_bands(0, 1.0, 0.1)


def test_kh_sweep_high_frequency_robin():
    """BASELINE config 5: k = 1000, Robin/impedance row, >= 10 points per wavelength."""
    for N in (1600, 4096):
        nodes, u = fh.solve_helmholtz_fem(N, 1000.0)
        ex = banded1d.discrete_exact(N, 1000.0)
        assert rel_err(u, ex) < 1e-10, (N, rel_err(u, ex))


def test_full_size_sweep_properties():
    """BASELINE config 3 at full size: 65,536 wavenumbers x 4,096 elements, staged pipeline.
    Size-independent properties: device-side backward error of every system, the exact discrete
    solution on a sample, |u_j| == 1 (the Neumann+Sommerfeld solution is a pure phase)."""
    N, n_k = 4096, 65536
    ks = np.linspace(1.0, 1000.0, n_k)
    nodes, U, status = fh.solve_helmholtz_sweep(N, ks, return_status=True)
    assert int(status.abs().sum()) == 0 and U.shape == (n_k, N + 1)
    dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
    res = fh.tridiagonal_residual(dl, d, du, b, U)
    # unpivoted CR / PCR: an interface pivot ~cos(2^s * 8 * theta) comes within ~1e-5 of zero for a few of the
    # 65,536 wavenumbers (worst backward error 5e-10 before refinement, 39 systems above 1e-12); those systems are
    # flagged by the kernel and refined once.  Bar: 1e-10; LAPACK: 2e-17.
    assert res["backward"].max() < 1e-13, res["backward"].max()
    assert np.median(res["backward"]) < 1e-15
    # the raw kernel flags every system it hurt: what is NOT flagged is already at 1e-13 without any refinement
    U0, st0 = fh.solve_tridiagonal(dl, d, du, b, refine=False)
    raw = fh.tridiagonal_residual(dl, d, du, b, U0)["backward"]
    flagged = (_np(st0) & 2) != 0
    assert raw[~flagged].max() < 1e-13 and raw.max() < 1e-7, (raw.max(), raw[~flagged].max(), flagged.sum())
    assert flagged.mean() < 0.05                           # a few per cent of the sweep
    assert float((U.abs() - 1.0).abs().max()) < 1e-6
    for i in (0, 1, 4097, 32768, 65535):
        ex = banded1d.discrete_exact(N, float(ks[i]))
        tol = 5e-9 if ks[i] < 50 else 1e-10
        assert rel_err(_np(U[i]), ex) < tol, (i, rel_err(_np(U[i]), ex))


def test_l2_error_and_convergence_rates_match_notebook(golden1d, capsys):
    """f2: compute_L2_error / convergence_test on the device reproduce the reference's numbers, including the
    notebook's stored stdout (nb:439-444) character for character."""
    import json
    import os
    from conftest import GOLDEN
    g = golden1d
    for idx, N, k, L in golden_cases(g):
        h, nodes, elements = fh.create_mesh(N, L)
        l2 = fh.compute_L2_error(k, N, elements, nodes, g[f"c{idx}_u"])
        # (N=4096, k=1: the error is 9e-9, i.e. the summed squares are ~8e-17 -- absolute floor for rounding)
        assert abs(l2 - g[f"c{idx}_l2"]) <= 1e-10 * abs(g[f"c{idx}_l2"]) + 1e-16, (idx, l2, g[f"c{idx}_l2"])
    rates = fh.convergence_test([1, 5, 10, 15, 30, 100], [10, 20, 40, 80, 160])
    stored = json.load(open(os.path.join(GOLDEN, "versions.json")))["notebook_stored_stdout"]
    assert capsys.readouterr().out == stored
    assert np.allclose(np.array(list(rates.values())), g["conv_rates"], rtol=1e-7)


def test_zero_pivot_is_resolved_by_the_pivoted_fallback():
    """k*h = 2 makes the first diagonal entry exactly zero (-1/h + k^2 h/4): harmless for LAPACK's pivoted zgesv
    that the reference calls, fatal for an unpivoted sweep.  The status flag routes such systems to the pivoted
    device fallback, so the result still matches the reference and no LinAlgError is raised."""
    for N, k in ((1, 2.0), (2, 4.0), (64, 128.0), (1000, 2000.0)):
        dl, d, du, b = banded1d.assemble_bands(N, k, 1.0 / N)
        assert d[0] == 0
        nodes, u = fh.solve_helmholtz_fem(N, k)
        assert rel_err(u, banded1d.solve_bands(dl, d, du, b)) < 1e-12, (N, k)
    # mixed batch: only the flagged systems take the fallback
    ks = [3.0, 128.0, 77.0]
    _, U, status = fh.solve_helmholtz_sweep(64, ks, return_status=True)
    assert _np(status).tolist() == [0, 0, 0]
    for i, k in enumerate(ks):
        assert rel_err(_np(U[i]), banded1d.solve_bands(*banded1d.assemble_bands(64, k, 1.0 / 64))) < 1e-11
    # without the fallback the raw kernel reports it
    bands = fh.assemble_bands(64, 128.0, 1.0 / 64)
    _, st = fh.solve_tridiagonal(*bands, refine=False)
    assert int(st[0]) & 1


def test_breakdown_flag_on_a_verified_solution_is_cleared_without_the_pivoted_resolve():
    """repair_flagged (device-driven): a system that carries the breakdown bit but whose (refined) solution verifiably is
    within the acceptance bar is kept untouched; a real breakdown in the same batch goes through the pivoted kernel; a
    finite but wrong solution does too."""
    N, ks = 512, [3.0, 40.0, 77.0, 21.5]
    dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
    u, status = fh.solve_tridiagonal(dl, d, du, b)
    assert int(status.abs().sum()) == 0
    ref = u.clone()
    status[1] = 1              # as if the probe had found a severe defect (3) and the refinement step had run since (-> 1)
    u[2] = float("nan")        # a real breakdown
    status[2] = 1
    u[3, 100] += 1e-6          # finite, but five digits short of the bar
    status[3] = 1
    counts = fh.fem1d.repair_flagged(dl, d, du, b, u, status)
    assert counts == {"flagged": 3, "kept": 1, "resolved": 2, "singular": 0}, counts
    assert int(status.abs().sum()) == 0
    assert torch.equal(u[0], ref[0]) and torch.equal(u[1], ref[1])
    assert rel_err(_np(u[2]), _np(ref[2])) < 1e-10 and rel_err(_np(u[3]), _np(ref[3])) < 1e-10
    assert fh.tridiagonal_residual(dl, d, du, b, u)["backward"].max() < 1e-15


def _call_pivoted(dl, d, du, b):
    import ctypes as C  # noqa: F401
    from fem_helmholtz_eqn_b200._lib import as_device, check, ptr, stream_ptr
    lib = fh._lib.require_cuda()
    dev = [as_device(np.atleast_2d(x), torch.complex128) for x in (dl, d, du, b)]
    batch, n = dev[1].shape
    x = torch.empty_like(dev[3])
    st = torch.full((batch,), 7, dtype=torch.int32, device="cuda")
    ws_bytes = lib.fh_tridiag_solve_pivoted_workspace_bytes(n, batch)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    check(lib.fh_tridiag_solve_pivoted_c128(*(ptr(t) for t in dev), ptr(x), n, batch, ptr(st), ptr(ws), ws_bytes,
                                            stream_ptr()), "fh_tridiag_solve_pivoted_c128")
    return _np(x), _np(st)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 127, 128, 129, 130, 255, 256, 257, 258, 259, 513, 4097, 4112, 5001, 100_003])
def test_pivoted_kernel_matches_lapack_on_systems_that_need_pivoting(n):
    """fh_tridiag_solve_pivoted_c128 (one CTA per system, elimination from both ends, tiles of 128 rows per side): sizes
    around the tile boundaries, exact zeros on the diagonal (every third entry), batch of 3 (grid-stride over slots)."""
    rng = np.random.default_rng(n)
    batch = 3 if n < 10_000 else 1
    dl, d, du, b = (rng.normal(size=(batch, n)) + 1j * rng.normal(size=(batch, n)) for _ in range(4))
    if n > 3:
        d[:, ::3] = 0
    x, st = _call_pivoted(dl, d, du, b)
    assert st.tolist() == [0] * batch
    for i in range(batch):
        ref = banded1d.solve_bands(dl[i], d[i], du[i], b[i])
        assert rel_err(x[i], ref) < 1e-9, (n, i, rel_err(x[i], ref))
        # (random hollow matrices are not well conditioned: the backward error is the sharp gate)
        assert banded1d.residual_metrics(dl[i], d[i], du[i], b[i], x[i])[0] < 5e-15, n


def test_pivoted_kernel_reports_singular_systems_only():
    n = 9
    one, zero = np.ones(n, complex), np.zeros(n, complex)
    rng = np.random.default_rng(1)
    ok = [rng.normal(size=n) + 1j * rng.normal(size=n) for _ in range(4)]
    dl, d, du, b = (np.stack([s, g]) for s, g in zip((one, zero, one, one), ok))  # system 0: odd hollow matrix
    x, st = _call_pivoted(dl, d, du, b)
    assert st.tolist() == [1, 0]
    assert rel_err(x[1], banded1d.solve_bands(*ok)) < 1e-10


def test_eight_rank_sweep_shards_near_the_partition_resonance():
    """in the 8-rank sweep (8 x 65,536 wavenumbers over [1, 1000]) rank 3 owns k = 806.8412377190356,
    a resonance of the batched kernel's partition (raw backward error 8e-8 -> severe flag).  Every shard's wavenumbers
    around it go through the real kernels here: after refinement + device-side repair nothing may be above 1e-13, no
    system may be left flagged, and the pivoted re-solve must not be needed (one refinement step reaches ~2e-15)."""
    N, world, n_k = 4096, 8, 65536
    ks_all = np.linspace(1.0, 1000.0, n_k * world)
    worst, resolved, severe = 0.0, 0, 0
    for r in range(world):
        ks = ks_all[r::world]
        ks = ks[np.abs(ks - 806.8412377190356) < 0.75]
        assert len(ks) > 90
        dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
        u, status = fh.solve_tridiagonal(dl, d, du, b, refine=False)
        severe += int(((status & 1) != 0).sum())
        fh.fem1d.refine_flagged(dl, d, du, b, u, status)
        counts = fh.fem1d.repair_flagged(dl, d, du, b, u, status)
        resolved += counts["resolved"]
        assert int(status.abs().sum()) == 0, (r, _np(status).nonzero())
        worst = max(worst, float(fh.tridiagonal_residual(dl, d, du, b, u)["backward"].max()))
    assert severe >= 1          # the resonance is in there
    assert worst < 1e-13, worst
    assert resolved == 0, resolved


@pytest.mark.parametrize("fused", [False, True])
def test_host_sweep_repairs_breakdowns_and_flagged_systems(fused):
    """solve_helmholtz_sweep_host and the fused sweep keep the accuracy of the staged device path: systems the
    unpivoted kernels flag (small pivot) or break on (k*h = 2: exact zero pivot) are solved again through the staged
    path and their rows of the result patched -- also in the fused variant, which keeps no bands to refine from."""
    N = 4096
    ks = np.concatenate([np.linspace(900.0, 1000.0, 2047), [2.0 * N]])        # last entry: k*h = 2
    nodes, U = fh.fem1d.solve_helmholtz_sweep_host(N, ks, fused=fused)
    dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
    res = fh.tridiagonal_residual(dl, d, du, b, U)["backward"]
    assert res.max() < 1e-13, res.max()
    _, U_dev = fh.solve_helmholtz_sweep(N, ks, fused=fused)
    assert fh.tridiagonal_residual(dl, d, du, b, U_dev)["backward"].max() < 1e-13
    assert rel_err(U[-1], banded1d.solve_bands(*banded1d.assemble_bands(N, float(ks[-1]), 1.0 / N))) < 1e-11


@pytest.mark.parametrize("N,bc", [(4096, {}), (4095, {}), (1000, {}), (2056, {}), (4111, {}), (64, {}),
                                  (999, dict(bc_left="dirichlet", bc_right="dirichlet", g_left=1.0, g_right=0.0))])
def test_one_pass_assemble_and_solve_keeps_identical_bands(N, bc):
    """fh_helmholtz1d_sweep_assemble_solve_c128: the bands it writes are what fh_assemble_1d_c128 writes (bit for bit:
    cut row, head rows, both cluster halves, single-CTA sizes, Dirichlet rows) and the solution is what the batched
    solver computes from those bands (bit for bit)."""
    ks = np.linspace(3.0, 700.0, 149 * 2 + 5)
    bands, U, status = fh.assemble_and_solve(N, ks, refine=False, **bc)
    ref = fh.assemble_bands(N, ks, 1.0 / N, **bc)
    for got, want in zip(bands, ref):
        assert torch.equal(got, want)
    U2, st2 = fh.solve_tridiagonal(*ref, refine=False)
    assert torch.equal(U, U2) and torch.equal(status, st2)
    _, U3, st3 = fh.assemble_and_solve(N, ks, **bc)          # with refinement from the stored bands
    assert int(st3.abs().sum()) == 0
    assert fh.tridiagonal_residual(*ref, U3)["backward"].max() < 1e-13
