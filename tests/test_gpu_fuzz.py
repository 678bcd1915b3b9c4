"""Randomised parity of the 1D path (fixed seeds): mesh sizes across every kernel shape (single CTA, cluster of two, head
rows, long meshes), wavenumbers from 0.1 to 20 points per wavelength and beyond, all boundary-row variants, all three
pipelines.  Gates: bands bit-exact against the CPU oracle's (nb:134-187 restated in oracle/banded1d.py), the three
pipelines bit-identical to each other, normwise backward error <= 1e-13 (1e-12 with a reflecting right end; north-star bar
1e-10), forward error against
LAPACK zgtsv within what cond(A) allows."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import banded1d

pytestmark = pytest.mark.gpu

fh = pytest.importorskip("fem_helmholtz_eqn_b200")

BCS = [{}, dict(bc_left="dirichlet", bc_right="dirichlet", g_left=1.0, g_right=0.0),
       dict(bc_left="dirichlet", bc_right="sommerfeld", g_left=0.5), dict(bc_left="neumann", bc_right="dirichlet", g_right=-0.25)]


def _cases(seed, count):
    rng = np.random.default_rng(seed)
    sizes = [2, 3, 7, 8, 9, 15, 16, 17, 63, 64, 65, 255, 1000, 2047, 2048, 2055, 2056, 2057, 3000, 4095, 4096, 4111, 4112, 4113,
             5000, 20011]
    out = []
    for _ in range(count):
        N = int(rng.choice(sizes)) if rng.random() < 0.7 else int(rng.integers(2, 6000))
        L = float(rng.choice([1.0, 2.0, 0.37]))
        ppw = float(rng.choice([4.0, 10.0, 25.0, 100.0, 1e4]))            # points per wavelength
        kmax = 2 * np.pi * N / (L * ppw)
        ks = np.sort(rng.uniform(0.05 * kmax, kmax, size=int(rng.integers(1, 7))))
        out.append((N, L, ks, BCS[int(rng.integers(0, len(BCS)))], float(rng.uniform(-0.4, 0.4))))
    return out


@pytest.mark.parametrize("seed", [11, 12, 13, 109])  # 109: a wavenumber whose k**2 (libm pow) is not the rounded square
def test_random_meshes_wavenumbers_and_boundary_rows(seed):
    for N, L, ks, bc, theta in _cases(seed, 14):
        tag = (N, L, list(ks), bc, theta)
        _, U1, st1 = fh.solve_helmholtz_sweep(N, ks, L, theta, pipeline="staged", return_status=True, check_singular=False, **bc)
        _, U2, st2 = fh.solve_helmholtz_sweep(N, ks, L, theta, pipeline="one_pass", return_status=True, check_singular=False, **bc)
        _, U3, st3 = fh.solve_helmholtz_sweep(N, ks, L, theta, pipeline="fused", return_status=True, check_singular=False, **bc)
        assert int(st1.abs().sum()) == 0 and int(st2.abs().sum()) == 0 and int(st3.abs().sum()) == 0, tag
        assert torch.equal(U1, U2), tag                       # same bands, same arithmetic: identical bits
        bands, _, _ = fh.assemble_and_solve(N, ks, L, theta, refine=False, **bc)
        absorbing = bc.get("bc_right", "sommerfeld") == "sommerfeld"
        for i, k in enumerate(ks):
            ref = banded1d.assemble_bands(N, float(k), L / N, theta, **bc)
            for got, want in zip(bands, ref):
                assert np.array_equal(got[i].cpu().numpy(), want), tag
            for U in (U1, U3):
                u = U[i].cpu().numpy()
                be, _ = banded1d.residual_metrics(*ref, u)
                # (a reflecting right end allows near-resonant, ill-conditioned systems: one refinement step was seen to
                # leave 1.06e-13 at N = 4095, k = 2194.42 in 4,200 random cases; the contract's bar is 1e-10)
                assert be < (1e-13 if absorbing else 1e-12), (tag, i, be)
            if absorbing:  # well conditioned: forward error gated too
                assert rel_err(U1[i].cpu().numpy(), banded1d.solve_bands(*ref)) < 1e-7, (tag, i)


def _permuted_oracle_mesh(ref2d, n_r, n_t, ri, ro, perm):
    """The oracle's mesh namespace on a re-numbered mesh, boundary lists by the reference's own rules (helmholtz.py:95-118)."""
    m = ref2d.structured_annulus(n_r, n_t, ri, ro)
    inv = np.argsort(perm)
    m.nodes, m.elements = m.nodes[perm], inv[m.elements]
    dist = np.linalg.norm(m.nodes, axis=1)
    m.inner_boundary_node_indices = np.where(np.isclose(dist, ri, atol=1e-5))[0]
    m.outer_boundary_node_indices = np.where(np.isclose(dist, ro, atol=1e-5))[0]
    ins, outs = set(m.inner_boundary_node_indices.tolist()), set(m.outer_boundary_node_indices.tolist())
    m.inner_boundary_element_indices = [e for e, c in enumerate(m.elements) if ins.intersection(c.tolist())]
    m.outer_boundary_element_indices = [e for e, c in enumerate(m.elements) if outs.intersection(c.tolist())]
    return m


@pytest.mark.parametrize("seed", [3, 4])
def test_random_annuli_against_the_2d_oracle(seed):
    """The four 2D classes on random annuli (radii, mesh sizes down to one ring of three quads, k**2, Fourier terms, ABC
    orders, every third mesh with a scrambled numbering) against oracle/ref2d.py, the restatement of src/solvers pinned by
    tests/golden: pattern and values of K and F bit-exact, solutions within 1e-9."""
    from fem_helmholtz_eqn_b200.helmholtz import HelmHoltz
    from fem_helmholtz_eqn_b200.utils import get_solver
    from oracle import ref2d
    rng = np.random.default_rng(seed)
    for trial in range(12):
        n_r, n_t = int(rng.integers(1, 9)), int(rng.integers(3, 40))
        ri = float(rng.choice([1.0, 0.5, 2.0]))
        ro = ri * float(rng.uniform(1.2, 2.5))
        k2, nf, order = float(rng.uniform(0.5, 60.0)), int(rng.integers(1, 5)), int(rng.integers(0, 4))
        kind = ref2d.SOLVER_KINDS[trial % 4]
        perm = rng.permutation((n_r + 1) * n_t) if trial % 3 == 0 else np.arange((n_r + 1) * n_t)
        m = _permuted_oracle_mesh(ref2d, n_r, n_t, ri, ro, perm)
        tag = (seed, trial, kind, n_r, n_t, ri, ro, k2, nf, order)
        Kpre, Fpre = ref2d.assemble(kind, m, k2, nf, order)
        ur, ui, _, _ = ref2d.solve(kind, m, k2, nf, order, ri, ro)
        s = get_solver(kind, HelmHoltz.from_arrays(m.nodes, m.elements, ri, ro), k_squared=k2, n_fourier=nf, abc_order=order,
                       inner_radius=ri, outer_radius=ro)
        s.assemble()
        r, c, v = s.K.coo()
        keep = v != 0
        order_ = np.lexsort((c[keep], r[keep]))
        r0, c0 = np.nonzero(Kpre)
        assert np.array_equal(r[keep][order_], r0) and np.array_equal(c[keep][order_], c0), tag
        assert np.array_equal(v[keep][order_], Kpre[r0, c0]) and np.array_equal(s.F, Fpre), tag
        gr, gi = s.solve()
        assert rel_err(gr + 1j * gi, ur + 1j * ui) < 1e-9, tag
