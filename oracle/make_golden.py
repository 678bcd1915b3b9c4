"""Generate tests/golden/*.npz by executing the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python -m oracle.make_golden

* 1D: code cell 2 of notebooks/1D_soln.ipynb is exec'd as is (matplotlib stubbed --
  it is only touched by the plotting helpers).
* 2D: ``src.utils.get_solver`` and the four solver classes are imported with a stub
  ``gmsh`` module (helmholtz.py:1 imports it; gmsh is not installable offline); the
  ``HelmHoltz`` object is created without its constructor and given a structured
  annulus mesh, then the reference classes run unmodified.

Nothing from the reference is copied into the repo: only numeric outputs are stored.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import scipy

REF = os.environ.get("FH_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_notebook_namespace():
    nb = json.load(open(os.path.join(REF, "notebooks", "1D_soln.ipynb")))
    code_cells = [c for c in nb["cells"] if c["cell_type"] == "code"]
    plt = types.ModuleType("matplotlib.pyplot")
    ns = {"np": np, "plt": plt, "__name__": "reference_nb"}
    exec(compile("".join(code_cells[1]["source"]), "1D_soln.ipynb#cell2", "exec"), ns)
    stored = "".join(code_cells[3]["outputs"][0]["text"])
    return ns, stored


def golden_1d(ns):
    out = {}
    # element / boundary matrices (a3, a4)
    hs = np.array([1.0 / 999, 2.0 / 100, 1.0 / 4096, 1.0 / 1600, 0.3, 1e-6, 1.0 / 3])
    out["em_h"] = hs
    out["em_Ke"] = np.array([ns["element_matrices"](float(h))[0] for h in hs])
    out["em_Me"] = np.array([ns["element_matrices"](float(h))[1] for h in hs])
    kth = np.array([[10.0, 0.0], [5.0, 0.0], [1000.0, 0.0], [37.5, 0.4], [2.0, 1.2]])
    out["bm_k_theta"] = kth
    out["bm_Ne"] = np.array([ns["boundary_matrices"](0.1, float(k), float(t))[0] for k, t in kth])
    out["bm_Se"] = np.array([ns["boundary_matrices"](0.1, float(k), float(t))[1] for k, t in kth])
    xi = np.array([-1.0, -0.5, 0.0, 0.3, 1.0])
    out["sf_xi"] = xi
    out["sf_phi"] = np.array([ns["shape_functions"](float(x))[0] for x in xi])
    out["sf_dphi"] = np.array([ns["shape_functions"](float(x))[1] for x in xi])

    # full solves: (N, k, L) -- bands of the reference's dense A, b and u (a1, a5, a6)
    cases = [
        (100, 5, 2.0),        # notebook demo cell (nb:400-402), integer k as written there
        (8, 3.0, 1.0),
        (10, 1, 1.0),         # first point of the notebook convergence sweep
        (64, 40.0, 1.0),
        (999, 10.0, 1.0),     # BASELINE config 1 (1,000 nodes)
        (1600, 1000.0, 1.0),  # BASELINE config 5 (k=1000, 10 points per wavelength)
        (2049, 77.25, 1.0),
        (4096, 1.0, 1.0),     # three units of BASELINE config 3's sweep
        (4096, 500.5, 1.0),
        (4096, 1000.0, 1.0),
    ]
    out["cases"] = np.array([[c[0], float(c[1]), c[2]] for c in cases])
    for idx, (N, k, L) in enumerate(cases):
        h, nodes, elements = ns["create_mesh"](N, L)
        A, b = ns["assembly"](N, k, h, elements)
        # the whole matrix must be tridiagonal: everything off the three bands is 0
        band_mass = (np.abs(np.diag(A)).sum() + np.abs(np.diag(A, 1)).sum()
                     + np.abs(np.diag(A, -1)).sum())
        assert np.isclose(np.abs(A).sum(), band_mass, rtol=1e-14)
        nodes2, u = ns["solve_helmholtz_fem"](N, k, L=L)
        assert np.array_equal(nodes, nodes2)
        out[f"c{idx}_h"] = np.float64(h)
        out[f"c{idx}_nodes"] = nodes
        out[f"c{idx}_elements_head"] = elements[:4]
        out[f"c{idx}_dl"] = np.concatenate([[0], np.diag(A, -1)])
        out[f"c{idx}_d"] = np.diag(A).copy()
        out[f"c{idx}_du"] = np.concatenate([np.diag(A, 1), [0]])
        out[f"c{idx}_b"] = b
        out[f"c{idx}_u"] = u
        if N <= 999:
            # Dirichlet row replacement, fem2d_dirichlet.py:20-27 convention, applied
            # to the reference's own (A, b): u(0)=1, u(L)=0  (BASELINE config 1), and
            # Dirichlet-left + Sommerfeld-right (fem2d_dirichlet_sommerfeld.py:17-22)
            A1, b1 = A.copy(), b.copy()
            A1[0] = 0; A1[0, 0] = 1; b1[0] = 1
            out[f"c{idx}_u_ds"] = np.linalg.solve(A1, b1)
            A1[-1] = 0; A1[-1, -1] = 1; b1[-1] = 0
            out[f"c{idx}_u_dd"] = np.linalg.solve(A1, b1)
        # L2 error with the reference's quirks (f2)
        out[f"c{idx}_l2"] = np.complex128(ns["compute_L2_error"](k, N, elements, nodes, u))
        print(f"1D case {idx}: N={N} k={k} L={L} done", flush=True)

    # convergence study exactly as notebook cell 4 (nb:469-471), without plotting
    N_values = [10, 20, 40, 80, 160]
    k_values = [1, 5, 10, 15, 30, 100]
    hs_, errs = [], {k: [] for k in k_values}
    for N in N_values:
        h, nodes, elements = ns["create_mesh"](N)
        hs_.append(h)
        for k in k_values:
            _, u = ns["solve_helmholtz_fem"](N, k, order=0)
            errs[k].append(np.sqrt(ns["compute_L2_error"](k, N, elements, nodes, u)))
    rates = [np.polyfit(np.log(hs_), np.log(errs[k]), 1)[0] for k in k_values]
    out["conv_k"] = np.array(k_values)
    out["conv_N"] = np.array(N_values)
    out["conv_rates"] = np.array(rates)
    out["conv_errs"] = np.array([errs[k] for k in k_values])
    return out


def golden_2d():
    sys.modules.setdefault("gmsh", types.ModuleType("gmsh"))
    sys.path.insert(0, REF)
    from src.helmholtz import HelmHoltz  # noqa: E402
    from src.utils import get_solver, solvers  # noqa: E402

    from oracle.ref2d import structured_annulus

    out = {"solver_names": np.array(solvers)}
    meshes = {"m0": (3, 12, 1.0, 1.5), "m1": (6, 24, 1.0, 1.5)}
    runs = [("default", 0), ("default", 1), ("default", 3), ("dirichlet", 3),
            ("dirichlet_sommerfeld", 0), ("dirichlet_sommerfeld", 2), ("neumann_dirichlet", 3)]
    for mname, (n_r, n_t, ri, ro) in meshes.items():
        m = structured_annulus(n_r, n_t, ri, ro)
        eqn = HelmHoltz.__new__(HelmHoltz)
        for key, val in vars(m).items():
            setattr(eqn, key, val)
        eqn.abc_order = 3
        out[f"{mname}_nodes"] = m.nodes
        out[f"{mname}_elements"] = m.elements
        out[f"{mname}_inner_el"] = np.array(m.inner_boundary_element_indices)
        out[f"{mname}_outer_el"] = np.array(m.outer_boundary_element_indices)
        out[f"{mname}_radii"] = np.array([ri, ro])
        for kind, order in runs:
            for k2 in (5.0, 30.25):
                s = get_solver(kind, eqn, k_squared=k2, abc_order=order, inner_radius=ri,
                               outer_radius=ro)
                s.assemble()
                tag = f"{mname}_{kind}_o{order}_k{k2}"
                Kpre, Fpre = s.K.copy(), s.F.copy()
                ur, ui = s.solve()           # re-assembles, applies BC rows, dense solve
                for label, M in (("Kpre", Kpre), ("K", s.K)):
                    r, c = np.nonzero(M)
                    out[f"{tag}_{label}_rows"] = r.astype(np.int32)
                    out[f"{tag}_{label}_cols"] = c.astype(np.int32)
                    out[f"{tag}_{label}_vals"] = M[r, c]
                out[f"{tag}_Fpre"] = Fpre
                out[f"{tag}_F"] = s.F
                out[f"{tag}_u"] = ur + 1j * ui
                if kind == "default" and order == 3 and k2 == 5.0:
                    Ke, Fe = s.get_element_matrices(0)
                    out[f"{mname}_Ke0"] = Ke
                    out[f"{mname}_Se_outer0"] = s.sommerfeld_element_matrices(
                        m.outer_boundary_element_indices[0])
                    out[f"{mname}_Ne_inner0"] = s.neumann_element_matrices(
                        m.inner_boundary_element_indices[0])
                print("2D", tag, "nnz", len(out[f"{tag}_K_vals"]), flush=True)
        for order in (0, 1, 2, 3):
            out[f"{mname}_abc{order}"] = np.complex128(eqn.get_abc_coeff(5.0, order))
    return out


def golden_2d_post():
    """What follows the solve in the reference's 2D drivers: every class's analytic solution at the mesh nodes and
    BaseSolver.compute_L2_error of the reference's own solution (base.py:119-177), the reference run unmodified."""
    sys.modules.setdefault("gmsh", types.ModuleType("gmsh"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.helmholtz import HelmHoltz  # noqa: E402
    from src.utils import get_solver  # noqa: E402

    from oracle.ref2d import structured_annulus

    out = {}
    meshes = {"m0": (3, 12, 1.0, 1.5), "m1": (6, 24, 1.0, 1.5)}
    runs = [("default", 0), ("default", 1), ("default", 3), ("dirichlet", 3),
            ("dirichlet_sommerfeld", 0), ("dirichlet_sommerfeld", 2), ("neumann_dirichlet", 3)]
    for mname, (n_r, n_t, ri, ro) in meshes.items():
        m = structured_annulus(n_r, n_t, ri, ro)
        eqn = HelmHoltz.__new__(HelmHoltz)
        for key, val in vars(m).items():
            setattr(eqn, key, val)
        eqn.abc_order = 3
        for kind, order in runs:
            for k2, nf in ((5.0, 1), (30.25, 2)):
                s = get_solver(kind, eqn, k_squared=k2, abc_order=order, inner_radius=ri, outer_radius=ro, n_fourier=nf)
                ur, ui = s.solve()
                tag = f"{mname}_{kind}_o{order}_k{k2}"
                out[f"{tag}_u"] = ur + 1j * ui
                out[f"{tag}_exact_nodes"] = np.array([s.get_analytical_solution(x, y) for x, y in m.nodes])
                out[f"{tag}_l2"] = np.float64(s.compute_L2_error(ur, ui))
                print("2D post", tag, "n_fourier", nf, "L2", out[f"{tag}_l2"], flush=True)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "post":   # only the (later) fixture of the steps after the solve
        np.savez_compressed(os.path.join(OUT, "ref2d_post.npz"), **golden_2d_post())
        print("wrote", os.path.join(OUT, "ref2d_post.npz"))
        return
    ns, stored = load_notebook_namespace()
    g1 = golden_1d(ns)
    np.savez_compressed(os.path.join(OUT, "ref1d.npz"), **g1)
    g2 = golden_2d()
    np.savez_compressed(os.path.join(OUT, "ref2d.npz"), **g2)
    np.savez_compressed(os.path.join(OUT, "ref2d_post.npz"), **golden_2d_post())
    meta = {
        "generator": "oracle/make_golden.py (executes /root/reference unmodified)",
        "numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0],
        "blas": str(np.show_config(mode="dicts").get("Build Dependencies", {}).get("blas", {})),
        "notebook_stored_stdout": stored,
    }
    json.dump(meta, open(os.path.join(OUT, "versions.json"), "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
