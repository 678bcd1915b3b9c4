"""End-to-end time of the 2D drop-in call `get_solver(kind, eqn).solve()` (assemble -> boundary rows -> banded LU ->
(u_real, u_imag) on the host) against mesh size, with its parts, next to the oracle's dense NumPy path on the host cores
(reference meshes: ~225 nodes).    python tools/solve2d_bench.py [--cpu]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_helmholtz_eqn_b200.helmholtz import HelmHoltz  # noqa: E402
from fem_helmholtz_eqn_b200.utils import get_solver  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        out.append(time.perf_counter() - t0)
    return 1e3 * min(out)


for nt, nr in ((15, 14), (60, 20), (120, 40), (240, 60)):
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=nt, n_r=nr, mesher="structured")
    for kind in ("default", "dirichlet_sommerfeld"):
        s = get_solver(kind, eqn, k_squared=30.25)
        t_all = timed(lambda: s.solve())
        t_asm = timed(lambda: s.assemble())
        line = "%6d nodes %-22s solve() %8.2f ms (assemble %6.2f ms)" % (eqn.n_nodes, kind, t_all, t_asm)
        if "--cpu" in sys.argv and eqn.n_nodes <= 1300:
            from oracle import ref2d  # noqa: E402  (development comparison only)
            m = ref2d.structured_annulus(nr, nt, 1.0, 1.5)
            t0 = time.perf_counter()
            ref2d.solve(kind, m, 30.25, 1, 3, 1.0, 1.5)
            line += "   oracle (dense NumPy, host): %.0f ms" % (1e3 * (time.perf_counter() - t0))
        print(line)
