"""Time of the 2D sparse (banded LU) device solve against mesh size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch
from fem_helmholtz_eqn_b200.helmholtz import HelmHoltz
from fem_helmholtz_eqn_b200.utils import get_solver
for nt, nr in ((15,14),(60,20),(120,40)):
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=nt, n_r=nr, mesher="structured")
    for mode in ("band",):
        s = get_solver("dirichlet_sommerfeld", eqn, k_squared=30.25); s.linear_solver=mode
        s.assemble(); s.apply_boundary_conditions()
        s._solve_device(); torch.cuda.synchronize()
        t0=time.perf_counter(); x=s._solve_device(); torch.cuda.synchronize(); t1=time.perf_counter()
        print(eqn.n_nodes, mode, s._band_plan()[1:] if mode=="band" else "", "%.2f ms"%((t1-t0)*1e3))
