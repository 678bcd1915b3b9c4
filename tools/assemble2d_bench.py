"""Throughput of the 2D Q4 assembly (element matrices -> deterministic CSR scatter) against mesh size."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_helmholtz_eqn_b200.helmholtz import HelmHoltz  # noqa: E402
from fem_helmholtz_eqn_b200.utils import get_solver  # noqa: E402

for nt, nr in ((60, 20), (256, 64), (1024, 256), (2048, 512)):
    t0 = time.perf_counter()
    eqn = HelmHoltz(inner_radius=1.0, outer_radius=1.5, n_theta=nt, n_r=nr, mesher="structured")
    t_mesh = time.perf_counter() - t0
    s = get_solver("dirichlet_sommerfeld", eqn, k_squared=30.25)
    s.assemble()                       # symbolic phase (sort, CSR pattern) is cached after the first call
    s.assemble()                       # (the second call still grows the allocator's pool)
    torch.cuda.synchronize()
    t_num = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        s.assemble()
        torch.cuda.synchronize()
        t_num = min(t_num, time.perf_counter() - t0)
    s2 = get_solver("dirichlet_sommerfeld", eqn, k_squared=30.25)
    t0 = time.perf_counter()
    s2.assemble()
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("nodes %8d elements %8d nnz %9d | mesh %.3f s | symbolic+numeric %.1f ms | numeric only %.1f ms = %.2e elements/s" % (
        eqn.n_nodes, eqn.n_elements, s.K.nnz, t_mesh, t_all * 1e3, t_num * 1e3, eqn.n_elements / t_num), flush=True)
