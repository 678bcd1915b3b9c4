"""Largest relative defect seen by the interface-row probe of the large-system solvers (needs a library built with
-DFH_PROBE_STAT: tools/build_variant.sh stat -DFH_PROBE_STAT; FH_LIB=build/stat/libfemhelmholtz_b200.so), next to the
normwise backward error of the same solve -- what the probe threshold has to separate."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402
from fem_helmholtz_eqn_b200 import fem1d  # noqa: E402

lib = fh._lib.require_cuda()
lib.fh_debug_probe_max.restype = ctypes.c_double
lib.fh_debug_probe_max.argtypes = [ctypes.c_int]

for N in (5000, 131071, 300_017, 1_000_000, 2**24, 2**27):
    for k, bc in ((10.0, {}), (1000.0, {}), (37.25, {}), (3.0, dict(bc_left="dirichlet", bc_right="dirichlet", g_right=0.5))):
        plan = fem1d.LargeMeshPlan(N, fused=True, use_graph=False, **bc)
        lib.fh_debug_probe_max(1)
        plan.solve(k)
        torch.cuda.synchronize()
        pf, bf, sf = lib.fh_debug_probe_max(1), plan.backward_error(), int(plan.status.item())
        del plan
        line = f"N={N} k={k} {'DD' if bc else 'NS'}: fused probe max {pf:.2e} backward {bf:.2e} status {sf}"
        if N <= 2**24:
            plan = fem1d.LargeMeshPlan(N, use_graph=False, **bc)
            lib.fh_debug_probe_max(1)
            plan.solve(k)
            torch.cuda.synchronize()
            ps, bs, ss = lib.fh_debug_probe_max(1), plan.backward_error(), int(plan.status.item())
            del plan
            line += f" | staged probe max {ps:.2e} backward {bs:.2e} status {ss}"
        print(line, flush=True)
        torch.cuda.empty_cache()
