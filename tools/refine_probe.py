"""Time of the safety nets alone (refinement of the flagged systems, then the repair pass) on the bench's sweep: the
one-pass kernel runs once, its status flags are restored before every repetition.
    python tools/refine_probe.py [n_k]        (FH_LIB=... for a variant build)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402
from fem_helmholtz_eqn_b200 import fem1d  # noqa: E402

n_k = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = 4096
ks = np.linspace(1.0, 1000.0, n_k)
bands, U, status = fh.assemble_and_solve(N, ks, refine=False)
dl, d, du, b = bands
st0 = status.clone()
U0 = U.clone()
lib = fh._lib.require_cuda()
ws = torch.empty(lib.fh_tridiag_refine_flagged_workspace_bytes(N + 1, n_k, max(256, n_k // 8)), dtype=torch.uint8, device="cuda")
times = []
for rep in range(8):
    status.copy_(st0)
    U.copy_(U0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fem1d.refine_flagged(dl, d, du, b, U, status, sync=False, ws=ws)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
flagged, done = (int(v) for v in ws[:8].view(torch.int32).tolist())
res = fh.tridiagonal_residual(dl, d, du, b, U)["backward"]
print("refine: %d flagged, %d refined, median %.1f us, min %.1f us | backward error max %.2e | flags left %d"
      % (flagged, done, 1e3 * float(np.median(times[2:])), 1e3 * min(times[2:]), float(res.max()), int((status & 2).sum())))
