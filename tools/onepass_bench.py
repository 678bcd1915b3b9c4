"""One-pass assemble+solve (bands stored, never read back) against the two-kernel staged pipeline, 65,536 x 4,097."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402

N, n_k = 4096, 65536
ks = np.linspace(1.0, 1000.0, n_k)
B = torch.empty((4, n_k, N + 1), dtype=torch.complex128, device="cuda")
U = torch.empty((n_k, N + 1), dtype=torch.complex128, device="cuda")


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_raw = timed(lambda: fh.assemble_and_solve(N, ks, out_bands=B, out=U, refine=False))
t_ref = timed(lambda: fh.assemble_and_solve(N, ks, out_bands=B, out=U))
dof = n_k * (N + 1)
print("one-pass kernel %.3f ms (%.0f GB/s of 80 B/DOF), with refinement %.3f ms -> %.3e DOF/s" % (
    t_raw, 80.0 * dof / t_raw / 1e6, t_ref, dof / t_ref * 1e3))
res = fh.tridiagonal_residual(B[0], B[1], B[2], B[3], U)["backward"]
print("backward error max %.2e" % res.max())
