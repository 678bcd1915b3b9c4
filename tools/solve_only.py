"""The staged batched solve alone (65,536 x 4,097 stored bands), a few launches: the command behind the ncu captures
of the pipelined solver (ncu -k regex:tridiag_pipe -c 1 ...).   python tools/solve_only.py [n_k] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402

n_k = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ks = np.linspace(1.0, 1000.0, n_k)
bands = fh.assemble_bands(4096, ks, 1.0 / 4096)
U = torch.empty((n_k, 4097), dtype=torch.complex128, device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fh.solve_tridiagonal(*bands, out=U, refine=False)
torch.cuda.synchronize()
a.record()
for _ in range(reps):
    fh.solve_tridiagonal(*bands, out=U, refine=False)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print("staged solve %.3f ms = %.0f GB/s of 80 B/DOF" % (ms, 80.0 * n_k * 4097 / ms / 1e6))
