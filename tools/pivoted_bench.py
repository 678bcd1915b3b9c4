"""Time of the pivoted fallback kernel (fh_tridiag_solve_pivoted_c128) and of the device-driven repair pass.
    gpurun -- python tools/pivoted_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402
from fem_helmholtz_eqn_b200._lib import check, ptr, stream_ptr  # noqa: E402

lib = fh._lib.require_cuda()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


for N, batch in ((4096, 1), (4096, 16), (4096, 256), (4096, 1024), (1000, 1), (100_000, 1), (1_000_000, 1)):
    ks = np.linspace(800.0, 810.0, batch)
    dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
    n = N + 1
    x = torch.empty_like(b)
    st = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ws_bytes = lib.fh_tridiag_solve_pivoted_workspace_bytes(n, batch)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: check(lib.fh_tridiag_solve_pivoted_c128(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(x), n, batch, ptr(st),
                                                               ptr(ws), ws_bytes, stream_ptr()), "pivoted"))
    be = fh.tridiagonal_residual(dl, d, du, b, x)["backward"].max()
    print(f"pivoted n={n} batch={batch}: {ms:.3f} ms  ({1e3 * ms / n:.1f} ns per row and system-in-flight), backward error {be:.1e}")

# repair pass on a clean batch (nothing flagged): the cost every step pays
N, batch = 4096, 65536
ks = np.linspace(1.0, 1000.0, batch)
dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
u, status = fh.solve_tridiagonal(dl, d, du, b)
ms0 = timed(lambda: fh.fem1d.repair_flagged(dl, d, du, b, u, status, sync=False))
status[1000] = 1
u[1000] = float("nan")
def one():
    status[1000] = 1
    fh.fem1d.repair_flagged(dl, d, du, b, u, status, sync=False)
ms1 = timed(one)
print(f"repair pass over {batch} systems: nothing flagged {ms0:.3f} ms; one system re-solved {ms1:.3f} ms")
