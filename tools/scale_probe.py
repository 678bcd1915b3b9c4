import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import fem_helmholtz_eqn_b200 as fh
N = 4096
def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for n_k in (4440, 8880, 17760, 35520, 65536):
    ks = np.linspace(1.0, 1000.0, n_k)
    B = torch.empty((4, n_k, N + 1), dtype=torch.complex128, device="cuda")
    U = torch.empty((n_k, N + 1), dtype=torch.complex128, device="cuda")
    t1 = timed(lambda: fh.assemble_and_solve(N, ks, out_bands=B, out=U, refine=False))
    t2 = timed(lambda: fh.solve_helmholtz_sweep(N, ks, fused=True, check_singular=False, out=U))
    t3 = timed(lambda: fh.solve_tridiagonal(B[0], B[1], B[2], B[3], out=U, refine=False))
    print("n_k %6d: one-pass %.3f ms (%.2f us/system-iteration)  fused %.3f (%.2f)  staged solve %.3f (%.2f)" % (
        n_k, t1, t1 * 1e3 / (n_k / 74), t2, t2 * 1e3 / (n_k / 74), t3, t3 * 1e3 / (n_k / 74)))
    del B, U
