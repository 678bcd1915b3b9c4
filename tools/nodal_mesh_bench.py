import sys, numpy as np, torch
sys.path.insert(0, '.')
import fem_helmholtz_eqn_b200 as fh
N, n_k = 4096, 65536
ks = np.linspace(1.0, 1000.0, n_k)
nodes = np.linspace(0, 1.0, N + 1)
rng = np.random.default_rng(0)
nodes[1:-1] += rng.uniform(-0.25, 0.25, N - 1) / N
B = torch.empty((4, n_k, N + 1), dtype=torch.complex128, device="cuda")
def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ksd = (ks, ks**2)
t_s = timed(lambda: fh.assemble_bands(N, ksd, 1.0 / N, out=B))
t_n = timed(lambda: fh.assemble_bands(N, ksd, 1.0 / N, nodes=nodes, out=B))
dof = n_k * (N + 1)
print("assembly scalar-h %.3f ms (%.0f GB/s of 64 B/DOF); nodal-h %.3f ms (%.0f GB/s of 72 B/DOF)" % (t_s, 64 * dof / t_s / 1e6, t_n, 72 * dof / t_n / 1e6))
U, st = fh.solve_tridiagonal(B[0], B[1], B[2], B[3])
print("jittered-mesh sweep: backward error max %.2e, flagged-after %d" % (fh.tridiagonal_residual(B[0], B[1], B[2], B[3], U)["backward"].max(), int((st != 0).sum())))
