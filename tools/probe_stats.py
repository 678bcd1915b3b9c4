import sys, numpy as np, torch
sys.path.insert(0, '.')
import fem_helmholtz_eqn_b200 as fh
N, n_k = 4096, 65536
ks = np.linspace(1.0, 1000.0, n_k)
dl, d, du, b = fh.assemble_bands(N, ks, 1.0 / N)
U, st = fh.solve_tridiagonal(dl, d, du, b, refine=False)
res = fh.tridiagonal_residual(dl, d, du, b, U)["backward"]
flag = (st.cpu().numpy() & 2) != 0
print("flagged", flag.sum(), "raw max", res.max(), "raw max unflagged", res[~flag].max(), "median", np.median(res))
for thr in (1e-16, 3e-16, 1e-15, 3e-15, 1e-14, 1e-13, 1e-12):
    print("backward >= %.0e: all %6d flagged %6d" % (thr, (res >= thr).sum(), (res[flag] >= thr).sum()))
q = np.quantile(res[flag], [0.1, 0.5, 0.9, 0.99])
print("flagged quantiles", q)
