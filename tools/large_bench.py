"""Throughput of the single-large-system solver (fh_tridiag_solve_large_c128 and the fused generator) against N."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402


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
    return float(np.median(ts))


def main():
    out = []
    k = 1000.0
    for logn in (20, 22, 24, 26, 27):
        N = 2 ** logn
        n = N + 1
        dl, d, du, b = fh.assemble_bands(N, k, 1.0 / N)
        t_asm = timed(lambda: fh.assemble_bands(N, k, 1.0 / N)) if logn <= 24 else float("nan")
        u = torch.empty_like(d)
        t_solve = timed(lambda: fh.solve_tridiagonal(dl, d, du, b, out=u, refine=False))
        res = fh.tridiagonal_residual(dl, d, du, b, u)["backward"].max()
        del dl, du, b
        t_fused = timed(lambda: fh.solve_helmholtz_sweep(N, [k], fused=True, check_singular=False, out=u.reshape(1, -1)))
        row = {"N": N, "assemble_ms": t_asm, "solve_ms": t_solve, "solve_gbs": 144.0 * n / t_solve / 1e6,
               "solve_dof_per_s": n / t_solve * 1e3, "fused_ms": t_fused, "fused_dof_per_s": n / t_fused * 1e3,
               "backward_error": float(res)}
        print(json.dumps(row), flush=True)
        out.append(row)
        del d, u
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
