"""numpy model of cta_pivoted_solve (csrc/fh_residual.cu): Gaussian elimination with partial pivoting between adjacent
rows (LAPACK zgtsv's rule, cabs1), run from BOTH ends of the system; the two sweeps meet in a pivoted 2 x 2 system for
(x_m, x_{m+1}) and back-substitute outwards.  Same index arithmetic as the kernel (side, local row j, factor-row slot),
so the CPU suite checks the bookkeeping; the GPU tests check the kernel against LAPACK."""
import numpy as np


def cabs1(z):
    return abs(z.real) + abs(z.imag)


def solve(dl, d, du, b):
    n = len(d)
    x = np.zeros(n, dtype=complex)
    if n == 1:
        x[0] = b[0] / d[0]
        return x
    m = (n - 2) // 2
    mp = n - 2 - m
    cnt = (m + 1, mp + 1)
    fac = np.zeros((n, 4), dtype=complex)
    mid = [None, None]
    with np.errstate(all="ignore"):
        for side in (0, 1):
            dd = u1 = rr = 0j
            for j in range(cnt[side]):
                g = j if side == 0 else n - 1 - j
                lo = (dl if side == 0 else du)[g]
                di = d[g]
                hi = (du if side == 0 else dl)[g]
                rh = b[g]
                if j == 0:
                    dd, u1, rr = di, hi, rh
                    continue
                swap = cabs1(dd) < cabs1(lo)
                if swap:
                    p0, p1, p2, pr, o0, o1, o2, orr = lo, di, hi, rh, dd, u1, 0j, rr
                else:
                    p0, p1, p2, pr, o0, o1, o2, orr = dd, u1, 0j, rr, lo, di, hi, rh
                inv = 1.0 / p0
                mlt = o0 * inv
                fac[(j - 1) if side == 0 else (n - j)] = (inv, p1, p2, pr)
                dd, u1, rr = o1 - mlt * p1, o2 - mlt * p2, orr - mlt * pr
            mid[side] = (dd, u1, rr)
        (a00, a01, r0), (a11, a10, r1) = mid
        if cabs1(a00) >= cabs1(a10):
            w = a10 / a00
            xm1 = (r1 - w * r0) / (a11 - w * a01)
            xm = (r0 - a01 * xm1) / a00
        else:
            w = a00 / a10
            xm1 = (r0 - w * r1) / (a01 - w * a11)
            xm = (r1 - a11 * xm1) / a10
        x[m], x[m + 1] = xm, xm1
        xmid = (xm, xm1)
        for side in (0, 1):
            xa, xb = xmid[side], xmid[side ^ 1]
            for j in range(cnt[side] - 2, -1, -1):
                g = j if side == 0 else n - 1 - j
                inv, f1, f2, fr = fac[g]
                v = (fr - f2 * xb - f1 * xa) * inv
                x[g] = v
                xb, xa = xa, v
    return x
