"""numpy model of the periodic recursion of csrc/fh_tridiag_large.cu (generated rows on a uniform mesh): per level two
constant rows (E, O), at most two special chunks, the plain chunk's Thomas sweep carried out symbolically in its end
values.  `plan` is the planner's table ([levels][6] = rows, chunks, plo, phi, clo, chi -- fh_large_uniform_plan or any
stand-in).  Nothing is stored per plain row on the way down; on the way up plain chunks are
x_i = alpha_i x_first + beta_i x_last + gamma_i.  Test infrastructure only."""
import numpy as np

from .recursive_partition_model import chunk_bounds

M = 16


def _reduce_chunk(rows):
    """rows: list of (a, b, c, d) -> (first-unknown row, last-unknown row); the arithmetic of reduce_level."""
    ln = len(rows)
    pa, pb, pc, pd = rows[1]
    for i in range(2, ln):
        a, b, c, d = rows[i]
        w = a / pb
        pa, pb, pc, pd = -w * pa, b - w * pc, c, d - w * pd
    qa, qb, qc, qd = rows[ln - 2]
    for i in range(ln - 3, -1, -1):
        a, b, c, d = rows[i]
        w = c / qb
        qa, qb, qc, qd = a, b - w * qa, -w * qc, d - w * qd
    return (qa, qb, qc, qd), (pa, pb, pc, pd)


def _symbolic_table(E, O):
    """alpha, beta, gamma of a plain chunk (rows E O E O ...): x_i = alpha_i x_0 + beta_i x_15 + gamma_i."""
    eo = (E, O)
    fc, fP, fG = np.zeros(M, complex), np.zeros(M, complex), np.zeros(M, complex)
    cp, Pp, Gp = 0j, 1 + 0j, 0j
    for i in range(1, M - 1):
        a, b, c, d = eo[i & 1]
        inv = 1 / (b - a * cp)
        Pp, Gp, cp = -(a * Pp) * inv, (d - a * Gp) * inv, c * inv
        fc[i], fP[i], fG[i] = cp, Pp, Gp
    al, be, ga = np.zeros(M, complex), np.zeros(M, complex), np.zeros(M, complex)
    al[0], be[M - 1] = 1, 1
    a_, b_, g_ = 0j, 1 + 0j, 0j
    for i in range(M - 2, 0, -1):
        a_, b_, g_ = fP[i] - fc[i] * a_, -(fc[i] * b_), fG[i] - fc[i] * g_
        al[i], be[i], ga[i] = a_, b_, g_
    return al, be, ga


def reduce_slab(plan, interior, row0, rowN, has_row0=True, has_rowN=True):
    """Rows of level 0: `interior` everywhere except mesh row 0 / N (when the row range holds them).  Returns
    (iface, backsub): the slab's two interface rows as complex[4][2] = (a, b, c, d) x (first, last) -- the layout of
    fh_helmholtz1d_spike_reduce_c128 -- and the function that turns the slab's two end values into its solution."""
    L = len(plan) - 1
    eo = [(tuple(interior), tuple(interior))]
    special = [dict()]  # per level: row index -> (a, b, c, d)
    n0 = int(plan[0][0])
    if has_row0:
        special[0][0] = tuple(row0)
    if has_rowN:
        special[0][n0 - 1] = tuple(rowN)

    def row(l, r):
        rows, chunks, plo, phi, clo, chi = (int(v) for v in plan[l])
        if plo <= r < phi:
            return eo[l][r & 1]
        return special[l][r]

    tables = []
    for l in range(L):
        rows, chunks, plo, phi, clo, chi = (int(v) for v in plan[l])
        bounds = chunk_bounds(rows)
        E, O = eo[l]
        first, last = _reduce_chunk([eo[l][i & 1] for i in range(M)])
        eo.append((first, last))
        tables.append(_symbolic_table(E, O))
        nxt = {}
        for c in list(range(0, clo)) + list(range(chi, chunks)):
            r0, ln = bounds[c]
            f, s = _reduce_chunk([row(l, r0 + i) for i in range(ln)])
            nxt[2 * c], nxt[2 * c + 1] = f, s
        special.append(nxt)
    top = [row(L, 0), row(L, 1)]

    def backsub(x_top):
        x = np.asarray(x_top, complex)
        for l in range(L - 1, -1, -1):
            rows, chunks, plo, phi, clo, chi = (int(v) for v in plan[l])
            bounds = chunk_bounds(rows)
            out = np.empty(rows, complex)
            al, be, ga = tables[l]
            if chi > clo:  # plain chunks: the streaming map
                c = np.arange(clo, chi)
                out[16 * clo:16 * chi] = (ga[None, :] + al[None, :] * x[2 * c, None] + be[None, :] * x[2 * c + 1, None]).ravel()
            for c in list(range(0, clo)) + list(range(chi, chunks)):
                r0, ln = bounds[c]
                x0, xl = x[2 * c], x[2 * c + 1]
                cp, dp = np.zeros(ln, complex), np.zeros(ln, complex)
                pc, pd = 0j, x0
                for i in range(1, ln - 1):
                    a, b, cc, d = row(l, r0 + i)
                    inv = 1 / (b - a * pc)
                    pd, pc = (d - a * pd) * inv, cc * inv
                    cp[i], dp[i] = pc, pd
                out[r0], out[r0 + ln - 1] = x0, xl
                xn = xl
                for i in range(ln - 2, 0, -1):
                    xn = dp[i] - cp[i] * xn
                    out[r0 + i] = xn
            x = out
        return x

    return np.array(top, complex).T.copy(), backsub


def solve(plan, interior, row0, rowN):
    """The whole mesh as one slab: reduce, solve the 2 x 2 system of the last level, back-substitute."""
    iface, backsub = reduce_slab(plan, interior, row0, rowN)
    (b0, b1), (c0, _), (_, a1), (d0, d1) = iface[1], iface[2], iface[0], iface[3]
    det = b0 * b1 - c0 * a1
    return backsub([(d0 * b1 - c0 * d1) / det, (b0 * d1 - a1 * d0) / det])
