"""numpy model of csrc/fh_tridiag_large.cu: recursive partition with two interface rows per 16-row chunk,
single-GPU (closed ends, coarsest level <= 4112 rows solved directly) and the per-rank SPIKE pieces
(open ends, reduced down to 2 rows, gathered 2P-row interface system).  Test infrastructure only."""
import numpy as np
from scipy.linalg import lapack

M = 16
COARSE = 4112


def _gtsv(dl, d, du, b):
    if len(d) == 1:
        return b / d
    return lapack.zgtsv(dl[1:].copy(), d.copy(), du[:-1].copy(), b.copy())[3]


def chunk_bounds(n):
    """Chunks of 16 rows inside tiles of 1024 rows; a trailing 1-row chunk / tile is merged into its
    predecessor (chunk of 17 rows, tile of 1025), so every chunk has >= 2 rows (same rule as the kernel)."""
    TILE = 1024
    n_tiles = -(-n // TILE)
    if n % TILE == 1 and n > TILE:
        n_tiles -= 1
    out = []
    for tile in range(n_tiles):
        row0 = tile * TILE
        rows = n - row0 if tile == n_tiles - 1 else TILE
        nch = -(-rows // M)
        if rows % M == 1 and rows > M:
            nch -= 1
        for t in range(nch):
            r0 = row0 + t * M
            ln = (row0 + rows - r0) if t == nch - 1 else M
            out.append((r0, ln))
    return out


def reduce_level(a, b, c, d):
    n = len(b)
    bounds = chunk_bounds(n)
    P = len(bounds)
    ra = np.zeros(2 * P, complex); rb = np.ones(2 * P, complex); rc = np.zeros(2 * P, complex); rd = np.zeros(2 * P, complex)
    for t, (r0, ln) in enumerate(bounds):
        A, B, C, D = (x[r0:r0 + ln] for x in (a, b, c, d))
        assert ln >= 2
        pa, pb, pc, pd = A[1], B[1], C[1], D[1]
        for i in range(2, ln):
            w = A[i] / pb
            pa, pb, pc, pd = -w * pa, B[i] - w * pc, C[i], D[i] - w * pd
        qa, qb, qc, qd = A[ln - 2], B[ln - 2], C[ln - 2], D[ln - 2]
        for i in range(ln - 3, -1, -1):
            w = C[i] / qb
            qa, qb, qc, qd = A[i], B[i] - w * qa, -w * qc, D[i] - w * qd
        ra[2 * t], rb[2 * t], rc[2 * t], rd[2 * t] = qa, qb, qc, qd
        ra[2 * t + 1], rb[2 * t + 1], rc[2 * t + 1], rd[2 * t + 1] = pa, pb, pc, pd
    return ra, rb, rc, rd


def backsub_level(a, b, c, d, xr):
    n = len(b)
    x = np.empty(n, complex)
    for t, (r0, ln) in enumerate(chunk_bounds(n)):
        A, B, C, D = (v[r0:r0 + ln] for v in (a, b, c, d))
        x0, xl = xr[2 * t], xr[2 * t + 1]
        cp, dp = np.zeros(ln, complex), np.zeros(ln, complex)
        pc, pd = 0j, x0
        for i in range(1, ln - 1):
            inv = 1 / (B[i] - A[i] * pc)
            pd = (D[i] - A[i] * pd) * inv
            pc = C[i] * inv
            cp[i], dp[i] = pc, pd
        x[r0] = x0
        x[r0 + ln - 1] = xl
        xn = xl
        for i in range(ln - 2, 0, -1):
            xn = dp[i] - cp[i] * xn
            x[r0 + i] = xn
    return x


def _close(a, c, open_first, open_last):
    a, c = a.copy(), c.copy()
    if not open_first:
        a[0] = 0
    if not open_last:
        c[-1] = 0
    return a, c


def solve_large(dl, d, du, b, stop=COARSE):
    levels = [(*_close(dl, du, False, False), d, b)]
    levels[0] = (levels[0][0], d, levels[0][1], b)
    while len(levels[-1][1]) > stop:
        levels.append(reduce_level(*levels[-1]))
    a, bb, c, dd = levels[-1]
    x = _gtsv(a, bb, c, dd)
    for lv in reversed(levels[:-1]):
        x = backsub_level(*lv, x)
    return x


def spike_reduce(dl, d, du, b, open_first, open_last):
    a, c = _close(dl, du, open_first, open_last)
    levels = [(a, d, c, b)]
    while len(levels[-1][1]) > 2:
        levels.append(reduce_level(*levels[-1]))
    top = levels[-1]
    iface = np.stack([top[0], top[1], top[2], top[3]])  # [4][2]
    return levels, iface


def spike_solve_interface(iface_all):
    P = len(iface_all)
    a = np.concatenate([q[0] for q in iface_all]); bb = np.concatenate([q[1] for q in iface_all])
    c = np.concatenate([q[2] for q in iface_all]); dd = np.concatenate([q[3] for q in iface_all])
    a[0] = 0; c[-1] = 0
    return _gtsv(a, bb, c, dd)


def spike_backsub(levels, x_first_last):
    x = np.asarray(x_first_last)
    for lv in reversed(levels[:-1]):
        x = backsub_level(*lv, x)
    return x


def slab_ranges(n, P):
    """Contiguous, near-equal row ranges [begin, end) per rank (same rule as the product's host code)."""
    base, rem = divmod(n, P)
    out, s = [], 0
    for r in range(P):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def solve_spike(dl, d, du, b, P):
    n = len(d)
    parts = []
    for r, (s, e) in enumerate(slab_ranges(n, P)):
        parts.append(spike_reduce(dl[s:e], d[s:e], du[s:e], b[s:e], r > 0, r < P - 1))
    xi = spike_solve_interface([p[1] for p in parts])
    return np.concatenate([spike_backsub(p[0], xi[2 * r:2 * r + 2]) for r, p in enumerate(parts)])
