"""numpy model of the algorithm in csrc/fh_tridiag_batched.cu (thread-chunk partition + PCR + 2-CTA
reversed-half coupling + head rows).  Test infrastructure: lets the CPU suite check the MATH of the
kernel (index maps, elimination order, cluster exchange) without a GPU."""
import numpy as np

T, M, SLOTS, HEADMAX = 256, 8, 2048, 8


def _ceil_log2(v):
    l = 0
    while (1 << l) < v:
        l += 1
    return l


def solve_half(a, b, c, d, cluster, levels):
    """One CTA: local rows in local order (a[0] forced 0); returns closures for the exchange."""
    n_loc = len(b)
    head = max(0, n_loc - SLOTS)
    pad = SLOTS - (n_loc - head)
    A = np.zeros((T, M), complex); B = np.ones((T, M), complex)
    C = np.zeros((T, M), complex); D = np.zeros((T, M), complex)
    body = slice(head, n_loc)
    A.reshape(-1)[pad:] = a[body]; B.reshape(-1)[pad:] = b[body]
    C.reshape(-1)[pad:] = c[body]; D.reshape(-1)[pad:] = d[body]
    first_active = pad // M
    if head == 0:
        A[first_active, pad - first_active * M] = 0
    if cluster == 1:
        C[T - 1, M - 1] = 0
    hc = np.zeros(head, complex); hd = np.zeros(head, complex)
    if head > 0:
        cp = dp = 0j
        for j in range(head):
            ha = 0 if j == 0 else a[j]
            inv = 1 / (b[j] - ha * cp)
            dp = (d[j] - ha * dp) * inv
            cp = c[j] * inv
            hc[j], hd[j] = cp, dp
        B[0, 0] -= A[0, 0] * cp; D[0, 0] -= A[0, 0] * dp; A[0, 0] = 0
    for i in range(1, M):
        w = A[:, i] / B[:, i - 1]
        B[:, i] -= w * C[:, i - 1]; D[:, i] -= w * D[:, i - 1]; A[:, i] = -w * A[:, i - 1]
    for i in range(M - 3, -1, -1):
        w = C[:, i] / B[:, i + 1]
        A[:, i] -= w * A[:, i + 1]; D[:, i] -= w * D[:, i + 1]; C[:, i] = -w * C[:, i + 1]
    A0, C0, D0 = A[:, 0] / B[:, 0], C[:, 0] / B[:, 0], D[:, 0] / B[:, 0]
    rb = B[:, M - 1].copy(); rA = A[:, M - 1].copy(); rC = np.zeros(T, complex)
    rD = D[:, M - 1].copy(); rV = np.zeros(T, complex)
    cl = C[:-1, M - 1]
    rb[:-1] -= cl * A0[1:]; rC[:-1] = -cl * C0[1:]; rD[:-1] -= cl * D0[1:]
    if cluster > 1:
        rV[-1] = C[-1, M - 1]
    rA, rC, rD, rV = rA / rb, rC / rb, rD / rb, rV / rb
    s = 1
    for _ in range(levels):
        def sh(x, k):
            out = np.zeros_like(x)
            if k > 0: out[k:] = x[:-k]
            else: out[:k] = x[-k:]
            return out
        Am, Cm, Dm, Vm = (sh(x, s) for x in (rA, rC, rD, rV))
        Ap, Cp, Dp, Vp = (sh(x, -s) for x in (rA, rC, rD, rV))
        inv = 1 / (1 - rA * Cm - rC * Ap)
        nD = (rD - rA * Dm - rC * Dp) * inv
        nV = (rV - rA * Vm - rC * Vp) * inv
        nA = -(rA * Am) * inv
        nC = -(rC * Cp) * inv
        rA, rC, rD, rV = nA, nC, nD, nV
        s *= 2

    def finish(x_partner):
        X = rD - rV * x_partner
        xl = np.concatenate([[0], X[:-1]])
        x = np.empty((T, M), complex)
        for i in range(M - 1):
            x[:, i] = (D[:, i] - A[:, i] * xl - C[:, i] * X) / B[:, i]
        x[:, M - 1] = X
        out = np.empty(n_loc, complex)
        out[head:] = x.reshape(-1)[pad:]
        xn = x[0, 0]
        for j in range(head - 1, -1, -1):
            xn = hd[j] - hc[j] * xn
            out[j] = xn
        return out

    return rD[-1], rV[-1], finish


def solve(dl, d, du, b):
    n = len(d)
    if n <= SLOTS + HEADMAX:
        chunks = -(-min(n, SLOTS) // M)
        _, _, fin = solve_half(dl, d, du, b, 1, min(8, _ceil_log2(chunks)))
        return fin(0)
    assert n <= 2 * (SLOTS + HEADMAX)
    n0 = n // 2
    n1 = n - n0
    chunks = -(-min(n1, SLOTS) // M)
    lv = min(8, _ceil_log2(chunks))
    Y0, V0, f0 = solve_half(dl[:n0], d[:n0], du[:n0], b[:n0], 2, lv)
    r = slice(n - 1, n0 - 1, -1) if n0 > 0 else slice(n - 1, None, -1)
    Y1, V1, f1 = solve_half(du[r], d[r], dl[r], b[r], 2, lv)
    x0 = (Y0 - V0 * Y1) / (1 - V0 * V1)
    x1 = Y1 - V1 * x0
    out = np.empty(n, complex)
    out[:n0] = f0(x1)
    out[n0:] = f1(x0)[::-1]
    return out
