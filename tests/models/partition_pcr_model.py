"""numpy model of the algorithm in csrc/fh_tridiag_batched.cu (thread-chunk partition; one cyclic-reduction step
on the interface rows, PCR over the odd ones, even rows from their own equation; 2-CTA reversed-half coupling
with the cut row of odd-length systems; head rows).  Test infrastructure: lets the CPU suite check the MATH of
the kernel (index maps, elimination order, cluster exchange) without a GPU."""
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
    chunks = min(n_loc // M, T)          # whole 8-row chunks, right-aligned: threads T - chunks .. T - 1
    head = n_loc - M * chunks            # 0..8 rows in front of them, eliminated serially ("head rows")
    pad = SLOTS - M * chunks
    A = np.zeros((T, M), complex); B = np.ones((T, M), complex)
    C = np.zeros((T, M), complex); D = np.zeros((T, M), complex)
    body = slice(head, n_loc)
    A.reshape(-1)[pad:] = a[body]; B.reshape(-1)[pad:] = b[body]
    C.reshape(-1)[pad:] = c[body]; D.reshape(-1)[pad:] = d[body]
    first_active = T - chunks
    if head == 0 and chunks > 0:
        A[first_active, 0] = 0
    if cluster == 1:
        C[T - 1, M - 1] = 0
    hc = np.zeros(head, complex); hd = np.zeros(head, complex)
    if head > 0:
        cp = dp = 0j
        for j in range(head):
            ha = 0 if j == 0 else a[j]
            hcj = 0 if (cluster == 1 and j == n_loc - 1) else c[j]   # system shorter than one chunk
            inv = 1 / (b[j] - ha * cp)
            dp = (d[j] - ha * dp) * inv
            cp = hcj * inv
            hc[j], hd[j] = cp, dp
        if chunks > 0:
            f = first_active  # the thread whose chunk follows the head rows
            B[f, 0] -= A[f, 0] * cp; D[f, 0] -= A[f, 0] * dp; A[f, 0] = 0
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
    def sh(x, k):
        out = np.zeros_like(x)
        if k > 0: out[k:] = x[:-k]
        elif k < 0: out[:k] = x[-k:]
        else: out[:] = x
        return out

    # level 0 = cyclic reduction: only the odd rows are reduced (neighbours = the even rows on either side)
    eA, eC, eD = rA[0::2], rC[0::2], rD[0::2]                 # even rows keep their own equation
    oA, oC, oD, oV = rA[1::2], rC[1::2], rD[1::2], rV[1::2]
    Am, Cm, Dm = rA[0::2], rC[0::2], rD[0::2]                 # row 2j   (left neighbour of odd row 2j+1)
    Ap, Cp, Dp = sh(rA[0::2], -1), sh(rC[0::2], -1), sh(rD[0::2], -1)   # row 2j+2 (zero row past the end)
    inv = 1 / (1 - oA * Cm - oC * Ap)
    qA, qC = -(oA * Am) * inv, -(oC * Cp) * inv
    qD, qV = (oD - oA * Dm - oC * Dp) * inv, oV * inv         # (V lives on the last row only: even rows carry none)
    s = 1
    for _ in range(levels):                                   # PCR over the T/2 odd rows
        Am, Cm, Dm, Vm = (sh(x, s) for x in (qA, qC, qD, qV))
        Ap, Cp, Dp, Vp = (sh(x, -s) for x in (qA, qC, qD, qV))
        inv = 1 / (1 - qA * Cm - qC * Ap)
        qD, qV, qA, qC = ((qD - qA * Dm - qC * Dp) * inv, (qV - qA * Vm - qC * Vp) * inv,
                          -(qA * Am) * inv, -(qC * Cp) * inv)
        s *= 2

    def finish(x_partner):
        Xo = qD - qV * x_partner                              # odd rows
        X = np.empty(T, complex)
        X[1::2] = Xo
        X[0::2] = eD - eA * np.concatenate([[0], Xo[:-1]]) - eC * Xo   # even rows: own equation
        xl = np.concatenate([[0], X[:-1]])
        x = np.empty((T, M), complex)
        for i in range(M - 1):
            x[:, i] = (D[:, i] - A[:, i] * xl - C[:, i] * X) / B[:, i]
        x[:, M - 1] = X
        out = np.empty(n_loc, complex)
        out[head:] = x.reshape(-1)[pad:]
        xn = x[first_active, 0] if chunks > 0 else 0
        for j in range(head - 1, -1, -1):
            xn = hd[j] if (chunks == 0 and j == head - 1) else hd[j] - hc[j] * xn
            out[j] = xn
        return out

    return qD[-1], qV[-1], finish


def _levels(chunks):
    """PCR levels over the odd rows among `chunks` active interface rows (host: pcr_levels)."""
    return min(7, _ceil_log2((max(chunks, 1) + 1) // 2))


def solve(dl, d, du, b):
    n = len(d)
    if n <= SLOTS + HEADMAX:
        _, _, fin = solve_half(dl, d, du, b, 1, _levels(min(n, SLOTS) // M))
        return fin(0)
    assert n <= 2 * (SLOTS + HEADMAX)
    m = n // 2                      # both halves have m rows; odd n: row m is the cut row
    cut = n & 1
    lv = _levels(min(m, SLOTS) // M)
    Y0, V0, f0 = solve_half(dl[:m], d[:m], du[:m], b[:m], 2, lv)
    r = slice(n - 1, m + cut - 1, -1)
    Y1, V1, f1 = solve_half(du[r], d[r], dl[r], b[r], 2, lv)
    out = np.empty(n, complex)
    if cut:
        xc = (b[m] - dl[m] * Y0 - du[m] * Y1) / (d[m] - dl[m] * V0 - du[m] * V1)
        out[m] = xc
        out[:m] = f0(xc)
        out[m + 1:] = f1(xc)[::-1]
    else:
        x0 = (Y0 - V0 * Y1) / (1 - V0 * V1)
        x1 = Y1 - V1 * x0
        out[:m] = f0(x1)
        out[m:] = f1(x0)[::-1]
    return out
