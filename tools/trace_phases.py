"""Phase lengths of one iteration of the batched solver (needs a library built with -DFH_TRACE:
    make -C fem_helmholtz_eqn_b200/csrc clean && make -C fem_helmholtz_eqn_b200/csrc EXTRA=-DFH_TRACE).
Prints, per warp of the two CTAs of cluster 0, clock64 deltas between the trace points and the globaltimer skew
between the two CTAs at the exchange."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402
from fem_helmholtz_eqn_b200 import _lib  # noqa: E402

NAMES = ["wait tiles", "tile->reg + phase 1", "drain barrier + prefetch", "phase 2", "row0 sts + barrier",
         "iface row + barrier", "CR level + barrier", "PCR levels", "exchange", "back-sub", "or-barrier",
         "store + probe"]


def main():
    lib = _lib.load()
    n_k = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    buf = torch.zeros(2 * 8 * 16 * 2, dtype=torch.int64, device="cuda")
    lib.fh_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ks = torch.linspace(1.0, 1000.0, n_k, dtype=torch.float64)
    dl, d, du, b = fh.assemble_bands(4096, ks, 1.0 / 4096)
    fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
    onepass = len(sys.argv) > 3 and sys.argv[3] == "onepass"
    run = ((lambda: fh.solve_helmholtz_sweep(4096, ks, fused=True, check_singular=False)) if fused
           else (lambda: fh.assemble_and_solve(4096, ks, refine=False)) if onepass
           else (lambda: fh.solve_tridiagonal(dl, d, du, b, refine=False)))
    run()
    torch.cuda.synchronize()
    assert lib.fh_debug_set_trace(buf.data_ptr(), it) == 0
    run()
    torch.cuda.synchronize()
    lib.fh_debug_set_trace(None, 0)
    tr = buf.cpu().numpy().reshape(2, 8, 16, 2)
    for cta in range(2):
        print(f"CTA {cta} (cluster rank {cta}); columns = warps 0..7, clock64 deltas")
        for k in range(12):
            dt = tr[cta, :, k + 1, 0] - tr[cta, :, k, 0]
            print("  %-26s %s" % (NAMES[k], " ".join("%6d" % v for v in dt)))
        print("  %-26s %s" % ("(iface row compute)", " ".join("%6d" % v for v in tr[cta, :, 13, 0] - tr[cta, :, 5, 0])))
        print("  %-26s %s" % ("(tma store wait, t = 0)", " ".join("%6d" % v for v in tr[cta, :, 14, 0] - tr[cta, :, 13, 0])))
        tot = tr[cta, :, 12, 0] - tr[cta, :, 0, 0]
        print("  %-26s %s" % ("iteration", " ".join("%6d" % v for v in tot)))
    g = tr[:, 0, :, 1].astype(np.int64)
    print("globaltimer (ns) of warp 0 at PCR end / exchange end: CTA0 %d / %d, CTA1 %d / %d (relative to CTA0 PCR end)" % (
        0, g[0, 9] - g[0, 8], g[1, 8] - g[0, 8], g[1, 9] - g[0, 8]))
    print("iteration length by globaltimer: CTA0 %d ns, CTA1 %d ns" % (g[0, 12] - g[0, 0], g[1, 12] - g[1, 0]))


if __name__ == "__main__":
    main()
