"""Role traces of one iteration of the pipelined batched solver (fh_tridiag_pipe.cuh; needs a -DFH_TRACE build:
    tools/build_variant.sh trace -DFH_TRACE;  FH_LIB=build/trace/libfemhelmholtz_b200.so python tools/trace_pipe.py).
Per warp of the two CTAs of cluster 0: clock64 at the trace points of iteration `it`, relative to the E/B warp 0's
first stamp, and the length of the iteration (stamp 0 of iteration it + 1 is not recorded: two runs, it and it + 1).
    python tools/trace_pipe.py [n_k] [it] [staged|fused|onepass]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402
from fem_helmholtz_eqn_b200 import _lib  # noqa: E402

EB = ["E top", "tiles landed", "down sweep done", "drain+prefetch", "parked, row0 out", "FullFree passed", "iface row out",
      "B: X ready", "B: xs written (or-bar)", "B: probe done"]
IG = ["I top", "FullReady passed", "CR step done", "PCR done", "exchange done", "xi out"]
SG = ["S top", "tiles free", "tiles filled", "TMA issued", "boundary rows out"]


def main():
    lib = _lib.load()
    n_k = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    mode = sys.argv[3] if len(sys.argv) > 3 else "staged"
    lib.fh_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ks = torch.linspace(1.0, 1000.0, n_k, dtype=torch.float64)
    dl, d, du, b = fh.assemble_bands(4096, ks, 1.0 / 4096)
    run = {"fused": lambda: fh.solve_helmholtz_sweep(4096, ks, fused=True, check_singular=False),
           "onepass": lambda: fh.assemble_and_solve(4096, ks, refine=False),
           "staged": lambda: fh.solve_tridiagonal(dl, d, du, b, refine=False)}[mode]
    run()
    torch.cuda.synchronize()
    traces = []
    for i in (it, it + 1):
        buf = torch.zeros(2 * 16 * 16 * 2, dtype=torch.int64, device="cuda")
        assert lib.fh_debug_set_trace(buf.data_ptr(), i) == 0
        run()
        torch.cuda.synchronize()
        lib.fh_debug_set_trace(None, 0)
        traces.append(buf.cpu().numpy().reshape(2, 16, 16, 2))
    tr, nxt = traces
    for cta in range(2):
        t0 = tr[cta, 0, 0, 0]
        print(f"CTA {cta}: clock64 relative to E/B warp 0's first stamp of iteration {it}")
        for names, warps in ((EB, range(0, 8)), (IG, range(8, 12)), (SG, range(12, 16))):
            for k, name in enumerate(names):
                v = tr[cta, list(warps), k, 0]
                if v.any():
                    print("  %-26s %s" % (name, " ".join("%7d" % (x - t0) if x else "      -" for x in v)))
        g = tr[cta, :, 0, 1].astype(np.int64), nxt[cta, :, 0, 1].astype(np.int64)
        print("  iteration length by globaltimer (ns), per warp:", " ".join("%5d" % (b_ - a_) if a_ and b_ else "    -" for a_, b_ in zip(*g)))


if __name__ == "__main__":
    main()
