"""Per-level clock64 stamps of the single-CTA reduction of the periodic recursion (deep_reduce_levels) on one 2^24-row
slab of config 4 (needs a -DFH_DEEP_TRACE build: tools/build_variant.sh dtrace -DFH_DEEP_TRACE;
FH_LIB=build/dtrace/libfemhelmholtz_b200.so python tools/trace_deep.py).  Round 2: staging 500 + two 14-step reductions
1,900-2,000 + hand-over 250 cycles per level, 22,000 cycles for the eight levels."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fem_helmholtz_eqn_b200 import distributed as D, _lib
lib = _lib.load()
N, k = 2**27, 1000.0
ops = D.CudaSlabOps(N, k)
b0, e0 = D.slab_range(N + 1, 8, 3)
for _ in range(3):
    iface = ops.reduce(b0, e0)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
lib.fh_debug_deep_trace.argtypes = [ctypes.c_void_p]
assert lib.fh_debug_deep_trace(buf) == 0
v = np.array(buf[:], dtype=np.int64)
t0 = v[0]
for l in range(9):
    if v[3*l]:
        print("level %d: start %6d  staged %6d (+%d)  reduced %6d (+%d)" % (l, v[3*l]-t0, v[3*l+1]-t0, v[3*l+1]-v[3*l], v[3*l+2]-t0, v[3*l+2]-v[3*l+1]))
print("join wait: %d .. %d" % (v[60]-t0, v[61]-t0))
