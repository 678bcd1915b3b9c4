"""One slab of BASELINE config 4 (2^24 + 1 rows of an N = 2^27 mesh, generated rows): reduce + back-substitution through
the C ABI, a few repetitions -- the command profiles/ launch lists of the periodic recursion are taken from.

    python tools/slab_probe.py [slab_index] [n_slabs]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_helmholtz_eqn_b200 import distributed as D  # noqa: E402

N, k = 2**27, 1000.0
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
slabs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ops = D.CudaSlabOps(N, k)
b0, e0 = D.slab_range(N + 1, slabs, idx)
for _ in range(3):
    iface = ops.reduce(b0, e0)
    x = ops.solve_interface(torch.stack([iface]).contiguous())
    u = ops.backsub(x[0:2])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    iface = ops.reduce(b0, e0)
    u = ops.backsub(x[0:2])
b.record()
torch.cuda.synchronize()
print("slab %d of %d: %d rows, reduce + backsub %.1f us per repetition" % (idx, slabs, e0 - b0, a.elapsed_time(b) * 100))
