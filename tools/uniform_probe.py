import sys, torch
sys.path.insert(0, '.')
import fem_helmholtz_eqn_b200 as fh
from fem_helmholtz_eqn_b200 import distributed as D
N, k = 2**27, 1000.0
n = N + 1
ops = D.CudaSlabOps(N, k)
b0, e0 = D.slab_range(n, 8, 0)
for _ in range(3):
    iface = ops.reduce(b0, e0)
    x = ops.solve_interface(torch.stack([iface]).contiguous())
    ops.backsub(x[0:2])
torch.cuda.synchronize()
