import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fem_helmholtz_eqn_b200 as fh
from fem_helmholtz_eqn_b200 import fem1d
N=4096; n_k=20000
ks=np.linspace(1,1000,n_k)
nodes,U=fem1d.solve_helmholtz_sweep_host(N,ks)
_,sums=fem1d.solve_helmholtz_sweep_host(N,ks,reduce="l2")
sample=[0,n_k//3,n_k-1, 4094,4095,4096]
ref=fem1d.l2_error_sums(ks[sample],N,nodes,U[sample])
print(sums[sample]); print(ref); print(np.abs(sums[sample]-ref)/np.abs(ref))
allref=fem1d.l2_error_sums(ks,N,nodes,U)
print("max rel diff all", np.max(np.abs(sums-allref)/np.abs(allref)))
