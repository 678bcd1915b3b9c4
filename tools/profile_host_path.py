"""Dev tool: where does the host time of one small solve_helmholtz_fem call go?"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fem_helmholtz_eqn_b200 as fh
for _ in range(5):
    fh.solve_helmholtz_fem(999, 10.0, bc_left="dirichlet", bc_right="dirichlet")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    fh.solve_helmholtz_fem(999, 10.0, bc_left="dirichlet", bc_right="dirichlet")
print("ms per call", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    fh.solve_helmholtz_fem(999, 10.0, bc_left="dirichlet", bc_right="dirichlet")
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
