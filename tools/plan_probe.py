"""BASELINE config 4 through distributed.PartitionedPlan under torchrun: checks the plan against solve_partitioned
(same bits) and times solve(k) on the device (CUDA events, max over ranks).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/plan_probe.py [graph|eager|peer] [log2 N]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_helmholtz_eqn_b200 import distributed as D  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "eager"
N = 2 ** (int(sys.argv[2]) if len(sys.argv) > 2 else 27)
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
print("rank", rank, "start", mode, flush=True)
plan = (D.PartitionedPlan(N, exchange="peer") if mode == "peer" else D.PartitionedPlan(N, use_graph=(mode == "graph")))
print("rank", rank, "plan built, graph =", plan.graph is not None, flush=True)
for k in (1000.0, 37.5):
    b, e, u = plan.solve(k)
    b2, e2, ref = D.solve_partitioned(N, k)
    assert (b, e) == (b2, e2) and torch.equal(u, ref) and not plan.singular(), k
print("rank", rank, "plan == solve_partitioned", flush=True)
k, ts = 1000.0, []
for _ in range(10):
    dist.barrier()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    plan.solve(k)
    c.record()
    torch.cuda.synchronize()
    dt = torch.tensor([a.elapsed_time(c)], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ts.append(float(dt.item()))
if rank == 0:
    med = float(np.median(ts))
    print(json.dumps({"config": "N = 2^%d, k = 1000, %d ranks" % (int(np.log2(N)), world), "mode": mode,
                      "graph": plan.graph is not None, "median_ms": med, "min_ms": float(np.min(ts)),
                      "dof_per_s": (N + 1) / (med * 1e-3), "write_gbs_per_gpu": 16 * (N + 1) / world / (med * 1e-3) / 1e9,
                      "timing": "CUDA events around solve(k), max over ranks"}))
plan.close()  # a live graph with NCCL kernels would block the teardown
dist.destroy_process_group()
