"""Timings of the BASELINE configs that are not the headline workload (they are parity tests in the suite; this
script only records how long they take).  Single process: configs 1, 2, 5 and one slab of config 4; under torchrun
with N ranks: config 4 for real (N = 2^27 split into N slabs, one NCCL all-gather of 128 B per rank).

    python tools/bench_configs.py
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_configs.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fem_helmholtz_eqn_b200 as fh  # noqa: E402
from fem_helmholtz_eqn_b200 import distributed as D  # noqa: E402


def timed(fn, reps=10, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()          # 256 MB write: evicts L2 between repetitions
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    out = {"world": world}
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if world == 1:
        # config 1: 1,000 nodes, k = 10, Dirichlet rows, host to host through the reference-facing call
        t0 = time.perf_counter()
        for _ in range(20):
            fh.solve_helmholtz_fem(999, 10.0, bc_left="dirichlet", bc_right="dirichlet")
        out["c1_host_to_host_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        # config 2 / 5: N = 1e6, staged (bands stored, 152 MB) and fused; device-resident, L2 flushed between reps
        for name, k in (("c2_k10", 10.0), ("c5_k1000", 1000.0)):
            N = 1_000_000
            dl, d, du, b = fh.assemble_bands(N, k, 1.0 / N)
            med_a, _ = timed(lambda: fh.assemble_bands(N, k, 1.0 / N, out=torch.stack([dl, d, du, b])), flush=flush)
            med_s, min_s = timed(lambda: fh.solve_tridiagonal(dl, d, du, b, refine=False), flush=flush)
            med_f, min_f = timed(lambda: fh.solve_helmholtz_sweep(N, [k], fused=True, check_singular=False), flush=flush)
            plan = fh.fem1d.LargeMeshPlan(N)
            med_g, min_g = timed(lambda: plan.solve(k), flush=flush)
            warm_g, _ = timed(lambda: plan.solve(k), reps=20)
            assert torch.equal(plan.solve(k), fh.solve_tridiagonal(dl, d, du, b, refine=False)[0][0])
            out[name + "_graph"] = {"assemble_plus_solve_ms_cold_l2": med_g, "min_ms": min_g, "warm_l2_ms": warm_g,
                                    "dof_per_s": (N + 1) / (med_g * 1e-3)}
            del plan
            plan_f = fh.fem1d.LargeMeshPlan(N, fused=True)   # A never stored: the periodic recursion on generated rows
            med_fg, min_fg = timed(lambda: plan_f.solve(k), flush=flush)
            warm_fg, _ = timed(lambda: plan_f.solve(k), reps=20)
            out[name + "_fused_graph"] = {"solve_ms_cold_l2": med_fg, "min_ms": min_fg, "warm_l2_ms": warm_fg,
                                          "dof_per_s": (N + 1) / (med_fg * 1e-3)}
            del plan_f
            out[name] = {"assemble_ms": med_a, "solve_large_ms": med_s, "solve_large_min_ms": min_s,
                         "fused_ms": med_f, "fused_min_ms": min_f,
                         "staged_dof_per_s": (N + 1) / ((med_a + med_s) * 1e-3), "fused_dof_per_s": (N + 1) / (med_f * 1e-3)}
    # config 4: N = 2^27, k = 1000, split into `slabs` contiguous slabs (fused rows)
    N, k = 2**27, 1000.0
    slabs = world if world > 1 else 8
    n = N + 1
    if world == 1:
        ops = D.CudaSlabOps(N, k)
        b0, e0 = D.slab_range(n, slabs, 0)
        med_r, _ = timed(lambda: ops.reduce(b0, e0), reps=5)
        iface = ops.reduce(b0, e0)
        x = ops.solve_interface(torch.stack([iface]).contiguous())
        med_b, _ = timed(lambda: ops.backsub(x[0:2]), reps=5)
        out["c4_one_slab_of_8"] = {"rows": e0 - b0, "reduce_ms": med_r, "backsub_ms": med_b,
                                   "dof_per_s_per_gpu": (e0 - b0) / ((med_r + med_b) * 1e-3)}
        del ops
        plan = D.PartitionedPlan(N)                           # the whole mesh on one GPU, graph replay
        med_p, min_p = timed(lambda: plan.solve(k), reps=10)
        out["c4_whole_mesh_one_gpu_plan"] = {"rows": n, "graph": plan.graph is not None, "median_ms": med_p,
                                             "min_ms": min_p, "dof_per_s": n / (med_p * 1e-3),
                                             "write_gbs": 16 * n / (med_p * 1e-3) / 1e9}
    else:
        import torch.distributed as dist
        def run():
            D.solve_partitioned(N, k)
        run()
        torch.cuda.synchronize()
        dist.barrier()
        ts = []
        for _ in range(5):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            ts.append(float(dt.item()))
        out["c4_partitioned"] = {"ranks": world, "median_ms": float(np.median(ts)) * 1e3,
                                 "dof_per_s": n / float(np.median(ts)), "timing": "host wall clock, call by call"}
        # the same through PartitionedPlan (buffers kept, k in device memory, one CUDA graph per rank): device time by
        # CUDA events, max over ranks
        plan = D.PartitionedPlan(N, use_graph=True)
        for name, use in (("c4_partitioned_plan", plan),):
            use.solve(k)
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                dist.barrier()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                use.solve(k)
                b.record()
                torch.cuda.synchronize()
                dt = torch.tensor([a.elapsed_time(b)], device="cuda")
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                ts.append(float(dt.item()))
            out[name] = {"ranks": world, "graph": use.graph is not None, "median_ms": float(np.median(ts)),
                         "min_ms": float(np.min(ts)), "dof_per_s": n / (float(np.median(ts)) * 1e-3),
                         "write_gbs_per_gpu": 16 * n / world / (float(np.median(ts)) * 1e-3) / 1e9,
                         "singular": use.singular(), "timing": "CUDA events, max over ranks"}
        plan.close()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
