#!/usr/bin/env python
"""Benchmark of the Helmholtz FEM hot path (BASELINE.json: "assembled+solved DOF/s and Helmholtz
solves/s (k-sweep)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: assembly of A (kept in
HBM, bit-identical to the assembly kernel's bands) + the batched tridiagonal solve + the two safety
nets, through the pipeline `solve_helmholtz_sweep` uses by default on a uniform mesh (one pass: the
rows are generated, written out and solved in one kernel), over the wavenumber sweep of BASELINE
configs[2]: 65,536 wavenumbers x the 4,096-element mesh per GPU (k in [1, 1000], Neumann +
Sommerfeld rows, complex128).  The two-kernel staged pipeline (band assembly kernel -> batched solve
kernel, round 1's headline) is timed right after it and reported under "staged".  Multi-GPU: every rank takes its own contiguous slice of a sweep of
N x 65,536 wavenumbers (independent systems, no data-path collective) -> weak scaling.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the public
host-to-host API (pinned H2D of k, D2H of every solution) and is the number to hold against the
reference arm (`--impl reference`: the reference's dense NumPy algorithm on the host cores).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ELEM = 4096          # elements per mesh (BASELINE configs[2])
N_K_PER_GPU = 65536    # wavenumbers per GPU
K_MIN, K_MAX = 1.0, 1000.0
BYTES_ASSEMBLY = 64    # per DOF: dl, d, du, b written (scalar-h mode)
BYTES_SOLVE = 80       # per DOF: 4 bands read + u written
BYTES_ONE_PASS = 80    # per DOF: 4 bands + u written, nothing read
METRIC = "assembled+solved DOF/s (k-sweep: A assembled and kept, batched tridiagonal solve)"


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, reasons, smax, power = [], set(), None, 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                power = max(power, float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": power}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place besides tests/ and smoke() that may touch oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


_literal = None


def literal_reference():
    """The reference's OWN 1D code, if build() staged it under baseline/_ref (git-ignored, shipped with the snapshot):
    code cell 2 of notebooks/1D_soln.ipynb exec'd unmodified (matplotlib stubbed -- only the plotting helpers touch it).
    Returns its namespace, or None (then the oracle's line-by-line port oracle/ref1d.py stands in)."""
    global _literal
    if _literal is None:
        _literal = False
        path = os.path.join(ROOT, "baseline", "_ref", "notebooks", "1D_soln.ipynb")
        try:
            import types

            import numpy as np
            nb = json.load(open(path))
            code = [c for c in nb["cells"] if c["cell_type"] == "code"]
            ns = {"np": np, "plt": types.ModuleType("matplotlib.pyplot"), "__name__": "reference_nb"}
            exec(compile("".join(code[1]["source"]), "1D_soln.ipynb#cell2", "exec"), ns)
            _literal = ns
        except Exception:
            _literal = False
    return _literal or None


def cpu_reference_kind():
    return "reference" if literal_reference() is not None else "port"


def cpu_reference_step(ks_sample):
    """The reference's own algorithm (dense assembly + np.linalg.solve, nb:134-205) on a bounded sample of the sweep:
    the literal notebook code when staged, else its restatement oracle/ref1d.py.  Returns seconds."""
    ns = literal_reference()
    if ns is not None:
        solve = ns["solve_helmholtz_fem"]
    else:
        from oracle import ref1d
        solve = ref1d.solve_helmholtz_fem
    t0 = time.perf_counter()
    for k in ks_sample:
        solve(N_ELEM, float(k))
    return time.perf_counter() - t0


def cpu_banded_rate(n_systems=64):
    """Banded restatement (vectorised bands + LAPACK zgtsv, 1 thread): DOF/s."""
    import numpy as np
    from oracle import banded1d
    ks = np.linspace(K_MIN, K_MAX, n_systems)
    t0 = time.perf_counter()
    for k in ks:
        banded1d.solve_bands(*banded1d.assemble_bands(N_ELEM, float(k), 1.0 / N_ELEM))
    return n_systems * (N_ELEM + 1) / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    ks_all = np.linspace(K_MIN, K_MAX, N_K_PER_GPU)
    per_step = 1  # one dense 4097x4097 assembly+solve per step (~5 s on 8 cores): bounded sample
    pick = lambda i: ks_all[(i * 7919) % N_K_PER_GPU: (i * 7919) % N_K_PER_GPU + per_step]
    for i in range(args.warmup):
        cpu_reference_step(pick(i))
    t = 0.0
    for i in range(args.steps):
        t += cpu_reference_step(pick(args.warmup + i))
    dof = args.steps * per_step * (N_ELEM + 1)
    value = dof / t
    threads = cpu_threads()
    sample = (f"{per_step} wavenumber(s) of the sweep per step at N={N_ELEM} (dense complex128 "
              f"{N_ELEM + 1}^2 assembly + LAPACK zgesv, as nb:134-205); the full 65,536-k step would take "
              f"~{N_K_PER_GPU * t / (args.steps * per_step) / 3600:.0f} h")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "DOF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128",
        "data": "synthetic", "config": workload_config(args.gpus),
        "solves_per_s": args.steps * per_step / t,
        "cpu_baseline": {"value": value, "unit": "DOF/s", "cores": threads, "kind": cpu_reference_kind(), "sample": sample,
                         "code": ("baseline/_ref/notebooks/1D_soln.ipynb cell 2, exec'd unmodified" if literal_reference()
                                  else "oracle/ref1d.py (line-by-line port; the notebook was not staged)")},
        "e2e": {"value": value, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "banded_zgtsv_dof_per_s_1thread": cpu_banded_rate(),
        "note": "the dense figure is O(N^3) against O(N) kernels: the banded LAPACK restatement on one thread is the fairer "
                "CPU yardstick and is reported beside it",
    }))


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[2]: wavenumber sweep, 65,536 k x 4,096-element 1D mesh per GPU "
                        "(k in [1,1000], Neumann + Sommerfeld rows)",
            "n_elem": N_ELEM, "n_k_per_gpu": N_K_PER_GPU, "n_k_total": N_K_PER_GPU * n_gpus,
            "dof_per_step": N_K_PER_GPU * n_gpus * (N_ELEM + 1),
            "pipeline": "one pass (the default of solve_helmholtz_sweep): fh_helmholtz1d_sweep_assemble_solve_c128 writes the "
                        "bands of A and solves from registers, then refinement + breakdown repair from the stored bands; "
                        "the staged pair fh_assemble_1d_c128 -> fh_tridiag_solve_batched_c128 is reported under 'staged'",
            "parallelism": f"k-sharded x{n_gpus} (rank r owns ks[r::{n_gpus}]), no data-path collective",
            "l2": "outputs larger than L2 (17.2 GB of bands + 4.3 GB of solutions written per step; no flush needed)"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
BYTES_LARGE_SOLVE = 162   # stored bands, one long system: reduce 64 read; back-substitution 64 read + 16 written; +18 levels
BYTES_LARGE_ONE_PASS = 162  # A materialised by the level-0 reduction itself: 64 written; back-substitution 64 + 16; +18 levels
BYTES_LARGE_FUSED = 16    # generated rows (periodic recursion): the solution is written, nothing is read at level 0


def other_configs(world, rank, dev, dist, peak):
    """The BASELINE configs that are not the headline workload, timed after it (they are parity tests in tests/; this
    records how long they take on this box, as a fraction of the HBM roofline, with the backward error of what was timed):
    config 1 (1,000 nodes, host-to-host latency), configs 2 / 5 (N = 1e6, one GPU) through fem1d.LargeMeshPlan, and
    config 4 (N = 2^27, k = 1000) through distributed.PartitionedPlan -- one slab per rank, 128-byte NCCL all-gather --,
    rows generated and with A materialised.  With more than one rank also the STRONG-scaling wavenumber sweep (65,536
    wavenumbers over all ranks) and the final all-gather of its solutions.  CUDA events, max over ranks, median of 10."""
    import numpy as np
    import torch

    import fem_helmholtz_eqn_b200 as fh
    from fem_helmholtz_eqn_b200 import distributed as D
    from fem_helmholtz_eqn_b200 import fem1d

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, reps=10, cold=False):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if cold:
                flush.zero_()  # evicts L2
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return statistics.median(ts)

    def rate(dof, ms, bytes_per_dof, gpus=1):
        gbs = bytes_per_dof * dof / (ms * 1e-3) / 1e9 / gpus
        return {"ms": ms, "dof_per_s": dof / (ms * 1e-3), "bytes_per_dof": bytes_per_dof, "gbs_per_gpu": gbs,
                "frac_of_hbm_peak": gbs / peak}

    res = {}
    if world == 1:
        # ---- config 1: the reference's CPU-runnable case, host numbers in, host solution out --------------------
        bc1 = dict(bc_left="dirichlet", bc_right="dirichlet", g_left=1.0, g_right=0.0)
        lat = {}
        for name, kw in (("dirichlet_rows", bc1), ("notebook_neumann_sommerfeld", {})):
            fh.solve_helmholtz_fem(999, 10.0, **kw)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter()
                nodes, u = fh.solve_helmholtz_fem(999, 10.0, **kw)
                ts.append(time.perf_counter() - t0)
            A, b = fh.assembly(999, 10.0, 1.0 / 999, **kw)
            be = float(A.residual(torch.from_numpy(u).reshape(1, -1))["backward"][0])
            lat[name] = {"host_to_host_ms": 1e3 * statistics.median(ts), "backward_error": be}
        res["config1_N999_k10"] = dict(lat, api="solve_helmholtz_fem(999, 10.0): mesh, assembly, solve, D2H of u",
                                       note="152 KB of traffic: launch- and PCIe-latency bound, no roofline fraction; "
                                            "the reference takes 117 ms (8 cores, SURVEY.md section 6)")
        # ---- configs 2 / 5: one mesh of 1e6 elements ------------------------------------------------------------
        N = 1_000_000
        for name, k in (("config2_N1e6_k10", 10.0), ("config5_N1e6_k1000_robin", 1000.0)):
            staged, fused = fem1d.LargeMeshPlan(N), fem1d.LargeMeshPlan(N, fused=True)
            entry = {"rows": N + 1, "api": "fem1d.LargeMeshPlan(N).solve(k): one CUDA graph replay per wavenumber"}
            for label, plan, bpd in (("staged", staged, BYTES_LARGE_ONE_PASS), ("fused", fused, BYTES_LARGE_FUSED)):
                cold, warm = timed(lambda: plan.solve(k), cold=True), timed(lambda: plan.solve(k))
                plan.solve(k)
                entry[label] = {"cold_l2": rate(N + 1, cold, bpd), "warm_l2": rate(N + 1, warm, bpd),
                                "backward_error": plan.backward_error(), "status": int(plan.status.item())}
            entry["max_dev_from_unit_modulus"] = float((fused.solve(k).abs() - 1.0).abs().max().item())
            entry["note"] = ("staged = A materialised (written by the first reduction level of the solver: "
                             "fh_helmholtz1d_assemble_solve_large_c128) and solved: 162 MB of traffic = 25 us at the HBM "
                             "peak and mostly inside the 126 MB L2 once warm: bound by its eight dependent launches, the "
                             "fractions say how far")
            res[name] = entry
            del staged, fused
    # ---- config 4: N = 2^27, one slab per rank ----------------------------------------------------------------------
    N4, k4 = 2**27, 1000.0
    entry = {"ranks": world, "rows": N4 + 1, "api": "distributed.PartitionedPlan(N).solve(k)",
             "timing": "CUDA events around solve(k) after a barrier, max over ranks, median of 10 (one solve: includes the "
                       "skew between the ranks' starts and the launch latency); back_to_back = 20 solves queued at once, "
                       "per solve; inputs larger than L2",
             "parity": "normwise backward error of the whole solution, measured on the device (rows re-generated; "
                       "all-reduced over the ranks); bar 1e-10"}
    for label, staged, bpd in (("generated_rows", False, BYTES_LARGE_FUSED), ("assembled_A", True, BYTES_LARGE_ONE_PASS)):
        sub = {}
        plan = D.PartitionedPlan(N4, use_graph=False, staged=staged)
        ms = timed(lambda: plan.solve(k4))
        plan.solve(k4)
        sub["call_by_call"] = rate(N4 + 1, ms, bpd, world)
        sub["backward_error"] = plan.backward_error()
        sub["status_flags"] = plan.flags()
        if not staged:
            dev_mod = (plan.u.abs() - 1.0).abs().max().reshape(1).to(torch.float64)
            if dist is not None:
                dist.all_reduce(dev_mod, op=dist.ReduceOp.MAX)
            sub["max_dev_from_unit_modulus"] = float(dev_mod.item())
        del plan
        torch.cuda.empty_cache()
        # the whole sequence (NCCL all-gather included) as one CUDA graph per rank
        plan = D.PartitionedPlan(N4, use_graph=True, staged=staged)
        try:
            has = torch.tensor([1 if plan.graph is not None else 0], dtype=torch.int32, device=dev)
            if dist is not None:
                dist.all_reduce(has, op=dist.ReduceOp.MIN)
            if int(has.item()) == 1:
                ms = timed(lambda: plan.solve(k4))
                ms_pipe = timed(lambda: [plan.solve(k4) for _ in range(20)], reps=5) / 20.0
                plan.solve(k4)
                sub["graph"] = dict(rate(N4 + 1, ms, bpd, world), back_to_back_ms_per_solve=ms_pipe,
                                    back_to_back_dof_per_s=(N4 + 1) / (ms_pipe * 1e-3))
                sub["graph_backward_error"] = plan.backward_error()
        finally:
            plan.close()
            del plan
            torch.cuda.empty_cache()
        if not staged and world > 1:
            # ... and without a library collective: the exchange carried out by the reduction kernel over peer memory
            # (symmetric memory; if a box cannot provide it the entry says so and the run goes on -- the ranks agree
            # on that before anybody launches a kernel that waits for its peers)
            plan, err = None, ""
            try:
                plan = D.PartitionedPlan(N4, exchange="peer")
            except Exception as exc:  # noqa: BLE001
                err = f"{type(exc).__name__}: {exc}"[:200]
            ok = torch.tensor([1 if plan is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                try:
                    ms = timed(lambda: plan.solve(k4))
                    ms_pipe = timed(lambda: [plan.solve(k4) for _ in range(20)], reps=5) / 20.0
                    plan.solve(k4)
                    sub["peer_exchange"] = dict(rate(N4 + 1, ms, bpd, world), graph=plan.graph is not None,
                                                backward_error=plan.backward_error(), status_flags=plan.flags(),
                                                back_to_back_ms_per_solve=ms_pipe,
                                                back_to_back_dof_per_s=(N4 + 1) / (ms_pipe * 1e-3))
                finally:
                    plan.close()
            else:
                sub["peer_exchange"] = {"unavailable": err or "symmetric memory could not be set up on another rank"}
            del plan
            torch.cuda.empty_cache()
        entry[label] = sub
    res["config4_N2^27_k1000"] = entry
    # ---- strong-scaling sweep + the final gather ------------------------------------------
    if world > 1:
        n, n_k = N_ELEM + 1, N_K_PER_GPU // world
        ks = np.linspace(K_MIN, K_MAX, N_K_PER_GPU)[rank::world]
        out = torch.empty((n_k, n), dtype=torch.complex128, device=dev)

        def sweep():
            fem1d.solve_helmholtz_sweep(N_ELEM, ks, out=out, check_singular=False)

        ms = timed(sweep, reps=5)
        full = torch.empty((world * n_k, n, 2), dtype=torch.float64, device=dev)
        src = torch.view_as_real(out)
        ms_g = timed(lambda: dist.all_gather_into_tensor(full, src), reps=5)
        total_bytes = world * n_k * n * 16
        res["sweep_strong_scaling"] = {
            "n_k_total": N_K_PER_GPU, "n_k_per_rank": n_k, "ranks": world,
            "api": "fem1d.solve_helmholtz_sweep on ks[rank::world] (default pipeline: one pass, A kept), host call included",
            "ms": ms, "dof_per_s": N_K_PER_GPU * n / (ms * 1e-3),
            "gather": {"api": "dist.all_gather_into_tensor of the solution shards (every rank ends with all 4.3 GB)",
                       "ms": ms_g, "bytes_total": total_bytes,
                       "algbw_gbs": total_bytes / (ms_g * 1e-3) / 1e9,
                       "per_gpu_receive_gbs": total_bytes * (world - 1) / world / (ms_g * 1e-3) / 1e9}}
        del full, out
    return res


def run_gpu(args):
    import ctypes as C

    import numpy as np
    import torch

    import fem_helmholtz_eqn_b200 as fh
    from fem_helmholtz_eqn_b200 import fem1d
    from fem_helmholtz_eqn_b200._lib import check, ptr, stream_ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator comes up (NCCL_DEBUG=VERSION on some boxes): the
        # contract is ONE JSON line on stdout, so stdout points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    lib = fh._lib.require_cuda()
    dev = torch.device("cuda", local_rank)

    n, n_k = N_ELEM + 1, args.n_k
    if args.scaling == "strong":
        n_k = max(1, args.n_k // world)
    ks_global = np.linspace(K_MIN, K_MAX, n_k * world)
    # cyclic sharding: every rank sees the same mix of wavenumbers (a contiguous slice of the top of the range flags
    # 13 % of its systems for refinement, one of the bottom 0.2 % -- the slowest rank sets the step time)
    ks_h = np.ascontiguousarray(ks_global[rank::world])
    if os.environ.get("FH_BENCH_EMULATE_SHARD"):  # development: the wavenumbers of rank R of a W-rank job, on one GPU
        w_, r_ = (int(v) for v in os.environ["FH_BENCH_EMULATE_SHARD"].split(","))
        ks_h = np.ascontiguousarray(np.linspace(K_MIN, K_MAX, n_k * w_)[r_::w_])
    k2s_h = fem1d.squares_like_the_reference(ks_h)  # nb:185: k**2 of a scalar = libm pow, not the rounded square
    ks_d = torch.from_numpy(ks_h).to(dev)
    k2s_d = torch.from_numpy(k2s_h).to(dev)
    bands = torch.empty((4, n_k, n), dtype=torch.complex128, device=dev)
    U = torch.empty((n_k, n), dtype=torch.complex128, device=dev)
    status = torch.zeros(n_k, dtype=torch.int32, device=dev)
    prob = fem1d._problem(N_ELEM, 1.0 / N_ELEM, 0.0, "neumann", "sommerfeld", 1.0, 0.0, None)

    def assemble():
        check(lib.fh_assemble_1d_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), n_k, ptr(bands[0]), ptr(bands[1]),
                                      ptr(bands[2]), ptr(bands[3]), stream_ptr()), "assemble")

    def solve():
        check(lib.fh_tridiag_solve_batched_c128(ptr(bands[0]), ptr(bands[1]), ptr(bands[2]), ptr(bands[3]),
                                                ptr(U), n, n_k, ptr(status), None, 0, stream_ptr()), "solve")

    def fused():
        check(lib.fh_helmholtz1d_sweep_fused_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), n_k, ptr(U), ptr(status),
                                                  None, 0, stream_ptr()), "fused")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # workspaces of the two stream-ordered safety nets, allocated once (nothing is allocated inside the timed region)
    refine_ws = torch.empty(lib.fh_tridiag_refine_flagged_workspace_bytes(n, n_k, max(256, n_k // 8)),
                            dtype=torch.uint8, device=dev)
    repair_ws = torch.empty(lib.fh_tridiag_repair_flagged_workspace_bytes(n, n_k, 16), dtype=torch.uint8, device=dev)

    def safety_nets():
        # (1) systems whose unpivoted elimination met a small pivot (status bit 1, 2-3 % of the sweep) get one step of
        # iterative refinement; (2) breakdowns (status bit 0: a non-finite value, or a probe defect beyond what one
        # refinement step is trusted to repair -- about one wavenumber in 10^5) are verified on the device and re-solved
        # with partial pivoting unless the refined solution is within 1e-13.  Both are device-driven and stream-ordered:
        # no host round trip, no allocation; the counters are read after the timed region.
        fem1d.refine_flagged(bands[0], bands[1], bands[2], bands[3], U, status, sync=False, ws=refine_ws)
        fem1d.repair_flagged(bands[0], bands[1], bands[2], bands[3], U, status, sync=False, ws=repair_ws)

    # one pass that KEEPS A: rows generated, written out as the bands and solved from registers (80 B/DOF of traffic
    # instead of 144) -- the default pipeline of solve_helmholtz_sweep on a uniform mesh, and the headline step
    def one_pass():
        check(lib.fh_helmholtz1d_sweep_assemble_solve_c128(C.byref(prob), ptr(ks_d), ptr(k2s_d), n_k, ptr(bands[0]),
                                                           ptr(bands[1]), ptr(bands[2]), ptr(bands[3]), ptr(U),
                                                           ptr(status), None, 0, stream_ptr()), "one-pass")

    for _ in range(args.warmup):
        one_pass()
        safety_nets()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        t_begin.record()
        for s in range(args.steps):
            ev[s][0].record()
            one_pass()
            ev[s][1].record()
            safety_nets()
            ev[s][2].record()
        t_end.record()
        barrier()
    flagged, done = (int(v) for v in refine_ws[:8].view(torch.int32).tolist())
    assert flagged == done, "more systems flagged than the refinement workspace holds"
    repaired = fem1d.repair_counts(repair_ws)  # of the last step (every step sees the same wavenumbers)
    total_ms = t_begin.elapsed_time(t_end)
    op_ms = [e[0].elapsed_time(e[1]) for e in ev]
    ref_ms = [e[1].elapsed_time(e[2]) for e in ev]
    assert int((status & 1).sum()) == 0, "a system is singular to working precision (flag kept by the pivoted re-solve)"

    # parity inside the bench: device-side backward error of what the timed steps produced, against the bands the
    # ASSEMBLY kernel writes (the one-pass kernel's own bands are compared with them bit for bit)
    U_head = U.clone()
    band_sums = bands.view(torch.int64).sum(dim=(1, 2))  # wrap-around integer checksum of the bit patterns of each band
    res = fh.tridiagonal_residual(bands[0], bands[1], bands[2], bands[3], U)
    backward_max = float(res["backward"].max())

    # the two-kernel staged pipeline (round 1's headline): assembly kernel -> batched solve kernel -> safety nets
    for _ in range(2):
        assemble()
        solve()
        safety_nets()
    barrier()
    sev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for s in range(args.steps):
        sev[s][0].record()
        assemble()
        sev[s][1].record()
        solve()
        sev[s][2].record()
        safety_nets()
        sev[s][3].record()
    barrier()
    staged_ms = sev[0][0].elapsed_time(sev[-1][3]) / args.steps
    asm_ms = [e[0].elapsed_time(e[1]) for e in sev]
    sol_ms = [e[1].elapsed_time(e[2]) for e in sev]
    staged_ref_ms = [e[2].elapsed_time(e[3]) for e in sev]
    staged_berr = float(fh.tridiagonal_residual(bands[0], bands[1], bands[2], bands[3], U)["backward"].max())
    # same bands, same arithmetic: the two pipelines agree bit for bit (solutions compared directly, bands by checksum)
    same_as_staged = bool(torch.equal(U, U_head)) and bool(torch.equal(bands.view(torch.int64).sum(dim=(1, 2)), band_sums))
    del U_head

    # fused variant (bands never stored), reported separately
    for _ in range(2):
        fused()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        fused()
    f1.record()
    barrier()
    fused_ms = f0.elapsed_time(f1) / args.steps

    # end-to-end through the public host-to-host API
    e2e = None
    e2e_s = ceil_s = l2_s = 0.0
    if not args.no_e2e:
        del bands
        torch.cuda.empty_cache()
        plan = fem1d.HostSweepPlan(N_ELEM, n_k)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        fem1d.solve_helmholtz_sweep_host(N_ELEM, ks_h, plan=plan)  # warm-up (page-locks, first touch)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, U_host = fem1d.solve_helmholtz_sweep_host(N_ELEM, ks_h, plan=plan)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        sample = [0, n_k // 3, n_k - 1]
        sums_from_u = fem1d.l2_error_sums(ks_h[sample], N_ELEM, np.linspace(0, 1.0, n), U_host[sample])
        # the box's ceiling for that path: the same bytes, device -> the same pinned buffer, nothing else -- every rank
        # at once (one process per GPU; on this pool the host is a VM with one NUMA node: all ranks share its DMA path)
        def d2h_only():
            for c, lo in enumerate(range(0, n_k, plan.chunk)):
                hi = min(n_k, lo + plan.chunk)
                plan.U_host[lo:hi].copy_(plan.U_dev[c & 1][:hi - lo], non_blocking=True)
            torch.cuda.synchronize()
        d2h_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d2h_only()
        ceil_s = (time.perf_counter() - t0) / e2e_steps
        e2e = {"h2d": plan.h2d_bytes, "d2h": plan.d2h_bytes, "steps": e2e_steps, "pipeline": plan.pipeline}
        del plan, U_host
        torch.cuda.empty_cache()
        # the reduced variant: what the reference's own sweep driver (convergence_test, nb:284-362) takes from every
        # solution is its L2 error sum -- 16 bytes per wavenumber cross PCIe instead of 64 KB
        plan = fem1d.HostSweepPlan(N_ELEM, n_k, reduce="l2")
        fem1d.solve_helmholtz_sweep_host(N_ELEM, ks_h, plan=plan)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, sums = fem1d.solve_helmholtz_sweep_host(N_ELEM, ks_h, plan=plan)
        torch.cuda.synchronize()
        l2_s = (time.perf_counter() - t0) / e2e_steps
        e2e.update({"l2_h2d": plan.h2d_bytes, "l2_d2h": plan.d2h_bytes,
                    "l2_matches_full": bool(np.allclose(sums[sample], sums_from_u, rtol=1e-12, atol=0.0))})
        del plan

    # max over ranks
    vals = torch.tensor([total_ms, e2e_s, fused_ms, statistics.mean(sol_ms), statistics.mean(asm_ms), float(done),
                         backward_max, staged_ms, staged_berr, float(repaired["flagged"]), float(repaired["kept"]),
                         float(repaired["resolved"]), ceil_s, l2_s, statistics.mean(op_ms), statistics.mean(ref_ms),
                         statistics.mean(staged_ref_ms), 0.0 if same_as_staged else 1.0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    (total_ms, e2e_s, fused_ms, sol_mean, asm_mean, refined_max, backward_max, staged_ms, staged_berr, brk_flagged,
     brk_kept, brk_resolved, ceil_s, l2_s, op_mean, ref_mean, staged_ref_mean, differs) = vals.tolist()

    configs = None
    if not args.no_configs:  # every rank takes part (config 4 holds a collective)
        del U
        torch.cuda.empty_cache()
        configs = other_configs(world, rank, dev, dist, measured_peak_gbs()[0])

    if rank == 0:
        dof_rank = n_k * n
        dof_all = dof_rank * world
        ms_per_step = total_ms / args.steps
        value = dof_all / (ms_per_step * 1e-3)
        peak, peak_kind = measured_peak_gbs()
        achieved = BYTES_ONE_PASS * dof_rank / (op_mean * 1e-3) / 1e9
        achieved_solve = BYTES_SOLVE * dof_rank / (sol_mean * 1e-3) / 1e9
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        traffic_note = ("profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from a "
                        "committed `ncu --set full` capture of the same launch -- a constant of that capture, not "
                        "measured in this run")
        peak_note = (f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured"
                     else "fallback 6.65 TB/s (B200_PROFILING.md)")
        staged_ms_all = staged_ms
        out = {
            "metric": METRIC, "value": value, "unit": "DOF/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
            "config": dict(workload_config(world), n_k_per_gpu=n_k, n_k_total=n_k * world, dof_per_step=n_k * world * n),
            "solves_per_s": n_k * world / (ms_per_step * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "tridiag_pipe_kernel<2, HelmSourceT<true, true>> "
                                                   "(fh_helmholtz1d_sweep_assemble_solve_c128)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_note,
                         "algorithmic_bytes_per_launch": BYTES_ONE_PASS * dof_rank,
                         "algorithmic_bytes_per_dof": "64 written (dl, d, du, b) + 16 written (u), nothing read",
                         "traffic": traffic.get("one_pass_bytes_per_launch"), "traffic_source": traffic_note,
                         "launch_ms": op_mean},
            "stages": {"one_pass_kernel_ms": op_mean, "refine_and_repair_ms": ref_mean,
                       "one_pass_kernel_ms_per_step": [round(v, 3) for v in op_ms],
                       "refined_systems_per_step_max_over_ranks": int(refined_max),
                       "breakdown_flags_per_step_max_over_ranks": int(brk_flagged),
                       "breakdowns_kept_after_verification_max_over_ranks": int(brk_kept),
                       "pivoted_resolves_per_step_max_over_ranks": int(brk_resolved),
                       "step_gbs": BYTES_ONE_PASS * dof_rank / (ms_per_step * 1e-3) / 1e9,
                       "step_frac": BYTES_ONE_PASS * dof_rank / (ms_per_step * 1e-3) / 1e9 / peak},
            "staged": {"what": "the two-kernel pipeline (round 1's headline): fh_assemble_1d_c128 -> "
                               "fh_tridiag_solve_batched_c128 -> refinement + repair; same outputs, bit for bit",
                       "ms_per_step": staged_ms_all, "value": dof_all / (staged_ms_all * 1e-3), "unit": "DOF/s",
                       "bytes_per_dof": BYTES_ASSEMBLY + BYTES_SOLVE,
                       "assembly_ms": asm_mean, "solve_ms": sol_mean, "refine_and_repair_ms": staged_ref_mean,
                       "solve_ms_per_step": [round(v, 3) for v in sol_ms],
                       "assembly_gbs": BYTES_ASSEMBLY * dof_rank / (asm_mean * 1e-3) / 1e9,
                       "assembly_frac": BYTES_ASSEMBLY * dof_rank / (asm_mean * 1e-3) / 1e9 / peak,
                       "solve_roofline": {"bound": "hbm", "kernel": "tridiag_pipe_kernel<2, BandSource> "
                                                                    "(fh_tridiag_solve_batched_c128)",
                                          "achieved": achieved_solve, "peak": peak, "unit": "GB/s",
                                          "frac": achieved_solve / peak,
                                          "algorithmic_bytes_per_launch": BYTES_SOLVE * dof_rank,
                                          "traffic": traffic.get("solve_bytes_per_launch"), "launch_ms": sol_mean},
                       "staged_gbs": (BYTES_ASSEMBLY + BYTES_SOLVE) * dof_rank / (staged_ms_all * 1e-3) / 1e9,
                       "staged_frac": (BYTES_ASSEMBLY + BYTES_SOLVE) * dof_rank / (staged_ms_all * 1e-3) / 1e9 / peak,
                       "backward_error_max": staged_berr,
                       "identical_to_one_pass_outputs": not differs},
            "fused": {"ms_per_step": fused_ms, "value": dof_all / (fused_ms * 1e-3), "unit": "DOF/s",
                      "bytes_per_dof": 16, "what": "fh_helmholtz1d_sweep_fused_c128: A never stored (no safety nets: "
                                                   "nothing to refine from)"},
            "parity": {"backward_error_max": backward_max, "bar": 1e-10},
            # per step: the one-pass kernel, 2 refinement kernels (flag compaction; the one-pass refinement solve:
            # residual + correction solve + update), the repair kernel (scan + verify-or-pivot)
            "gpu_launches": 4 * args.steps,
            "clocks": clocks.summary(),
        }
        if configs is not None:
            out["other_configs"] = configs
        if e2e:
            d2h_all = e2e["d2h"] * world
            out["e2e"] = {"value": dof_all / e2e_s, "unit": "DOF/s", "h2d_bytes_per_step": e2e["h2d"],
                          "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": 1e3 * e2e_s, "steps": e2e["steps"],
                          "api": "fem1d.solve_helmholtz_sweep_host (pinned H2D of k, k^2; D2H of every solution), "
                                 f"pipeline {e2e['pipeline']}",
                          "d2h_gbs_all_gpus": d2h_all / e2e_s / 1e9,
                          "ceiling_gbs": d2h_all / ceil_s / 1e9,
                          "ceiling_ms_per_step": 1e3 * ceil_s,
                          "frac_of_ceiling": ceil_s / e2e_s,
                          "ceiling": "the same solution bytes, device -> the same pinned host buffers, nothing else, all "
                                     "ranks at once (max over ranks): what this box's PCIe / host-memory path delivers",
                          "reduced_l2": {"value": dof_all / l2_s, "unit": "DOF/s", "ms_per_step": 1e3 * l2_s,
                                         "h2d_bytes_per_step": e2e["l2_h2d"], "d2h_bytes_per_step": e2e["l2_d2h"],
                                         "matches_full_solution_path": e2e["l2_matches_full"],
                                         "api": "fem1d.solve_helmholtz_sweep_host(reduce='l2'): per wavenumber the "
                                                "16-byte sum of compute_L2_error (nb:259-281) comes back, which is all "
                                                "the reference's convergence_test (nb:284-362) takes from a solution"}}
        if world == 1 and not args.no_cpu:
            ks_cpu = ks_h[[0, n_k // 2, n_k - 1]]
            cpu_reference_step(ks_cpu[:1])  # warm the BLAS threads
            t = cpu_reference_step(ks_cpu)
            out["cpu_baseline"] = {
                "value": len(ks_cpu) * n / t, "unit": "DOF/s", "cores": cpu_threads(), "kind": cpu_reference_kind(),
                "sample": f"{len(ks_cpu)} of the {n_k} wavenumbers (k = first, middle, last) at N={N_ELEM}: the "
                          f"reference's dense assembly + np.linalg.solve ({'the notebook code itself' if literal_reference() else 'oracle/ref1d.py'}), {t:.1f} s",
                "banded_zgtsv_dof_per_s_1thread": cpu_banded_rate()}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-k", dest="n_k", type=int, default=N_K_PER_GPU, help="wavenumbers per GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): 65,536 wavenumbers PER GPU; strong: 65,536 in all, shared out over the GPUs")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the timings of BASELINE configs 2, 4 and 5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
