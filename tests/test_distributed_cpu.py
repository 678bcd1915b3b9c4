"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in fem_helmholtz_eqn_b200/distributed.py:
slab / k-slice bookkeeping and the collective wiring.  The device steps are replaced by the numpy model of
the kernels (tests/models/recursive_partition_model.py) -- the CUDA kernels themselves are covered by the
-m gpu tests, which emulate the ranks on one GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import banded1d

fh_dist = pytest.importorskip("fem_helmholtz_eqn_b200.distributed")


class NumpySlabOps:
    """Stand-in for CudaSlabOps: same three steps, computed with the numpy model of the kernels."""

    def __init__(self, N, k):
        self.bands = banded1d.assemble_bands(N, k, 1.0 / N)
        self.bad = False

    def reduce(self, begin, end):
        from models.recursive_partition_model import spike_reduce
        dl, d, du, b = (x[begin:end] for x in self.bands)
        n = len(self.bands[1])
        self.levels, iface = spike_reduce(dl, d, du, b, begin > 0, end < n)
        return torch.from_numpy(iface)

    def solve_interface(self, iface_all):
        from models.recursive_partition_model import spike_solve_interface
        return torch.from_numpy(spike_solve_interface([q.numpy().copy() for q in iface_all]))

    def backsub(self, x_first_last):
        from models.recursive_partition_model import spike_backsub
        return torch.from_numpy(spike_backsub(self.levels, x_first_last.numpy()))

    def singular(self):
        return self.bad


def _numpy_sweep(N, ks, L=1.0, theta=0.0, **kw):
    U = np.stack([banded1d.solve_bands(*banded1d.assemble_bands(N, float(k), L / N, theta)) for k in ks])
    return np.linspace(0, L, N + 1), torch.from_numpy(U)


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, k = 5000, 37.5
        begin, end, u_loc = fh_dist.solve_partitioned(N, k, ops=NumpySlabOps(N, k))
        _, _, u_all = fh_dist.solve_partitioned(N, k, ops=NumpySlabOps(N, k), gather=True)
        ks = np.linspace(1.0, 50.0, 7)
        lo, hi, U_loc = fh_dist.sweep_sharded(64, ks, solver=_numpy_sweep)
        _, n_k, U_all = fh_dist.sweep_sharded(64, ks, solver=_numpy_sweep, gather=True)
        r_c, w_c, U_cyc = fh_dist.sweep_sharded(64, ks, solver=_numpy_sweep, layout="cyclic")
        _, _, U_cyc_all = fh_dist.sweep_sharded(64, ks, solver=_numpy_sweep, layout="cyclic", gather=True)
        results[rank] = dict(begin=begin, end=end, u_loc=u_loc.numpy(), u_all=u_all.numpy(), lo=lo, hi=hi,
                             U_loc=U_loc.numpy(), U_all=U_all.numpy(), n_k=n_k, cyc=(r_c, w_c),
                             U_cyc=U_cyc.numpy(), U_cyc_all=U_cyc_all.numpy())
    finally:
        dist.destroy_process_group()


def _worker_bad_rank(rank, world, port, results):
    """A breakdown on ONE rank: every rank must raise (nobody may be left waiting in the gather)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, k = 5000, 37.5
        ops = NumpySlabOps(N, k)
        ops.bad = rank == 1
        try:
            fh_dist.solve_partitioned(N, k, ops=ops, gather=True)
            results[rank] = "returned"
        except np.linalg.LinAlgError:
            results[rank] = "raised"

        def sweep(N_, ks_, L_=1.0, theta_=0.0, **kw):
            nodes, U = _numpy_sweep(N_, ks_, L_, theta_)
            st = torch.zeros(len(ks_), dtype=torch.int32)
            if rank == 0:
                st[1] = 1
            return nodes, U, st
        try:
            fh_dist.sweep_sharded(64, np.linspace(1.0, 50.0, 7), solver=sweep, gather=True)
            results[rank] += "+returned"
        except np.linalg.LinAlgError:
            results[rank] += "+raised"
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_slab_and_k_slices_cover_everything():
    for n, world in ((10, 3), (4097, 2), (2**27 + 1, 8), (7, 7)):
        ranges = [fh_dist.slab_range(n, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_world_size_2_partitioned_solve_and_sharded_sweep():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    N, k = 5000, 37.5
    bands = banded1d.assemble_bands(N, k, 1.0 / N)
    ref = banded1d.solve_bands(*bands)
    glued = np.concatenate([results[r]["u_loc"] for r in range(world)])
    assert [results[r]["begin"] for r in range(world)] == [0, 2501] and results[1]["end"] == N + 1
    assert np.linalg.norm(glued - ref) / np.linalg.norm(ref) < 1e-10
    for r in range(world):
        assert np.linalg.norm(results[r]["u_all"] - ref) / np.linalg.norm(ref) < 1e-10
        assert banded1d.residual_metrics(*bands, results[r]["u_all"])[0] < 1e-15
    ks = np.linspace(1.0, 50.0, 7)
    _, Uref = _numpy_sweep(64, ks)
    assert (results[0]["lo"], results[0]["hi"], results[1]["lo"], results[1]["hi"]) == (0, 4, 4, 7)
    assert np.array_equal(np.concatenate([results[0]["U_loc"], results[1]["U_loc"]]), Uref.numpy())
    for r in range(world):
        assert results[r]["n_k"] == 7 and np.array_equal(results[r]["U_all"], Uref.numpy())
        # cyclic layout: rank r owns ks[r::world]; the gathered result is back in the order of ks
        assert results[r]["cyc"] == (r, world)
        assert np.array_equal(results[r]["U_cyc"], Uref.numpy()[r::world])
        assert np.array_equal(results[r]["U_cyc_all"], Uref.numpy())


def test_world_size_2_rank_local_breakdown_raises_on_every_rank():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_bad_rank, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: "raised+raised", 1: "raised+raised"}


def test_single_process_path_needs_no_process_group():
    N, k = 1000, 12.0
    begin, end, u = fh_dist.solve_partitioned(N, k, ops=NumpySlabOps(N, k))
    ref = banded1d.solve_bands(*banded1d.assemble_bands(N, k, 1.0 / N))
    assert (begin, end) == (0, N + 1) and np.linalg.norm(u.numpy() - ref) / np.linalg.norm(ref) < 1e-10
    with pytest.raises(ValueError):
        fh_dist.solve_partitioned(1, k, ops=NumpySlabOps(4, k))
