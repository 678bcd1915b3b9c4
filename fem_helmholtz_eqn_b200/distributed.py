"""Multi-GPU entry points (one process per GPU, torch.distributed; NCCL on GPUs).

* ``sweep_sharded``       -- wavenumber sweep: every rank solves a contiguous slice of ``ks``; there is no
                             data-path collective, only an optional final all-gather of the solutions.
* ``solve_partitioned``   -- ONE mesh split into contiguous slabs (BASELINE config 4, N = 2^27): every rank
                             reduces its slab to two interface rows with the recursive-partition kernels,
                             the 2P interface rows (128 B per rank) are all-gathered, each rank solves the
                             tiny interface system redundantly and back-substitutes its own slab.

* ``PartitionedPlan``     -- the same for many wavenumbers on one mesh: buffers allocated once, the wavenumber in
                             device memory, the whole sequence (NCCL all-gather included) captured as a CUDA graph.

The device work is injected through a small "ops" object so that the rank bookkeeping and the collective
wiring can be exercised on CPU (gloo, world_size 2) with a numpy stand-in for the kernels; the default ops
call the CUDA library and nothing else.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, fem1d
from ._lib import check, ptr, stream_ptr


def slab_range(n_rows, world, rank):
    """Contiguous, near-equal row range ``[begin, end)`` of ``rank`` (first ``n % world`` ranks get one more)."""
    base, rem = divmod(int(n_rows), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def k_slice(n_k, world, rank):
    return slab_range(n_k, world, rank)


def _world(group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather(t, world, group):
    """all_gather_into_tensor of a (possibly complex) tensor; complex goes over the wire as (re, im) pairs."""
    import torch.distributed as dist
    src = torch.view_as_real(t.contiguous()) if t.is_complex() else t.contiguous()
    out = torch.empty((world * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    dist.all_gather_into_tensor(out, src, group=group)  # concatenated along dim 0 (gloo and nccl agree)
    out = out.reshape((world,) + tuple(src.shape))
    return torch.view_as_complex(out) if t.is_complex() else out


class CudaSlabOps:
    """The three device steps of the partitioned solve, through the C ABI (fused Helmholtz rows)."""

    def __init__(self, N, k, L=1.0, theta=0.0, **bc):
        self.lib = _lib.require_cuda()
        self.N, self.k, self.k2 = int(N), float(k), float(k**2)
        self.prob = fem1d._problem(N, L / N, theta, bc.get("bc_left", "neumann"), bc.get("bc_right", "sommerfeld"),
                                   bc.get("g_left", 1.0), bc.get("g_right", 0.0), None)
        self.device = _lib.device()

    def reduce(self, begin, end):
        n_rows = end - begin
        self.begin, self.n_rows = begin, n_rows
        self.ws_bytes = self.lib.fh_spike_workspace_bytes(n_rows)
        self.ws = torch.empty(max(self.ws_bytes, 1), dtype=torch.uint8, device=self.device)
        iface = torch.empty((4, 2), dtype=torch.complex128, device=self.device)
        check(self.lib.fh_helmholtz1d_spike_reduce_c128(C.byref(self.prob), self.k, self.k2, begin, n_rows,
                                                        ptr(self.ws), self.ws_bytes, ptr(iface), stream_ptr()),
              "fh_helmholtz1d_spike_reduce_c128")
        return iface

    def solve_interface(self, iface_all):
        world = iface_all.shape[0]
        x = torch.empty(2 * world, dtype=torch.complex128, device=self.device)
        self.status = torch.zeros(2, dtype=torch.int32, device=self.device)
        check(self.lib.fh_spike_solve_interface_c128(ptr(iface_all), world, ptr(x), ptr(self.status[:1]),
                                                     stream_ptr()), "fh_spike_solve_interface_c128")
        return x

    def backsub(self, x_first_last):
        u = torch.empty(self.n_rows, dtype=torch.complex128, device=self.device)
        check(self.lib.fh_helmholtz1d_spike_backsub_c128(C.byref(self.prob), self.k, self.k2, self.begin,
                                                         self.n_rows, ptr(self.ws), self.ws_bytes,
                                                         ptr(x_first_last.contiguous()), ptr(u),
                                                         ptr(self.status[1:]), stream_ptr()),
              "fh_helmholtz1d_spike_backsub_c128")
        return u

    def singular(self):
        return bool(torch.any(self.status != 0).item())


class PartitionedPlan:
    """``solve_partitioned`` for one mesh and many wavenumbers (BASELINE config 4; the reference's sweep loop
    nb:300-317 on a mesh too long for one GPU).  The buffers are allocated once, ``(k, k**2)`` live in device memory,
    and the sequence  reduce -> all-gather (128 B per rank, NCCL) -> interface solve -> back-substitution  is captured
    once as a CUDA graph: ``solve(k)`` costs one 16-byte copy and one graph launch per rank instead of a dozen host
    calls (at 2^24 rows per GPU the device work is ~0.1 ms -- less than issuing it call by call takes).
    Results are bit-identical to ``solve_partitioned(N, k)``.  Uniform meshes (the reference's only mode).

    With more than one rank the graph holds NCCL kernels: it is opt-in there (``use_graph=True``; measured on 2 x B200:
    0.32 ms against 0.35 ms call by call at N = 2^27), and :meth:`close` must run before
    ``dist.destroy_process_group()`` -- tearing the communicator down under a live graph blocks."""

    def __init__(self, N, L=1.0, theta=0.0, *, group=None, use_graph=None, **bc):
        self.lib = _lib.require_cuda()
        self.N, self.n, self.L, self.group = int(N), int(N) + 1, float(L), group
        self.world, self.rank = _world(group)
        if self.n < 3 * self.world:
            raise ValueError(f"mesh of {self.n} nodes is too small to split over {self.world} ranks")
        dev = self.device = _lib.device()
        self.prob = fem1d._problem(N, L / N, theta, bc.get("bc_left", "neumann"), bc.get("bc_right", "sommerfeld"),
                                   bc.get("g_left", 1.0), bc.get("g_right", 0.0), None)
        self.begin, self.end = slab_range(self.n, self.world, self.rank)
        self.n_rows = self.end - self.begin
        self.k_dev = torch.zeros(2, dtype=torch.float64, device=dev)
        self.ws_bytes = self.lib.fh_spike_workspace_bytes(self.n_rows)
        self.ws = torch.empty(max(self.ws_bytes, 1), dtype=torch.uint8, device=dev)
        self.iface = torch.empty((4, 2, 2), dtype=torch.float64, device=dev)               # complex128[4][2] as (re, im)
        self.iface_all = torch.empty((self.world, 4, 2, 2), dtype=torch.float64, device=dev)
        self.x = torch.empty(2 * self.world, dtype=torch.complex128, device=dev)
        self.status = torch.zeros(2, dtype=torch.int32, device=dev)
        self.u = torch.empty(self.n_rows, dtype=torch.complex128, device=dev)
        self.graph = None
        if use_graph is None:                # default: graph on one rank; with a collective inside it is opt-in
            use_graph = self.world == 1
        self._set_k(1.0)
        self._launch()                       # warm-up: module load, NCCL communicator, driver entry points
        torch.cuda.synchronize()
        if use_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._launch()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._launch()
                self.graph = graph
            except RuntimeError:             # a collective that cannot be captured: keep the call-by-call path
                torch.cuda.synchronize()
                self.graph = None

    def _set_k(self, k):
        # nb:185: k**2 evaluated on the host; a pageable copy is staged before the call returns (see LargeMeshPlan)
        self.k_dev.copy_(torch.tensor([float(k), float(k**2)], dtype=torch.float64))

    def _launch(self):
        import torch.distributed as dist
        lib, st = self.lib, stream_ptr()
        self.status.zero_()
        check(lib.fh_helmholtz1d_spike_reduce_kdev_c128(C.byref(self.prob), ptr(self.k_dev), self.begin, self.n_rows,
                                                        ptr(self.ws), self.ws_bytes, ptr(self.iface), st),
              "fh_helmholtz1d_spike_reduce_kdev_c128")
        if self.world > 1:
            dist.all_gather_into_tensor(self.iface_all, self.iface, group=self.group)
        else:
            self.iface_all.copy_(self.iface.unsqueeze(0))
        check(lib.fh_spike_solve_interface_c128(ptr(self.iface_all), self.world, ptr(self.x), ptr(self.status[:1]), st),
              "fh_spike_solve_interface_c128")
        check(lib.fh_helmholtz1d_spike_backsub_kdev_c128(C.byref(self.prob), ptr(self.k_dev), self.begin, self.n_rows,
                                                         ptr(self.ws), self.ws_bytes, ptr(self.x[2 * self.rank:]),
                                                         ptr(self.u), ptr(self.status[1:]), st),
              "fh_helmholtz1d_spike_backsub_kdev_c128")

    def solve(self, k):
        """Solve for wavenumber ``k``; returns ``(begin, end, u_local)`` -- this rank's node range and its slab of the
        solution (a device tensor, valid until the next call).  Stream-ordered; :meth:`singular` reads the status."""
        self._set_k(k)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch()
        return self.begin, self.end, self.u

    def singular(self):
        return bool(torch.any(self.status != 0).item())

    def close(self):
        """Drop the captured graph (required before the process group is destroyed when it holds a collective)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


def solve_partitioned(N, k, L=1.0, theta=0.0, *, group=None, ops=None, gather=False, **bc):
    """Solve ONE 1D Helmholtz problem with ``N`` elements across all ranks of ``group``.

    Returns ``(begin, end, u_local)``: this rank's node range and its slab of the solution (a tensor on the
    ops' device); with ``gather=True`` returns ``(0, N+1, u)`` with the full solution on every rank."""
    import torch.distributed as dist
    world, rank = _world(group)
    n = int(N) + 1
    if n < 3 * world:
        raise ValueError(f"mesh of {n} nodes is too small to split over {world} ranks")
    ops = ops if ops is not None else CudaSlabOps(N, k, L, theta, **bc)
    begin, end = slab_range(n, world, rank)
    iface = ops.reduce(begin, end)                          # [4, 2] complex128 on this rank
    if world > 1:
        iface_all = _all_gather(iface, world, group)        # 128 B per rank
    else:
        iface_all = iface.reshape((1,) + tuple(iface.shape))
    x_iface = ops.solve_interface(iface_all)                # [2 * world], identical on every rank
    u_local = ops.backsub(x_iface[2 * rank:2 * rank + 2])
    if ops.singular():
        raise np.linalg.LinAlgError("Singular matrix (zero pivot / non-finite value) in the partitioned solve")
    if not gather:
        return begin, end, u_local
    if world == 1:
        return 0, n, u_local
    sizes = [slab_range(n, world, r)[1] - slab_range(n, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=u_local.dtype, device=u_local.device)
    buf[:u_local.numel()] = u_local
    out = _all_gather(buf, world, group)
    return 0, n, torch.cat([out[r, :sizes[r]] for r in range(world)])


def sweep_sharded(N, ks, L=1.0, theta=0.0, *, group=None, gather=False, solver=None, layout="contiguous", **kw):
    """Wavenumber sweep sharded over the ranks of ``group``.

    ``layout="contiguous"``: rank r solves ``ks[k_slice(...)]`` and gets ``(lo, hi, U_local)``.
    ``layout="cyclic"``: rank r solves ``ks[r::world]`` and gets ``(r, world, U_local)`` -- the shards then see the
    same mix of wavenumbers, which balances the one data-dependent cost of the path (the selective refinement: on a
    sweep over k in [1, 1000] the top eighth of the range flags 13 % of its systems, the bottom eighth 0.2 %).
    With ``gather=True`` every rank gets ``(0, n_k, U)`` in the order of ``ks`` (the only collective of the sweep, an
    all-gather of the solutions).  ``solver(N, ks_local, L, theta, **kw) -> (nodes, U)`` defaults to
    :func:`fem1d.solve_helmholtz_sweep`."""
    world, rank = _world(group)
    ks_h = np.asarray(ks.detach().cpu().numpy() if isinstance(ks, torch.Tensor) else ks, dtype=np.float64).ravel()
    if layout not in ("contiguous", "cyclic"):
        raise ValueError("layout must be 'contiguous' or 'cyclic'")
    solver = solver if solver is not None else fem1d.solve_helmholtz_sweep
    if layout == "cyclic":
        _, U = solver(N, ks_h[rank::world], L, theta, **kw)
        sizes = [len(range(r, len(ks_h), world)) for r in range(world)]
        head = (rank, world)
    else:
        lo, hi = k_slice(len(ks_h), world, rank)
        _, U = solver(N, ks_h[lo:hi], L, theta, **kw)
        sizes = [k_slice(len(ks_h), world, r)[1] - k_slice(len(ks_h), world, r)[0] for r in range(world)]
        head = (lo, hi)
    if not gather or world == 1:
        return head + (U,) if not gather else (0, len(ks_h), U)
    n = U.shape[1]
    pad = max(sizes)
    buf = torch.zeros((pad, n), dtype=U.dtype, device=U.device)
    buf[:U.shape[0]] = U
    out = _all_gather(buf, world, group)
    if layout == "cyclic":
        full = torch.empty((len(ks_h), n), dtype=U.dtype, device=U.device)
        for r in range(world):
            full[r::world] = out[r, :sizes[r]]
        return 0, len(ks_h), full
    return 0, len(ks_h), torch.cat([out[r, :sizes[r]] for r in range(world)])
