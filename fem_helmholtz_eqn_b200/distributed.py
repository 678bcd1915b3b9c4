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


def _any_rank(flag, device, world, group):
    """True on every rank if ``flag`` is true on any rank (a 4-byte all-reduce): a rank-local breakdown must not make
    one rank raise while the others walk into the next collective."""
    if world == 1:
        return bool(flag)
    import torch.distributed as dist
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def backward_error_from_parts(parts, world=1, group=None):
    """Normwise backward error ||b - A u|| / (||A||_inf ||u|| + ||b||) of a partitioned solution from the per-rank
    partial sums ``parts`` = device fp64[4] = (sum |r|^2, sum |u|^2, sum |b|^2, max row sum |A|) of
    ``fh_helmholtz1d_residual_c128``: sums added, the maximum taken over the ranks.  Collective when world > 1."""
    t = parts.clone()
    if world > 1:
        import torch.distributed as dist
        mx = t[3:4].clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        t[3] = mx[0]
    r2, u2, b2, an = (float(v) for v in t.tolist())
    return float(np.sqrt(r2) / (an * np.sqrt(u2) + np.sqrt(b2)))


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
        """Breakdown (non-finite value) on THIS rank; the small-pivot bit of the probe does not count."""
        return bool(torch.any((self.status & fem1d.STATUS_NONFINITE) != 0).item())

    def flagged(self):
        """Any status bit on THIS rank: breakdown, or the interface-row probe found a defect > 5e-14."""
        return bool(torch.any(self.status != 0).item())

    def residual_parts(self, u, u_left, u_right):
        """Partial sums of the backward error over this rank's slab (rows re-generated, 16 B per row read)."""
        ws = torch.empty(self.lib.fh_helmholtz1d_residual_workspace_bytes(), dtype=torch.uint8, device=self.device)
        out = torch.empty(4, dtype=torch.float64, device=self.device)
        check(self.lib.fh_helmholtz1d_residual_c128(C.byref(self.prob), self.k, self.k2, None, self.begin, self.n_rows,
                                                    ptr(u), ptr(u_left), ptr(u_right), ptr(out), ptr(ws), ws.numel(),
                                                    stream_ptr()), "fh_helmholtz1d_residual_c128")
        return out


def _slab_halo(x_iface, world, rank):
    """The unknowns next to rank's slab, from the interface solution (entries 2r, 2r+1 = first, last unknown of slab r):
    ``(u_left, u_right)`` as 1-element views, None at a mesh end."""
    left = x_iface[2 * rank - 1:2 * rank] if rank > 0 else None
    right = x_iface[2 * rank + 2:2 * rank + 3] if rank < world - 1 else None
    return left, right


def _spike_stored(lib, bands, rhs, n_rows, world, rank, group, ws, ws_bytes, status):
    """One partitioned solve on STORED bands (the slab's dl, d, du and a right-hand side): reduce -> all-gather ->
    interface solve -> back-substitution.  Returns ``(x_local, x_iface)``."""
    dev = rhs.device
    iface = torch.empty((4, 2), dtype=torch.complex128, device=dev)
    of, ol = int(rank > 0), int(rank < world - 1)
    check(lib.fh_spike_reduce_c128(ptr(bands[0]), ptr(bands[1]), ptr(bands[2]), ptr(rhs), n_rows, of, ol, ptr(ws),
                                   ws_bytes, ptr(iface), stream_ptr()), "fh_spike_reduce_c128")
    iface_all = _all_gather(iface, world, group) if world > 1 else iface.reshape(1, 4, 2)
    x_iface = torch.empty(2 * world, dtype=torch.complex128, device=dev)
    check(lib.fh_spike_solve_interface_c128(ptr(iface_all.contiguous()), world, ptr(x_iface), ptr(status[:1]),
                                            stream_ptr()), "fh_spike_solve_interface_c128")
    x = torch.empty(n_rows, dtype=torch.complex128, device=dev)
    check(lib.fh_spike_backsub_c128(ptr(bands[0]), ptr(bands[1]), ptr(bands[2]), ptr(rhs), n_rows, of, ol, ptr(ws),
                                    ws_bytes, ptr(x_iface[2 * rank:]), ptr(x), ptr(status[1:]), stream_ptr()),
          "fh_spike_backsub_c128")
    return x, x_iface


def refine_partitioned(prob, k, begin, n_rows, u, x_iface, world, rank, group=None):
    """One step of iterative refinement of a partitioned solution, every rank on its slab (collective): the slab's
    bands are assembled (64 B per row), r = b - A u uses the neighbours' boundary unknowns from ``x_iface``, A e = r
    goes through the stored-band SPIKE pieces, u += e.  Returns the updated ``x_iface``.  This is the safety net of the
    unpivoted generated-row path (its interface-row probe, status bit 1, asks for it); it is never on the fast path."""
    lib = _lib.require_cuda()
    dev = u.device
    bands = torch.empty((4, n_rows), dtype=torch.complex128, device=dev)
    check(lib.fh_assemble_1d_slab_c128(C.byref(prob), float(k), float(k**2), None, begin, n_rows, ptr(bands[0]),
                                       ptr(bands[1]), ptr(bands[2]), ptr(bands[3]), stream_ptr()),
          "fh_assemble_1d_slab_c128")
    r = torch.empty(n_rows, dtype=torch.complex128, device=dev)
    check(lib.fh_tridiag_residual_vector_c128(ptr(bands[0]), ptr(bands[1]), ptr(bands[2]), ptr(bands[3]), ptr(u), n_rows,
                                              1, ptr(r), stream_ptr()), "fh_tridiag_residual_vector_c128")
    left, right = _slab_halo(x_iface, world, rank)
    if left is not None:       # (the residual kernel treats the slab as a closed system: add the two halo terms)
        r[0:1] -= bands[0][0:1] * left
    if right is not None:
        r[-1:] -= bands[2][-1:] * right
    ws_bytes = lib.fh_spike_workspace_bytes(n_rows)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    e, e_iface = _spike_stored(lib, bands, r, n_rows, world, rank, group, ws, ws_bytes, status)
    u += e
    return x_iface + e_iface


class PartitionedPlan:
    """``solve_partitioned`` for one mesh and many wavenumbers (BASELINE config 4; the reference's sweep loop
    nb:300-317 on a mesh too long for one GPU).  The buffers are allocated once, ``(k, k**2)`` live in device memory,
    and the sequence  reduce -> all-gather (128 B per rank, NCCL) -> interface solve -> back-substitution  is captured
    once as a CUDA graph: ``solve(k)`` costs one 16-byte copy and one graph launch per rank instead of a dozen host
    calls (at 2^24 rows per GPU the device work is ~0.1 ms -- less than issuing it call by call takes).
    Results are bit-identical to ``solve_partitioned(N, k)``.  Uniform meshes (the reference's only mode).

    ``staged=True``: A is materialised, as ``assembly()`` does (nb:185-187) -- every rank generates, stores and reduces
    the bands of its slab in one pass (64 B per row written) and back-substitutes from the stored bands (64 read + 16
    written, + 18 for the interface levels: 162 B per row in all; with a separate assembly kernel 226): the
    "assembled + solved" figure of the metric for config 4.  Default: the rows are generated inside the solver
    (16 B per row written, the periodic recursion).

    ``exchange="peer"`` (generated rows, more than one rank): no library collective -- the single-CTA kernel that
    reduces the slab stores its two interface rows into every peer's exchange buffer (symmetric memory over NVLink),
    publishes / awaits a sequence number per peer, solves the interface system and carries on with the small
    back-substitution levels (``fh_helmholtz1d_spike_solve_peer_kdev_c128``); captured as a graph by default.  Every
    rank must call ``solve`` the same number of times.

    With ``exchange="nccl"`` and more than one rank the graph holds NCCL kernels: it is opt-in there
    (``use_graph=True``; measured on 2 x B200: 0.35 ms against 0.37 ms call by call at N = 2^27), and :meth:`close`
    must run before ``dist.destroy_process_group()`` -- tearing the communicator down under a live graph blocks.

    Accuracy: the solvers are unpivoted; their interface-row probe sets ``status`` (bit 1: defect > 5e-14, bit 0:
    breakdown).  :meth:`flags` combines it over the ranks, :meth:`backward_error` measures the normwise backward error of
    the whole solution on the device (rows re-generated, 16 B per row), :meth:`solve_checked` does both and repairs."""

    def __init__(self, N, L=1.0, theta=0.0, *, group=None, use_graph=None, staged=False, exchange="nccl", **bc):
        self.lib = _lib.require_cuda()
        self.N, self.n, self.L, self.group = int(N), int(N) + 1, float(L), group
        self.world, self.rank = _world(group)
        self.staged = bool(staged)
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        if exchange == "peer" and staged:
            raise ValueError("exchange='peer' is the generated-row path (staged=False)")
        self.exchange = exchange if self.world > 1 else "nccl"
        if self.n < 3 * self.world:
            raise ValueError(f"mesh of {self.n} nodes is too small to split over {self.world} ranks")
        dev = self.device = _lib.device()
        self.prob = fem1d._problem(N, L / N, theta, bc.get("bc_left", "neumann"), bc.get("bc_right", "sommerfeld"),
                                   bc.get("g_left", 1.0), bc.get("g_right", 0.0), None)
        self.begin, self.end = slab_range(self.n, self.world, self.rank)
        self.n_rows = self.end - self.begin
        self.k_dev = torch.zeros(2, dtype=torch.float64, device=dev)
        self.k = None
        self.ws_bytes = self.lib.fh_spike_workspace_bytes(self.n_rows)
        self.ws = torch.empty(max(self.ws_bytes, 1), dtype=torch.uint8, device=dev)
        self.bands = torch.empty((4, self.n_rows), dtype=torch.complex128, device=dev) if self.staged else None
        self.iface = torch.empty((4, 2, 2), dtype=torch.float64, device=dev)               # complex128[4][2] as (re, im)
        self.iface_all = torch.empty((self.world, 4, 2, 2), dtype=torch.float64, device=dev)
        self.x = torch.empty(2 * self.world, dtype=torch.complex128, device=dev)
        self.status = torch.zeros(2, dtype=torch.int32, device=dev)
        self.u = torch.empty(self.n_rows, dtype=torch.complex128, device=dev)
        self.res_ws = torch.empty(self.lib.fh_helmholtz1d_residual_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.res4 = torch.zeros(4, dtype=torch.float64, device=dev)
        self.graph = None
        self._feed = fem1d._WavenumberFeed(self.k_dev)
        if self.exchange == "peer":
            # one exchange buffer per rank in symmetric memory (the same allocation on every rank, every rank's mapped
            # into every process over NVLink); the kernels own it from here on
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self._peer_buf = symm_mem.empty(self.lib.fh_spike_peer_buffer_bytes(), dtype=torch.uint8, device=dev)
            self._peer_buf.zero_()
            self._peer = symm_mem.rendezvous(self._peer_buf, grp)
            self._peer_ptrs = int(self._peer.buffer_ptrs_dev)
            self.seq_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            torch.cuda.synchronize()
            dist.barrier(group=group)        # every buffer is zero before anybody's first kernel stores into it
        if use_graph is None:                # default: graph on one rank; with a collective inside it is opt-in
            use_graph = self.world == 1 or self.exchange == "peer"
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
        self.k = float(k)
        self._feed.push(k)                   # nb:185: k**2 evaluated on the host; non-blocking pinned copy

    def _launch(self):
        import torch.distributed as dist
        lib, st = self.lib, stream_ptr()
        self.status.zero_()
        of, ol = int(self.rank > 0), int(self.rank < self.world - 1)
        if self.exchange == "peer":  # reduce -> exchange over peer memory -> interface solve -> back-substitution
            check(lib.fh_helmholtz1d_spike_solve_peer_kdev_c128(
                C.byref(self.prob), ptr(self.k_dev), self.begin, self.n_rows, ptr(self.ws), self.ws_bytes,
                C.c_void_p(self._peer_ptrs), self.rank, self.world, ptr(self.seq_dev), 2.0, ptr(self.x), ptr(self.u),
                ptr(self.status), st), "fh_helmholtz1d_spike_solve_peer_kdev_c128")
            return
        if self.staged:
            # level 0 generates the slab's rows, writes them out as the bands (bit-identical to the assembly kernel's) and
            # reduces them in one pass: A is written once and read once (by the back-substitution)
            B = self.bands
            check(lib.fh_helmholtz1d_spike_reduce_store_c128(C.byref(self.prob), 0.0, 0.0, ptr(self.k_dev), self.begin,
                                                             self.n_rows, ptr(B[0]), ptr(B[1]), ptr(B[2]), ptr(B[3]),
                                                             ptr(self.ws), self.ws_bytes, ptr(self.iface), st),
                  "fh_helmholtz1d_spike_reduce_store_c128")
        else:
            check(lib.fh_helmholtz1d_spike_reduce_kdev_c128(C.byref(self.prob), ptr(self.k_dev), self.begin, self.n_rows,
                                                            ptr(self.ws), self.ws_bytes, ptr(self.iface), st),
                  "fh_helmholtz1d_spike_reduce_kdev_c128")
        if self.world > 1:
            dist.all_gather_into_tensor(self.iface_all, self.iface, group=self.group)
        else:
            self.iface_all.copy_(self.iface.unsqueeze(0))
        check(lib.fh_spike_solve_interface_c128(ptr(self.iface_all), self.world, ptr(self.x), ptr(self.status[:1]), st),
              "fh_spike_solve_interface_c128")
        if self.staged:
            B = self.bands
            check(lib.fh_spike_backsub_c128(ptr(B[0]), ptr(B[1]), ptr(B[2]), ptr(B[3]), self.n_rows, of, ol, ptr(self.ws),
                                            self.ws_bytes, ptr(self.x[2 * self.rank:]), ptr(self.u), ptr(self.status[1:]),
                                            st), "fh_spike_backsub_c128")
        else:
            check(lib.fh_helmholtz1d_spike_backsub_kdev_c128(C.byref(self.prob), ptr(self.k_dev), self.begin, self.n_rows,
                                                             ptr(self.ws), self.ws_bytes, ptr(self.x[2 * self.rank:]),
                                                             ptr(self.u), ptr(self.status[1:]), st),
                  "fh_helmholtz1d_spike_backsub_kdev_c128")

    def solve(self, k):
        """Solve for wavenumber ``k``; returns ``(begin, end, u_local)`` -- this rank's node range and its slab of the
        solution (a device tensor, valid until the next call).  Stream-ordered; :meth:`flags` reads the status."""
        self._set_k(k)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch()
        return self.begin, self.end, self.u

    def flags(self):
        """Status bits of the last solve, OR-ed over the ranks (collective; synchronises): 1 = breakdown (non-finite
        value or severe probe defect), 2 = the probe found a defect > 5e-14."""
        both = torch.stack([(self.status & 1).max(), (self.status & 2).max()]).to(torch.int32)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(both, op=dist.ReduceOp.MAX, group=self.group)
        b0, b1 = both.tolist()
        return int(b0) | int(b1)

    def singular(self):
        """True (on every rank) if the last solve broke down on any rank."""
        return bool(self.flags() & fem1d.STATUS_NONFINITE)

    def residual_parts(self):
        """Launch the residual of the last solve over this rank's slab (stream-ordered); returns device fp64[4]."""
        left, right = _slab_halo(self.x, self.world, self.rank)
        check(self.lib.fh_helmholtz1d_residual_c128(C.byref(self.prob), 0.0, 0.0, ptr(self.k_dev), self.begin, self.n_rows,
                                                    ptr(self.u), ptr(left), ptr(right), ptr(self.res4), ptr(self.res_ws),
                                                    self.res_ws.numel(), stream_ptr()), "fh_helmholtz1d_residual_c128")
        return self.res4

    def backward_error(self):
        """Normwise backward error of the last solve over the WHOLE mesh (collective; synchronises)."""
        return backward_error_from_parts(self.residual_parts(), self.world, self.group)

    def solve_checked(self, k, tol=fem1d.ACCEPT_BACKWARD_ERROR):
        """:meth:`solve`, then what keeps the unpivoted path at ``np.linalg.solve`` quality (nb:204): if the probe
        flagged the solve on any rank, one refinement step through the stored-band pieces (collective), then the
        backward error is measured; ``numpy.linalg.LinAlgError`` -- raised on every rank alike -- if it is not finite or
        above ``tol`` even so."""
        begin, end, u = self.solve(k)
        if self.flags() != 0:
            self.x.copy_(refine_partitioned(self.prob, k, self.begin, self.n_rows, self.u, self.x, self.world, self.rank,
                                            self.group))
            be = self.backward_error()
            if not (be <= tol):
                raise np.linalg.LinAlgError(f"Singular matrix: the partitioned solve could not be repaired "
                                            f"(backward error {be:.1e})")
        return begin, end, u

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
    ops' device); with ``gather=True`` returns ``(0, N+1, u)`` with the full solution on every rank.  A breakdown on
    any rank raises ``numpy.linalg.LinAlgError`` on EVERY rank (the status is combined before anybody decides), so no
    rank is left waiting in a collective; a probe flag (CUDA ops) triggers one collective refinement step."""
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
    dev = u_local.device
    if _any_rank(ops.singular(), dev, world, group):
        raise np.linalg.LinAlgError("Singular matrix (zero pivot / non-finite value) in the partitioned solve")
    if hasattr(ops, "flagged") and _any_rank(ops.flagged(), dev, world, group):
        x_iface = refine_partitioned(ops.prob, k, begin, end - begin, u_local, x_iface, world, rank, group)
        left, right = _slab_halo(x_iface, world, rank)
        be = backward_error_from_parts(ops.residual_parts(u_local, left, right), world, group)
        if not (be <= fem1d.ACCEPT_BACKWARD_ERROR):
            raise np.linalg.LinAlgError(f"Singular matrix: the partitioned solve could not be repaired "
                                        f"(backward error {be:.1e})")
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
    default = solver is None
    if default:  # a singular system on one rank must raise on every rank, not leave the others in the gather below
        def solver(N_, ks_, L_, theta_, **kw_):
            nodes_, U_, st_ = fem1d.solve_helmholtz_sweep(N_, ks_, L_, theta_, check_singular=False, return_status=True,
                                                          **kw_)
            return nodes_, U_, st_
    if layout == "cyclic":
        res = solver(N, ks_h[rank::world], L, theta, **kw)
        sizes = [len(range(r, len(ks_h), world)) for r in range(world)]
        head = (rank, world)
    else:
        lo, hi = k_slice(len(ks_h), world, rank)
        res = solver(N, ks_h[lo:hi], L, theta, **kw)
        sizes = [k_slice(len(ks_h), world, r)[1] - k_slice(len(ks_h), world, r)[0] for r in range(world)]
        head = (lo, hi)
    U = res[1]
    if len(res) > 2:  # (nodes, U, status)
        bad = bool(torch.any((res[2] & fem1d.STATUS_NONFINITE) != 0).item()) if res[2].numel() else False
        if _any_rank(bad, U.device, world, group):
            raise np.linalg.LinAlgError("Singular matrix (zero pivot / non-finite solution) in the sharded sweep"
                                        + (f": local systems {torch.nonzero(res[2] & 1).flatten()[:8].tolist()}" if bad
                                           else " (on another rank)"))
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
