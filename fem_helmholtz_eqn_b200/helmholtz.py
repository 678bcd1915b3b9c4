"""Problem object of the 2D annulus solvers: mesh arrays, boundary index lists and the ABC coefficient.

Mirrors ``src/helmholtz.py`` of the reference: same constructor signature, same attribute names
(``nodes, elements, n_nodes, n_elements, inner/outer_boundary_node_indices,
inner/outer_boundary_element_indices, inner_radius, outer_radius``) and the same ``get_abc_coeff``.

Mesh generation itself is outside the accelerated path and out of scope (the reference calls the gmsh CAD mesher,
helmholtz.py:31-62): the constructor produces a structured counter-clockwise annulus of
``n_r x n_theta`` quads, and ``HelmHoltz.from_arrays`` wraps a quad mesh made elsewhere (e.g. exported from gmsh).
The boundary lists are derived from the node coordinates by the reference's rules (helmholtz.py:95-118) -- on the
host with numpy, as the reference does (``topology="host"``, the default: this is mesh data, L1 in SURVEY.md), or by
the device kernels ``fh_q4_boundary_lists`` + ``fh_compact_flags_i32`` (``topology="device"``; the reference's
O(n_elements x n_boundary) Python loops are the slow step of its set-up on fine meshes).  Same lists either way.
"""
from __future__ import annotations

import numpy as np


def structured_annulus(n_r, n_theta, inner_radius=1.0, outer_radius=1.5):
    """``(nodes[n,2], elements[ne,4])`` of a ring-major structured annulus (quads counter-clockwise)."""
    r = np.linspace(inner_radius, outer_radius, n_r + 1)
    t = 2.0 * np.pi * np.arange(n_theta) / n_theta
    R, T = np.meshgrid(r, t, indexing="ij")
    nodes = np.stack([R.ravel() * np.cos(T.ravel()), R.ravel() * np.sin(T.ravel())], axis=1)
    i, j = np.meshgrid(np.arange(n_r), np.arange(n_theta), indexing="ij")
    i, j = i.ravel(), j.ravel()
    jp = (j + 1) % n_theta
    elements = np.stack([i * n_theta + j, (i + 1) * n_theta + j, (i + 1) * n_theta + jp, i * n_theta + jp],
                        axis=1).astype(np.int64)
    return nodes, elements


class HelmHoltz:
    def __init__(self, inner_radius: float = 1.0, outer_radius: float = 1.5, n_theta: int = 15, n_r: int = 14,
                 abc_order: int = 3, mesher: str = "structured", topology: str = "host") -> None:
        self.inner_radius = inner_radius
        self.outer_radius = outer_radius
        self.abc_order = abc_order
        self.n_theta = n_theta
        self.n_r = n_r
        self.n_theta_nodes = 2 * self.n_theta + 1
        self.n_r_nodes = 2 * self.n_r + 1
        self.mesher = mesher
        self.topology = topology
        self.generate_mesh()
        self.u_real = np.zeros(self.n_nodes)
        self.u_imag = np.zeros(self.n_nodes)

    @classmethod
    def from_arrays(cls, nodes, elements, inner_radius=1.0, outer_radius=1.5, abc_order=3, topology="host"):
        """Wrap an existing quad mesh (e.g. one exported from gmsh elsewhere)."""
        self = cls.__new__(cls)
        self.inner_radius, self.outer_radius, self.abc_order = inner_radius, outer_radius, abc_order
        self.n_theta = self.n_r = None
        self.topology = topology
        self._set_mesh(np.asarray(nodes, dtype=np.float64), np.asarray(elements))
        self.u_real = np.zeros(self.n_nodes)
        self.u_imag = np.zeros(self.n_nodes)
        return self

    def generate_mesh(self):
        """helmholtz.py:31-120 without the CAD mesher: a structured annulus with the reference's mesh-size formula."""
        if self.mesher not in ("structured", "auto"):
            raise NotImplementedError("mesh generation with gmsh is out of scope (SURVEY.md section 2 #8): mesh elsewhere "
                                      "and wrap the arrays with HelmHoltz.from_arrays(nodes, elements, ...)")
        self.mesh_size = min((self.outer_radius - self.inner_radius) / self.n_r,
                             (2 * np.pi * self.inner_radius) / self.n_theta,
                             (2 * np.pi * self.outer_radius) / self.n_theta)
        nodes, elements = structured_annulus(self.n_r, self.n_theta, self.inner_radius, self.outer_radius)
        self._set_mesh(nodes, elements)

    def _set_mesh(self, nodes, elements):
        self.nodes = nodes
        self.elements = elements
        self.n_nodes = len(nodes)
        self.n_elements = len(elements)
        if self.topology not in ("host", "device"):
            raise ValueError("topology must be 'host' or 'device'")
        if self.topology == "device":
            lists = [self._boundary_lists_device(r) for r in (self.inner_radius, self.outer_radius)]
            (self.inner_boundary_node_indices, inner_el), (self.outer_boundary_node_indices, outer_el) = lists
        else:
            dist = np.linalg.norm(self.nodes, axis=1)                       # helmholtz.py:95
            self.inner_boundary_node_indices = np.where(np.isclose(dist, self.inner_radius, atol=1e-5))[0]
            self.outer_boundary_node_indices = np.where(np.isclose(dist, self.outer_radius, atol=1e-5))[0]
            # helmholtz.py:114-118: elements touching a boundary node (vectorised; same lists, ascending)
            inner_el = np.nonzero(np.isin(elements, self.inner_boundary_node_indices).any(axis=1))[0]
            outer_el = np.nonzero(np.isin(elements, self.outer_boundary_node_indices).any(axis=1))[0]
        self.inner_boundary_nodes = self.nodes[self.inner_boundary_node_indices]
        self.outer_boundary_nodes = self.nodes[self.outer_boundary_node_indices]
        self.inner_boundary_element_indices = inner_el.tolist()
        self.outer_boundary_element_indices = outer_el.tolist()

    def _boundary_lists_device(self, radius):
        """helmholtz.py:95-118 for one circle on the device: ``(node indices, element indices)`` as int64 numpy arrays."""
        import torch

        from . import _lib
        from ._lib import check, ptr, stream_ptr
        lib = _lib.require_cuda()
        nodes_d = _lib.as_device(self.nodes, torch.float64)
        elems_d = _lib.as_device(np.asarray(self.elements).astype(np.int32), torch.int32)
        n, ne = self.n_nodes, self.n_elements
        nf = torch.empty(max(n, 1), dtype=torch.int32, device=nodes_d.device)
        ef = torch.empty(max(ne, 1), dtype=torch.int32, device=nodes_d.device)
        check(lib.fh_q4_boundary_lists(ptr(nodes_d), n, ptr(elems_d), ne, float(radius), 1e-5, ptr(nf), ptr(ef),
                                       stream_ptr()), "fh_q4_boundary_lists")
        out = []
        for flags, size in ((nf, n), (ef, ne)):
            lst = torch.empty(max(size, 1), dtype=torch.int32, device=flags.device)
            cnt = torch.zeros(1, dtype=torch.int32, device=flags.device)
            check(lib.fh_compact_flags_i32(ptr(flags), size, ptr(lst), ptr(cnt), stream_ptr()), "fh_compact_flags_i32")
            out.append(lst[:int(cnt.item())].cpu().numpy().astype(np.int64))
        return out[0], out[1]

    def get_abc_coeff(self, k_squared, order: int) -> complex:
        """helmholtz.py:122-153.  The reference's O(n_outer^2) loops for order 2 and 3 can only ever assign one
        constant (every node pairs with itself), so it is returned directly."""
        k = np.sqrt(k_squared)
        R = self.outer_radius
        if order == 0:
            return 1j * k
        if order == 1:
            return 1j * k - 1.0 / (2.0 * R)
        if order == 2:
            return 1j * k - (1.0 / (2.0 * R)) + (1j / (8 * k * R**2))
        if order == 3:
            return 1j * k - (1.0 / (2.0 * R)) - (1 / (8 * k * R**2)) + (1j / (8 * k**2 * R**3))
        raise ValueError("Invali d order for ABC condition. Enter 1, 2 or 3.")
