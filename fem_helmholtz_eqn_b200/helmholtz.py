"""Problem object of the 2D annulus solvers: mesh arrays, boundary index lists and the ABC coefficient.

Mirrors ``src/helmholtz.py`` of the reference: same constructor signature, same attribute names
(``nodes, elements, n_nodes, n_elements, inner/outer_boundary_node_indices,
inner/outer_boundary_element_indices, inner_radius, outer_radius``) and the same ``get_abc_coeff``.

Mesh generation itself is outside the accelerated path (the reference calls the gmsh CAD mesher,
helmholtz.py:31-62).  If gmsh is importable it is used the same way; otherwise -- gmsh cannot be installed in
the offline build image -- a structured counter-clockwise annulus of ``n_r x n_theta`` quads is produced.  The
boundary lists are derived from the node coordinates by the reference's rules in both cases.
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
                 abc_order: int = 3, mesher: str = "auto") -> None:
        self.inner_radius = inner_radius
        self.outer_radius = outer_radius
        self.abc_order = abc_order
        self.n_theta = n_theta
        self.n_r = n_r
        self.n_theta_nodes = 2 * self.n_theta + 1
        self.n_r_nodes = 2 * self.n_r + 1
        self.mesher = mesher
        self.generate_mesh()
        self.u_real = np.zeros(self.n_nodes)
        self.u_imag = np.zeros(self.n_nodes)

    @classmethod
    def from_arrays(cls, nodes, elements, inner_radius=1.0, outer_radius=1.5, abc_order=3):
        """Wrap an existing quad mesh (e.g. one exported from gmsh elsewhere)."""
        self = cls.__new__(cls)
        self.inner_radius, self.outer_radius, self.abc_order = inner_radius, outer_radius, abc_order
        self.n_theta = self.n_r = None
        self._set_mesh(np.asarray(nodes, dtype=np.float64), np.asarray(elements))
        self.u_real = np.zeros(self.n_nodes)
        self.u_imag = np.zeros(self.n_nodes)
        return self

    def generate_mesh(self):
        """helmholtz.py:31-120.  gmsh when available (mesher 'auto'/'gmsh'), else the structured annulus."""
        use_gmsh = False
        if self.mesher in ("auto", "gmsh"):
            try:
                import gmsh  # noqa: F401
                use_gmsh = True
            except ImportError:
                if self.mesher == "gmsh":
                    raise
        if use_gmsh:
            nodes, elements = self._gmsh_annulus()
        else:
            self.mesh_size = min((self.outer_radius - self.inner_radius) / self.n_r,
                                 (2 * np.pi * self.inner_radius) / self.n_theta,
                                 (2 * np.pi * self.outer_radius) / self.n_theta)
            nodes, elements = structured_annulus(self.n_r, self.n_theta, self.inner_radius, self.outer_radius)
        self._set_mesh(nodes, elements)

    def _gmsh_annulus(self):
        import gmsh
        gmsh.initialize()
        try:
            gmsh.model.add("annulus")
            loops = [gmsh.model.occ.addCurveLoop([gmsh.model.occ.addCircle(0, 0, 0, rad)])
                     for rad in (self.outer_radius, self.inner_radius)]
            gmsh.model.occ.addPlaneSurface(loops)
            gmsh.model.occ.synchronize()
            self.mesh_size = min((self.outer_radius - self.inner_radius) / self.n_r,
                                 (2 * np.pi * self.inner_radius) / self.n_theta,
                                 (2 * np.pi * self.outer_radius) / self.n_theta)
            gmsh.model.mesh.setSize(gmsh.model.getEntities(0), self.mesh_size)
            gmsh.model.mesh.generate(2)
            gmsh.model.mesh.recombine()
            gmsh.model.mesh.setOrder(1)
            _, coords, _ = gmsh.model.mesh.getNodes()
            nodes = np.array(coords).reshape(-1, 3)[:, :2]
            _, etags, ntags = gmsh.model.mesh.getElements(2)
            blocks = [np.array(nt).reshape(-1, len(nt) // len(et)) - 1 for et, nt in zip(etags, ntags)]
            elements = np.vstack(blocks) if blocks else np.array([])
        finally:
            gmsh.finalize()
        return nodes, elements

    def _set_mesh(self, nodes, elements):
        self.nodes = nodes
        self.elements = elements
        self.n_nodes = len(nodes)
        self.n_elements = len(elements)
        dist = np.linalg.norm(self.nodes, axis=1)                       # helmholtz.py:95
        self.inner_boundary_node_indices = np.where(np.isclose(dist, self.inner_radius, atol=1e-5))[0]
        self.outer_boundary_node_indices = np.where(np.isclose(dist, self.outer_radius, atol=1e-5))[0]
        self.inner_boundary_nodes = self.nodes[self.inner_boundary_node_indices]
        self.outer_boundary_nodes = self.nodes[self.outer_boundary_node_indices]
        # helmholtz.py:114-118: elements touching a boundary node (vectorised; same lists, ascending)
        touches_in = np.isin(elements, self.inner_boundary_node_indices).any(axis=1)
        touches_out = np.isin(elements, self.outer_boundary_node_indices).any(axis=1)
        self.inner_boundary_element_indices = np.nonzero(touches_in)[0].tolist()
        self.outer_boundary_element_indices = np.nonzero(touches_out)[0].tolist()

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
