"""``StructuredMesh`` and its curved element container (host side, built once).

Mirrors ``src/meshes/structured_mesh.jl`` (constructors :50-72 mapping, :96-121 faces/transfinite,
:148-160 coordinates; ``linear_interpolate`` :213-215, ``bilinear_mapping`` :254-266, ``trilinear_mapping``
:268-286, ``transfinite_mapping`` :289-332) and ``src/solvers/dgsem_structured/containers*.jl``:
  * calc_node_coordinates!        containers_3d.jl:37-61 / containers_2d.jl:37-57
  * calc_jacobian_matrix!         containers_3d.jl:63-123 / containers_2d.jl:60-88
  * calc_contravariant_vectors!   containers_3d.jl:125-285 (curl-invariant form) / containers_2d.jl:90-107
  * calc_inverse_jacobian!        containers_3d.jl:288-342 / containers_2d.jl:109-120
  * initialize_left_neighbor_connectivity! containers_3d.jl:344-400 / containers_2d.jl:151-190
Mappings are NumPy-vectorised callables ``mapping(xi, eta[, zeta]) -> (x, y[, z])``.
Elements are ordered like ``LinearIndices(size(mesh))`` (x fastest).
"""
from __future__ import annotations

import numpy as np


def linear_interpolate(s, left_value, right_value):
    return 0.5 * ((1 - s) * left_value + (1 + s) * right_value)


def coordinates2mapping(coordinates_min, coordinates_max):
    def mapping(*xi):
        return tuple(linear_interpolate(s, a, b) for s, a, b in zip(xi, coordinates_min, coordinates_max))
    return mapping


def _arr(v):
    return np.stack([np.asarray(c, dtype=np.float64) for c in v])


def transfinite_mapping(faces):
    """structured_mesh.jl:289-332; each face function returns the ndims coordinates of a face point."""
    if len(faces) == 4:
        def mapping(x, y):
            f = [lambda s, k=k: _arr(np.broadcast_arrays(*faces[k](s))) for k in range(4)]
            x1, x2, x3, x4 = f[0](-1.0), f[1](-1.0), f[0](1.0), f[1](1.0)
            shape = np.broadcast(x, y).shape
            bil = 0.25 * (x1.reshape(2, *[1] * len(shape)) * (1 - x) * (1 - y) + x2.reshape(2, *[1] * len(shape)) * (1 + x) * (1 - y)
                          + x3.reshape(2, *[1] * len(shape)) * (1 - x) * (1 + y) + x4.reshape(2, *[1] * len(shape)) * (1 + x) * (1 + y))
            out = (linear_interpolate(x, f[0](y), f[1](y)) + linear_interpolate(y, f[2](x), f[3](x)) - bil)
            return tuple(out)
        return mapping
    if len(faces) == 6:
        def F(k, a, b):
            a, b = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
            return _arr(np.broadcast_arrays(*[np.asarray(c, dtype=np.float64) + 0 * a for c in faces[k](a, b)]))

        def mapping(x, y, z):
            x, y, z = np.broadcast_arrays(x, y, z)
            li = linear_interpolate
            m1, p1 = -np.ones_like(x), np.ones_like(x)
            corners = [F(0, -1.0, -1.0), F(1, -1.0, -1.0), F(0, 1.0, -1.0), F(1, 1.0, -1.0),
                       F(0, -1.0, 1.0), F(1, -1.0, 1.0), F(0, 1.0, 1.0), F(1, 1.0, 1.0)]
            c = [cc.reshape(3, *[1] * x.ndim) for cc in corners]
            tri = 0.125 * (c[0] * (1 - x) * (1 - y) * (1 - z) + c[1] * (1 + x) * (1 - y) * (1 - z)
                           + c[2] * (1 - x) * (1 + y) * (1 - z) + c[3] * (1 + x) * (1 + y) * (1 - z)
                           + c[4] * (1 - x) * (1 - y) * (1 + z) + c[5] * (1 + x) * (1 - y) * (1 + z)
                           + c[6] * (1 - x) * (1 + y) * (1 + z) + c[7] * (1 + x) * (1 + y) * (1 + z))
            c_x = li(x, li(y, F(2, m1, z), F(3, m1, z)) + li(z, F(4, m1, y), F(5, m1, y)),
                     li(y, F(2, p1, z), F(3, p1, z)) + li(z, F(4, p1, y), F(5, p1, y)))
            c_y = li(y, li(x, F(0, m1, z), F(1, m1, z)) + li(z, F(4, x, m1), F(5, x, m1)),
                     li(x, F(0, p1, z), F(1, p1, z)) + li(z, F(4, x, p1), F(5, x, p1)))
            c_z = li(z, li(x, F(0, y, m1), F(1, y, m1)) + li(y, F(2, x, m1), F(3, x, m1)),
                     li(x, F(0, y, p1), F(1, y, p1)) + li(y, F(2, x, p1), F(3, x, p1)))
            corr = 0.5 * (c_x + c_y + c_z)
            out = (li(x, F(0, y, z), F(1, y, z)) + li(y, F(2, x, z), F(3, x, z)) + li(z, F(4, x, y), F(5, x, y))
                   - corr + tri)
            return tuple(out)
        return mapping
    raise ValueError("faces must have 4 (2D) or 6 (3D) entries")


class StructuredMesh:
    """``StructuredMesh(cells_per_dimension, mapping | faces | (coordinates_min, coordinates_max);
    periodicity)`` (structured_mesh.jl:50-160)."""

    def __init__(self, cells_per_dimension, mapping=None, coordinates_max=None, periodicity=False, faces=None):
        self.cells_per_dimension = tuple(int(c) for c in cells_per_dimension)
        self.ndims = len(self.cells_per_dimension)
        if faces is not None:
            mapping = transfinite_mapping(tuple(faces))
        elif coordinates_max is not None:
            mapping = coordinates2mapping(tuple(mapping), tuple(coordinates_max))
        elif isinstance(mapping, (tuple, list)) and callable(mapping[0]):
            mapping = transfinite_mapping(tuple(mapping))
        if not callable(mapping):
            raise TypeError("StructuredMesh needs a mapping function, a tuple of face functions or min/max coordinates")
        self.mapping = mapping
        if isinstance(periodicity, bool):
            periodicity = (periodicity,) * self.ndims
        self.periodicity = tuple(bool(p) for p in periodicity)

    @property
    def ncells(self):
        return int(np.prod(self.cells_per_dimension))

    def __repr__(self):
        return f"StructuredMesh{{{self.ndims}}} {self.cells_per_dimension}"


def compute_metric_terms(X, D):
    """jacobian_matrix, contravariant_vectors and inverse_jacobian from nodal coordinates
    ``X[dim, nodes.., element]`` (containers_3d.jl:63-342 / containers_2d.jl:60-120); shared by the
    StructuredMesh and P4estMesh element containers (dgsem_p4est/containers_3d.jl:9-27)."""
    nd = X.shape[0]
    shape_nodes = X.shape[1:1 + nd]
    nelem = X.shape[-1]
    # jacobian_matrix[dim, index, nodes.., element] = d x_dim / d xi_index (containers_3d.jl:63-123)
    J = np.empty((nd, nd) + tuple(shape_nodes) + (nelem,), order="F")
    for a in range(nd):
        J[:, a] = np.moveaxis(np.tensordot(D, X, axes=([1], [1 + a])), 0, 1 + a)
    if nd == 2:
        # containers_2d.jl:90-107
        Ja = np.empty_like(J)
        Ja[0, 0] = J[1, 1]
        Ja[1, 0] = -J[0, 1]
        Ja[0, 1] = -J[1, 0]
        Ja[1, 1] = J[0, 0]
        inv_jac = 1.0 / (J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0])
    else:
        # curl-invariant form (containers_3d.jl:125-285):
        # Ja[n, 1] = d/deta (0.5 (x_m dx_l/dzeta - x_l dx_m/dzeta)) - d/dzeta (0.5 (x_m dx_l/deta - x_l dx_m/deta)) ...
        def ddx(a, field):  # derivative along reference axis a of field[nodes.., elem]
            return np.moveaxis(np.tensordot(D, field, axes=([1], [a])), 0, a)
        Ja = np.empty_like(J)
        for nn in range(3):
            m = (nn + 1) % 3
            l = (nn + 2) % 3
            def term(c):  # 0.5 * (x_m * J[l, c] - x_l * J[m, c])
                return 0.5 * (X[m] * J[l, c] - X[l] * J[m, c])
            Ja[nn, 0] = ddx(1, term(2)) - ddx(2, term(1))
            Ja[nn, 1] = ddx(2, term(0)) - ddx(0, term(2))
            Ja[nn, 2] = ddx(0, term(1)) - ddx(1, term(0))
        det = (J[0, 0] * J[1, 1] * J[2, 2] + J[0, 1] * J[1, 2] * J[2, 0] + J[0, 2] * J[1, 0] * J[2, 1]
               - J[2, 0] * J[1, 1] * J[0, 2] - J[2, 1] * J[1, 2] * J[0, 0] - J[2, 2] * J[1, 0] * J[0, 1])
        inv_jac = 1.0 / det
    return J, np.asfortranarray(Ja), np.asfortranarray(inv_jac)


class StructuredElementContainer:
    pass


def init_elements_structured(mesh, basis):
    nd, n = mesh.ndims, basis.nnodes
    cells = mesh.cells_per_dimension
    nelem = mesh.ncells
    nodes, D = basis.nodes, basis.derivative_matrix
    # reference coordinates of every node: cell offset + dx/2 * node (containers_3d.jl:41-58)
    ref = []
    for d in range(nd):
        dx = 2 / cells[d]
        offs = -1 + (np.arange(cells[d])) * dx + dx / 2  # cell_x_offset
        r = offs[None, :] + dx / 2 * nodes[:, None]        # [n, cells_d]
        ref.append(r)
    # broadcast to [n (i), n (j), (n (k)), cx, cy, (cz)]
    shape_nodes = [n] * nd
    grids = []
    for d in range(nd):
        sh = [1] * (2 * nd)
        sh[d] = n
        sh[nd + d] = cells[d]
        grids.append(ref[d].reshape(sh))
    grids = np.broadcast_arrays(*grids)
    xyz = mesh.mapping(*grids)
    coords = np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), grids[0].shape) for c in xyz])
    # -> [nd, n.., nelem] with the element index x fastest: reverse the cell axes, then merge them
    perm = (0,) + tuple(range(1, nd + 1)) + tuple(range(2 * nd, nd, -1))
    coords = coords.transpose(perm).reshape((nd,) + tuple(shape_nodes) + (nelem,))
    el = StructuredElementContainer()
    el.nelements = nelem
    el.node_coordinates = np.asfortranarray(coords)
    J, Ja, inv_jac = compute_metric_terms(el.node_coordinates, D)
    el.jacobian_matrix = J
    el.contravariant_vectors = Ja   # [dim, index, nodes.., element]
    el.inverse_jacobian = inv_jac   # [nodes.., element]
    # left neighbours (containers_3d.jl:344-400): 1-based, 0 at a non-periodic domain boundary
    lin = np.arange(1, nelem + 1, dtype=np.int64).reshape(cells, order="F")
    left = np.empty((nd, nelem), dtype=np.int64, order="F")
    for d in range(nd):
        ln = np.roll(lin, 1, axis=d)
        if not mesh.periodicity[d]:
            idx = [slice(None)] * nd
            idx[d] = 0
            ln[tuple(idx)] = 0
        left[d] = ln.ravel(order="F")
    el.left_neighbors = left
    return el


def calc_normalvectors_subcell_fv(contravariant_vectors, basis):
    """``calc_normalvectors_subcell_fv!`` (dgsem_structured/containers_2d.jl / containers_3d.jl:352-486, container
    ``NormalVectorContainer{2,3}D`` :488-541): the free-stream preserving normal vectors of the subcell finite-volume
    interfaces inside every element, used by ``calcflux_fv!`` on curved meshes (dgsem_structured/dg_3d.jl:377-436).
    For direction a, interface i (between nodes i and i + 1 along a):
    n_i = Ja^a(node 1) + sum_{l <= i} sum_m w_l D[l, m] Ja^a(node m).
    Returns one array per direction, [ndims, n.. (n - 1 along a) .., nelements], Fortran-ordered."""
    Ja = np.asarray(contravariant_vectors)  # [component, index, nodes.., element]
    nd = Ja.shape[0]
    n = basis.nnodes
    w, D = basis.weights, basis.derivative_matrix
    out = []
    for a in range(nd):
        J = np.moveaxis(Ja[:, a], 1 + a, 1)  # [component, node along a, others.., element]
        nv = np.empty((nd, n - 1) + J.shape[2:])
        cur = J[:, 0].copy()
        for i in range(n - 1):
            for m in range(n):
                cur = cur + (w[i] * D[i, m]) * J[:, m]
            nv[:, i] = cur
        out.append(np.asfortranarray(np.moveaxis(nv, 1, 1 + a)))
    return out
