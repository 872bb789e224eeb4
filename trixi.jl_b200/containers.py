"""DG containers for the TreeMesh solver, built on the host once (``create_cache``
``src/solvers/dgsem_tree/dg_2d.jl:14-37``) and uploaded by ``trixi_b200_create``.

Follows ``src/solvers/dgsem_tree/containers_3d.jl`` (and ``containers_2d.jl``):
  * init_elements!    :87-133   inverse_jacobian = 2/dx, node_coordinates[dim, i, j, k, element]
  * init_interfaces!  :229-281  element-major, positive directions only; neighbor_ids[1]=left, [2]=right
  * init_boundaries!  :391-471  direction-major so boundaries are sorted -x,+x,-y,...
Indices stored 1-based int64 like Julia ``Int``; arrays are Fortran-ordered (column-major).
"""
from __future__ import annotations

import numpy as np


class ElementContainer:
    pass


class InterfaceContainer:
    pass


class BoundaryContainer:
    pass


def init_elements(mesh, basis):
    """containers_3d.jl:87-133.  Returns inverse_jacobian[nelem], node_coordinates[ndims, n^d, nelem]."""
    nodes = basis.nodes
    n = basis.nnodes
    nd = mesh.ndims
    # integrate(one, nodes, basis): sequential sum of the weights (basis_lobatto_legendre.jl:142-150)
    reference_length = 0.0
    for w in basis.weights:
        reference_length += 1.0 * w
    reference_offset = (nodes[0] + nodes[-1]) / 2
    dx = mesh.length_at_level(mesh.levels)
    jacobian = dx / reference_length
    el = ElementContainer()
    el.inverse_jacobian = 1.0 / jacobian
    centers = mesh.cell_coordinates()  # [nd, nelem]
    nelem = mesh.ncells
    el.nelements = nelem
    xi = nodes - reference_offset  # [n]
    # node_coordinates[d, i, j, k, e] = center[d, e] + jacobian[e] * xi[index along d]
    shape = (nd,) + (n,) * nd + (nelem,)
    coords = np.empty(shape, order="F")
    for d in range(nd):
        bshape = [1] * nd + [nelem]
        nshape = [1] * (nd + 1)
        nshape[d] = n
        coords[d] = centers[d].reshape(bshape) + jacobian.reshape(bshape) * xi.reshape(nshape)
    el.node_coordinates = coords
    return el


def init_interfaces(mesh):
    """containers_3d.jl:229-281 (count :195-227)."""
    nd = mesh.ndims
    nelem = mesh.ncells
    nb = np.empty((nelem, nd), dtype=np.int64)
    for d in range(nd):
        nb[:, d] = mesh._face_neighbors(2 * d + 1, "same")  # positive direction
    valid = nb >= 0
    left = np.repeat(np.arange(nelem, dtype=np.int64)[:, None], nd, axis=1)[valid]
    right = nb[valid]
    orient = np.repeat(np.arange(1, nd + 1, dtype=np.int64)[None, :], nelem, axis=0)[valid]
    ic = InterfaceContainer()
    ic.neighbor_ids = np.asfortranarray(np.stack([left + 1, right + 1]))  # [2, I], 1-based
    ic.orientations = orient
    ic.ninterfaces = orient.shape[0]
    return ic


def init_boundaries(mesh, elements, basis):
    """containers_3d.jl:391-471."""
    nd = mesh.ndims
    n = basis.nnodes
    ids, orients, sides, coords, counts = [], [], [], [], []
    for direction in range(2 * nd):
        same = mesh._face_neighbors(direction, "same")
        coarse = mesh._face_neighbors(direction, "coarse")
        isb = (same < 0) & (coarse < 0)
        # a cell with a *refined* same-level neighbour has a non-leaf neighbour cell -> not a boundary
        if hasattr(mesh, "_has_refined_neighbor"):
            isb &= ~mesh._has_refined_neighbor(direction)
        el = np.nonzero(isb)[0]
        counts.append(el.shape[0])
        ids.append(el + 1)
        orients.append(np.full(el.shape[0], direction // 2 + 1, dtype=np.int64))
        sides.append(np.full(el.shape[0], 1 if direction % 2 == 1 else 2, dtype=np.int64))
        enc = elements.node_coordinates  # [nd, n.., nelem]
        idx = [slice(None)] * (nd + 2)
        idx[1 + direction // 2] = 0 if direction % 2 == 0 else n - 1
        coords.append(enc[tuple(idx)][..., el])
    bc = BoundaryContainer()
    bc.neighbor_ids = np.concatenate(ids).astype(np.int64)
    bc.orientations = np.concatenate(orients)
    bc.neighbor_sides = np.concatenate(sides)
    bc.node_coordinates = np.asfortranarray(np.concatenate(coords, axis=-1))
    bc.n_boundaries_per_direction = np.array(counts + [0] * (6 - len(counts)), dtype=np.int64)
    bc.nboundaries = int(bc.neighbor_ids.shape[0])
    return bc
