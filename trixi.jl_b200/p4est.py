"""``P4estMesh`` for programmatically built forests (host side).

The reference wraps libp4est (``src/meshes/p4est_mesh.jl``); mesh construction is out of the hot path, so
this module rebuilds exactly what the solver containers need for the constructor the reference's
library-free examples use -- ``P4estMesh(trees_per_dimension; polydeg, mapping | coordinates_min/max | faces,
initial_refinement_level, periodicity)`` (p4est_mesh.jl:189-247):
  * trees in linear (x-fastest) order -- ``connectivity_structured`` (p4est_mesh.jl:1149-1290), *not*
    p4est's Z-order brick -- each uniformly refined, quadrants of a tree in Morton order, elements numbered
    tree by tree;
  * geometry stored per tree as an interpolation polynomial of degree ``polydeg`` of the mapping
    (``calc_tree_node_coordinates!`` p4est_mesh.jl:1623-1678) and interpolated to every element's nodes
    (``calc_node_coordinates!`` dgsem_p4est/containers_3d.jl:39-77);
  * interface / boundary containers with symbolic ``node_indices`` (dgsem_p4est/containers.jl:226-262,
    init_interface_node_indices! containers_3d.jl:86-143), encoded as integers for the C ABI:
    :begin 0, :end 1, :i_forward 2, :i_backward 3, :j_forward 4, :j_backward 5.
Mesh files (``P4estMesh{NDIMS}(meshfile)``) and AMR need libp4est and stay with the reference.
"""
from __future__ import annotations

import numpy as np

from .basis import LobattoLegendreBasis, polynomial_interpolation_matrix
from .containers import BoundaryContainer, InterfaceContainer, MPIInterfaceContainer, owner_of
from .mesh import morton_key
from .structured import compute_metric_terms, coordinates2mapping, transfinite_mapping

IDX_BEGIN, IDX_END, IDX_I_FORWARD, IDX_I_BACKWARD, IDX_J_FORWARD, IDX_J_BACKWARD = range(6)


class P4estMesh:
    def __init__(self, trees_per_dimension, polydeg, mapping=None, faces=None, coordinates_min=None,
                 coordinates_max=None, initial_refinement_level=0, periodicity=False):
        self.trees_per_dimension = tuple(int(t) for t in trees_per_dimension)
        self.ndims = len(self.trees_per_dimension)
        if sum(x is not None for x in (mapping, faces, coordinates_min)) != 1:
            raise ValueError("Exactly one of mapping, faces and coordinates_min/max must be specified")
        if faces is not None:
            mapping = transfinite_mapping(tuple(faces))
        elif coordinates_min is not None:
            mapping = coordinates2mapping(tuple(coordinates_min), tuple(coordinates_max))
        self.mapping = mapping
        self.polydeg = int(polydeg)
        self.initial_refinement_level = int(initial_refinement_level)
        if isinstance(periodicity, bool):
            periodicity = (periodicity,) * self.ndims
        self.periodicity = tuple(bool(p) for p in periodicity)
        self.nodes = LobattoLegendreBasis(self.polydeg).nodes
        nd, L = self.ndims, self.initial_refinement_level
        ntrees = int(np.prod(self.trees_per_dimension))
        # quadrants of one tree in Morton order
        q = 1 << L
        grids = np.meshgrid(*[np.arange(q, dtype=np.int64)] * nd, indexing="ij")
        qc = np.stack([g.ravel() for g in grids])
        qc = qc[:, np.argsort(morton_key(qc, nd), kind="stable")]
        nq = qc.shape[1]
        tree_ids = np.repeat(np.arange(ntrees, dtype=np.int64), nq)
        tcoord = np.stack(np.unravel_index(tree_ids, self.trees_per_dimension, order="F"))
        self.tree_of_element = tree_ids
        self.quad_coords = np.tile(qc, (1, ntrees))                  # [nd, nelem] within the tree
        self.global_coords = tcoord * q + self.quad_coords           # [nd, nelem] on the global grid
        self.cells_per_dimension = tuple(t * q for t in self.trees_per_dimension)
        # global grid -> element id
        lut = np.full(self.cells_per_dimension, -1, dtype=np.int64)
        lut[tuple(self.global_coords)] = np.arange(self.ncells, dtype=np.int64)
        self._lut = lut
        self.tree_node_coordinates = self._calc_tree_node_coordinates()

    @property
    def ncells(self):
        return self.tree_of_element.shape[0]

    def _calc_tree_node_coordinates(self):
        nd, n = self.ndims, self.nodes.shape[0]
        tp = self.trees_per_dimension
        ref = []
        for d in range(nd):
            dx = 2 / tp[d]
            offs = -1 + np.arange(tp[d]) * dx + dx / 2
            ref.append(offs[None, :] + dx / 2 * self.nodes[:, None])
        grids = []
        for d in range(nd):
            sh = [1] * (2 * nd)
            sh[d] = n
            sh[nd + d] = tp[d]
            grids.append(ref[d].reshape(sh))
        grids = np.broadcast_arrays(*grids)
        xyz = self.mapping(*grids)
        coords = np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), grids[0].shape) for c in xyz])
        perm = (0,) + tuple(range(1, nd + 1)) + tuple(range(2 * nd, nd, -1))
        return np.asfortranarray(coords.transpose(perm).reshape((nd,) + (n,) * nd + (int(np.prod(tp)),)))

    def __repr__(self):
        return f"P4estMesh{{{self.ndims}}} trees {self.trees_per_dimension} level {self.initial_refinement_level}"


class P4estElementContainer:
    pass


def init_elements_p4est(mesh, basis, first=0, last=None):
    """init_elements! dgsem_p4est/containers.jl:58-76 + containers_3d.jl:9-77 for the elements
    [first, last) of this rank (p4est partitions the space-filling curve into contiguous chunks,
    ``global_first_quadrant`` dg_parallel.jl:290-296)."""
    nd, n = mesh.ndims, basis.nnodes
    if n < mesh.nodes.shape[0]:
        raise ValueError("The solver can't have a lower polydeg than the mesh")
    L = mesh.initial_refinement_level
    quad_length = 1.0 / (1 << L)
    last = mesh.ncells if last is None else last
    nelem = last - first
    X = np.empty((nd,) + (n,) * nd + (nelem,), order="F")
    # interpolation matrices only depend on the quadrant coordinate along one axis
    mats = []
    for c in range(1 << L):
        nodes_out = 2 * (quad_length * 1 / 2 * (basis.nodes + 1) + c * quad_length) - 1
        mats.append(polynomial_interpolation_matrix(mesh.nodes, nodes_out))
    mats = np.stack(mats)  # [2^L, n, n_mesh]
    T = mesh.tree_node_coordinates[..., mesh.tree_of_element[first:last]]  # [nd, nm.., nelem]
    data = T
    for d in range(nd):
        M = mats[mesh.quad_coords[d, first:last]]  # [nelem, n, nm]
        # contract axis 1+d of data with the last axis of M, per element
        data = np.moveaxis(data, 1 + d, -2)            # [..., nm, nelem]
        data = np.einsum("eij,...je->...ie", M, data)  # [..., n, nelem]
        data = np.moveaxis(data, -2, 1 + d)
    X[...] = data
    el = P4estElementContainer()
    el.nelements = nelem
    el.node_coordinates = X
    J, Ja, inv_jac = compute_metric_terms(X, basis.derivative_matrix)
    el.jacobian_matrix, el.contravariant_vectors, el.inverse_jacobian = J, Ja, inv_jac
    return el


def _face_indices(nd, d, side):
    """node_indices tuple of the face normal to dimension d (0-based); side 0: :begin, 1: :end; the surface
    axes run forward (primary alignment, containers_3d.jl:86-143)."""
    idx = []
    fw = [IDX_I_FORWARD, IDX_J_FORWARD]
    k = 0
    for c in range(nd):
        if c == d:
            idx.append(IDX_END if side else IDX_BEGIN)
        else:
            idx.append(fw[k])
            k += 1
    return idx


def init_interfaces_p4est(mesh, first=0, last=None, world_size=1):
    """init_interfaces! (dgsem_p4est/containers.jl:264-300) for a conforming brick forest: one interface per
    interior (or periodic) face, primary = the element on the negative side.  Faces whose other element
    belongs to another rank become MPI interfaces (init_mpi_interfaces! containers_parallel.jl:85-127,
    ``local_neighbor_ids``/``local_sides``/``node_indices`` of the local side), sorted by
    (neighbour rank, global interface id) (init_mpi_neighbor_connectivity dg_parallel.jl:313-390).
    Returns (interfaces, mpi_interfaces)."""
    nd = mesh.ndims
    cells = mesh.cells_per_dimension
    gc = mesh.global_coords
    last = mesh.ncells if last is None else last
    prim, sec, dims = [], [], []
    for d in range(nd):
        nb = gc.copy()
        nb[d] += 1
        outside = nb[d] >= cells[d]
        if mesh.periodicity[d]:
            nb[d] %= cells[d]
            valid = np.ones(mesh.ncells, dtype=bool)
        else:
            nb[d] = np.minimum(nb[d], cells[d] - 1)
            valid = ~outside
        nbid = mesh._lut[tuple(nb)]
        el = np.nonzero(valid)[0]
        prim.append(el)
        sec.append(nbid[valid])
        dims.append(np.full(el.shape[0], d, dtype=np.int64))
    prim, sec, dims = np.concatenate(prim), np.concatenate(sec), np.concatenate(dims)
    face_idx = np.array([[_face_indices(nd, d, 1), _face_indices(nd, d, 0)] for d in range(nd)],
                        dtype=np.int64)  # [d, side (0 primary: :end face, 1 secondary: :begin face), nd]
    ploc = (prim >= first) & (prim < last)
    sloc = (sec >= first) & (sec < last)
    both = ploc & sloc
    ic = InterfaceContainer()
    ic.neighbor_ids = np.asfortranarray(np.stack([prim[both] - first + 1, sec[both] - first + 1]))
    ic.node_indices = np.asfortranarray(face_idx[dims[both]].transpose(2, 1, 0))  # [nd, 2, I]
    ic.orientations = np.zeros(ic.neighbor_ids.shape[1], dtype=np.int64)
    ic.ninterfaces = ic.neighbor_ids.shape[1]

    mi = MPIInterfaceContainer()
    only_p, only_s = ploc & ~sloc, sloc & ~ploc
    loc = np.concatenate([prim[only_p], sec[only_s]])
    remote = np.concatenate([sec[only_p], prim[only_s]])
    side = np.concatenate([np.ones(only_p.sum(), dtype=np.int64), np.full(only_s.sum(), 2, dtype=np.int64)])
    dim = np.concatenate([dims[only_p], dims[only_s]])
    gif = np.concatenate([prim[only_p], prim[only_s]]) * nd + dim  # global interface id: primary element, dimension
    peer = owner_of(remote, mesh.ncells, world_size) if remote.size else remote
    order = np.lexsort((gif, peer))
    mi.local_neighbor_ids = (loc[order] - first + 1).astype(np.int64)
    mi.local_sides = side[order]
    mi.orientations = (dim[order] + 1).astype(np.int64)
    mi.node_indices = np.asfortranarray(face_idx[dim[order], side[order] - 1].T.reshape(nd, -1))  # [nd, MI]
    mi.neighbor_ranks = peer[order].astype(np.int64)
    mi.global_interface_ids = gif[order]
    mi.nmpiinterfaces = int(order.shape[0])
    return ic, mi


def init_boundaries_p4est(mesh, first=0, last=None):
    """init_boundaries! (dgsem_p4est/containers.jl:302-345), sorted by boundary name
    :x_neg, :x_pos, :y_neg, ... (structured_boundary_names! p4est_mesh.jl:298-365); elements [first, last)."""
    nd = mesh.ndims
    cells = mesh.cells_per_dimension
    last = mesh.ncells if last is None else last
    gc = mesh.global_coords[:, first:last]
    ids, nidx, counts = [], [], []
    for direction in range(2 * nd):
        d, side = direction // 2, direction % 2
        if mesh.periodicity[d]:
            counts.append(0)
            continue
        el = np.nonzero(gc[d] == (cells[d] - 1 if side else 0))[0]
        ids.append(el)
        counts.append(el.shape[0])
        ni = np.empty((nd, el.shape[0]), dtype=np.int64)
        ni[:] = np.array(_face_indices(nd, d, side))[:, None]
        nidx.append(ni)
    bc = BoundaryContainer()
    bc.neighbor_ids = (np.concatenate(ids) + 1).astype(np.int64) if ids else np.zeros(0, dtype=np.int64)
    bc.node_indices = np.asfortranarray(np.concatenate(nidx, axis=1)) if nidx else np.zeros((nd, 0), dtype=np.int64)
    orient, side_l = [], []
    for direction, c in enumerate(counts):
        orient += [direction // 2 + 1] * c
        side_l += [2 if direction % 2 == 0 else 1] * c
    bc.orientations = np.array(orient, dtype=np.int64)
    bc.neighbor_sides = np.array(side_l, dtype=np.int64)
    bc.node_coordinates = np.zeros((nd, 0))
    bc.n_boundaries_per_direction = np.array(counts + [0] * (6 - len(counts)), dtype=np.int64)
    bc.nboundaries = int(bc.neighbor_ids.shape[0])
    return bc
