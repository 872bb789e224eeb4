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
A forest may be refined non-uniformly before the semidiscretization is built -- ``refine(refine_fn)`` is
``refine_p4est!(mesh.p4est, recursive, refine_fn_c, C_NULL)`` with the callback's (which_tree, quadrant.x/.y/.z,
quadrant.level) (p4est_mesh.jl:2457-2482), ``balance()`` the 2:1 face balance ``create_cache`` applies
(``balance!`` p4est_mesh.jl:2484-2493, dgsem_p4est/dg.jl:13-16).  Hanging faces become L2 mortars
(``init_mortars!`` dgsem_p4est/containers.jl:686-689,902-948, node indices containers_3d.jl:159-201).
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
        self.levels = np.full(self.ncells, L, dtype=np.int64)
        self.is_uniform = True
        self.tree_node_coordinates = self._calc_tree_node_coordinates()

    # P4EST_MAXLEVEL = 30, P8EST_MAXLEVEL = 19: quadrant.x/.y/.z are multiples of root_len / 2^level
    @property
    def root_len(self):
        return 1 << (30 if self.ndims == 2 else 19)

    def _set_leaves(self, tree, level, coords):
        """Store the leaves sorted tree by tree, Morton order inside a tree (p4est's quadrant order)."""
        nd = self.ndims
        lmax = int(level.max()) if level.size else 0
        fine = coords << (lmax - level)[None, :]
        order = np.lexsort((morton_key(fine, nd), tree))
        self.tree_of_element = tree[order]
        self.levels = level[order]
        self.quad_coords = coords[:, order]
        self.is_uniform = bool(np.all(self.levels == self.levels[0]))
        if self.is_uniform:
            L = int(self.levels[0])
            self.initial_refinement_level = L
            q = 1 << L
            tcoord = np.stack(np.unravel_index(self.tree_of_element, self.trees_per_dimension, order="F"))
            self.global_coords = tcoord * q + self.quad_coords
            self.cells_per_dimension = tuple(t * q for t in self.trees_per_dimension)
            lut = np.full(self.cells_per_dimension, -1, dtype=np.int64)
            lut[tuple(self.global_coords)] = np.arange(self.ncells, dtype=np.int64)
            self._lut = lut
        else:
            self._lut = self.global_coords = self.cells_per_dimension = None
        self._leaf_index = None
        self._surfaces = None

    def refine(self, refine_fn, recursive=True):
        """``refine_p4est!``: ``refine_fn(which_tree, x, y[, z], level) -> bool`` with p4est's integer quadrant
        coordinates; ``recursive`` re-offers the children.  Call ``balance()`` (or build a semidiscretization,
        which balances like the reference's ``create_cache``) afterwards."""
        nd = self.ndims
        todo = [(int(t), int(l), tuple(int(c) for c in self.quad_coords[:, e]))
                for e, (t, l) in enumerate(zip(self.tree_of_element, self.levels))]
        done = []
        while todo:
            nxt = []
            for t, l, c in todo:
                h = self.root_len >> l
                if refine_fn(t, *[ci * h for ci in c], l):
                    children = [(t, l + 1, tuple(2 * c[d] + ((child >> d) & 1) for d in range(nd)))
                                for child in range(1 << nd)]
                    (nxt if recursive else done).extend(children)
                else:
                    done.append((t, l, c))
            todo = nxt
        self._set_leaves(np.array([x[0] for x in done], dtype=np.int64), np.array([x[1] for x in done], dtype=np.int64),
                         np.array([x[2] for x in done], dtype=np.int64).T.reshape(nd, -1))
        return self

    def leaf_index(self):
        if getattr(self, "_leaf_index", None) is None:
            self._leaf_index = {(int(t), int(l)) + tuple(int(c) for c in self.quad_coords[:, e]): e
                                for e, (t, l) in enumerate(zip(self.tree_of_element, self.levels))}
        return self._leaf_index

    def neighbor_cell(self, tree, level, coords, d, side):
        """(tree, coords) of the same-size cell across face (d, side) of a quadrant, or None at a domain boundary."""
        tp = self.trees_per_dimension
        tc = list(np.unravel_index(tree, tp, order="F"))
        c = list(coords)
        c[d] += 1 if side else -1
        q = 1 << level
        if c[d] < 0 or c[d] >= q:
            tc[d] += 1 if side else -1
            c[d] %= q
            if tc[d] < 0 or tc[d] >= tp[d]:
                if not self.periodicity[d]:
                    return None
                tc[d] %= tp[d]
        return int(np.ravel_multi_index(tuple(int(x) for x in tc), tp, order="F")), tuple(c)

    def find_leaf(self, tree, level, coords):
        """The leaf covering cell (tree, level, coords): (element, its level), or None if the cell is subdivided."""
        idx = self.leaf_index()
        for k in range(level + 1):
            e = idx.get((tree, level - k) + tuple(ci >> k for ci in coords))
            if e is not None:
                return e, level - k
        return None

    def balance(self):
        """2:1 balance across faces (P4EST_CONNECT_FACE): a leaf whose face neighbour is more than one level finer is
        split, until nothing changes."""
        nd = self.ndims
        while True:
            split = set()
            for e in range(self.ncells):
                t, l = int(self.tree_of_element[e]), int(self.levels[e])
                c = tuple(int(x) for x in self.quad_coords[:, e])
                for d in range(nd):
                    for side in (0, 1):
                        nb = self.neighbor_cell(t, l, c, d, side)
                        if nb is None:
                            continue
                        found = self.find_leaf(nb[0], l, nb[1])
                        if found is not None and found[1] < l - 1:
                            split.add(found[0])
            if not split:
                return self
            keep = np.array([e not in split for e in range(self.ncells)])
            tree, level, coords = [self.tree_of_element[keep]], [self.levels[keep]], [self.quad_coords[:, keep]]
            for e in sorted(split):
                for child in range(1 << nd):
                    tree.append(self.tree_of_element[e:e + 1])
                    level.append(self.levels[e:e + 1] + 1)
                    coords.append((2 * self.quad_coords[:, e] + np.array([(child >> d) & 1 for d in range(nd)]))[:, None])
            self._set_leaves(np.concatenate(tree), np.concatenate(level), np.concatenate(coords, axis=1))

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
    last = mesh.ncells if last is None else last
    nelem = last - first
    X = np.empty((nd,) + (n,) * nd + (nelem,), order="F")
    levels = mesh.levels[first:last]
    for L in np.unique(levels):
        sel = np.nonzero(levels == L)[0]
        quad_length = 1.0 / (1 << int(L))
        # interpolation matrices only depend on the quadrant coordinate along one axis
        # (quad_length = p4est_quadrant_len(level) / p4est_root_len, anchor = quad.x / p4est_root_len)
        used = np.unique(mesh.quad_coords[:, first:last][:, sel])
        mats = np.zeros((1 << int(L), n, mesh.nodes.shape[0]))
        for c in used:
            nodes_out = 2 * (quad_length * 1 / 2 * (basis.nodes + 1) + c * quad_length) - 1
            mats[c] = polynomial_interpolation_matrix(mesh.nodes, nodes_out)
        data = mesh.tree_node_coordinates[..., mesh.tree_of_element[first:last][sel]]  # [nd, nm.., nsel]
        for d in range(nd):
            M = mats[mesh.quad_coords[d, first:last][sel]]  # [nsel, n, nm]
            # contract axis 1+d of data with the last axis of M, per element
            data = np.moveaxis(data, 1 + d, -2)            # [..., nm, nsel]
            data = np.einsum("eij,...je->...ie", M, data)  # [..., n, nsel]
            data = np.moveaxis(data, -2, 1 + d)
        X[..., sel] = data
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


class P4estMortarContainer:
    """P4estMortarContainer (dgsem_p4est/containers.jl:563-613): ``neighbor_ids [2^(d-1)+1, M]`` (small elements by
    position, the large element last), ``node_indices [ndims, 2, M]`` (1: small side, 2: large side)."""
    pass


def _init_surfaces_general(mesh, first=0, last=None, world_size=1, basis=None):
    """init_surfaces! (dgsem_p4est/containers.jl:786-838; with ranks init_surfaces_iter_face_parallel
    containers_parallel.jl:380-520) for a forest with hanging faces: p4est's face iteration is replaced by a walk over
    the leaves and their 2*ndims faces.  A face whose same-size neighbour cell is a leaf of the same level is an
    interface (taken from the negative side: primary = the element whose + face it is), a subdivided neighbour cell
    makes a mortar with this leaf as the large element, a coarser neighbour is handled from its side.  Trees of a
    brick share their axes (orientation code 0, opposite faces), so both sides index forward
    (orientation_to_indices_p4est containers_3d.jl:206-300 with flipped = false, code 0).
    Elements [first, last) belong to this rank: surfaces with elements of two ranks become MPI interfaces / MPI
    mortars.  Returns (interfaces, mortars, boundaries, mpi_interfaces, mpi_mortars)."""
    last = mesh.ncells if last is None else last
    cache_key = (first, last, world_size)
    if getattr(mesh, "_surfaces", None) is not None and mesh._surfaces[0] == cache_key:
        return mesh._surfaces[1]
    from .containers import MPIMortarContainer, merge_mpi_mortar_pieces
    nd = mesh.ndims
    npos = 1 << (nd - 1)
    surf_dims = [[c for c in range(nd) if c != d] for d in range(nd)]
    local = lambda e: first <= e < last  # noqa: E731
    owner = lambda e: int(owner_of(np.array([e]), mesh.ncells, world_size)[0])  # noqa: E731
    prim, sec, idim = [], [], []
    m_ids, m_idx = [], []
    bnd = [[] for _ in range(2 * nd)]
    mpi_rows = []          # conforming faces shared with another rank
    mm_records, pieces, slots = [], [], []
    base_key = mesh.ncells * nd
    for e in range(mesh.ncells):
        t, l = int(mesh.tree_of_element[e]), int(mesh.levels[e])
        c = tuple(int(x) for x in mesh.quad_coords[:, e])
        for d in range(nd):
            for side in (0, 1):
                nb = mesh.neighbor_cell(t, l, c, d, side)
                if nb is None:
                    if local(e):
                        bnd[2 * d + side].append(e - first)
                    continue
                found = mesh.find_leaf(nb[0], l, nb[1])
                if found is not None:
                    if found[1] == l and side == 1:
                        o = found[0]
                        if local(e) and local(o):
                            prim.append(e - first)
                            sec.append(o - first)
                            idim.append(d)
                        elif local(e):  # local element is the primary one
                            mpi_rows.append((e - first, 1, d + 1, owner(o), e * nd + d, _face_indices(nd, d, 1)))
                        elif local(o):
                            mpi_rows.append((o - first, 2, d + 1, owner(e), e * nd + d, _face_indices(nd, d, 0)))
                    elif found[1] < l - 1:
                        raise ValueError("the forest is not 2:1 balanced; call mesh.balance()")
                    continue
                # hanging face: the 2^(d-1) small elements in the z-order of the face coordinates
                small = []
                for pos in range(npos):
                    cc = [2 * x for x in nb[1]]
                    cc[d] += 0 if side == 1 else 1
                    for k, sd in enumerate(surf_dims[d]):
                        cc[sd] += (pos >> k) & 1
                    f2 = mesh.find_leaf(nb[0], l + 1, tuple(cc))
                    if f2 is None or f2[1] != l + 1:
                        raise ValueError("the forest is not 2:1 balanced; call mesh.balance()")
                    small.append(f2[0])
                idx = [_face_indices(nd, d, 1 - side), _face_indices(nd, d, side)]
                members = small + [e]
                nloc = sum(local(x) for x in members)
                if nloc == len(members):
                    m_ids.append([x - first for x in members])
                    m_idx.append(idx)
                elif nloc > 0:
                    # MPI mortar (init_mpi_mortars! containers_parallel.jl:130-210): local elements by id, the others
                    # through exchange-only entries of the MPI interface list (see merge_mpi_mortar_pieces)
                    rec = [(x - first + 1) if local(x) else 0 for x in members]
                    owner_l = owner(e)
                    mortar_key = e * 2 * nd + 2 * d + side
                    for pos, sm in enumerate(small):
                        owner_s = owner(sm)
                        if owner_s == owner_l:
                            continue
                        key = base_key + mortar_key * npos + pos
                        if local(e):
                            pieces.append((e - first, 2, d + 1, owner_s, key, idx[1]))
                            slots.append((len(mm_records), pos))
                        elif local(sm):
                            pieces.append((sm - first, 1, d + 1, owner_l, key, idx[0]))
                            slots.append((len(mm_records), npos))
                    mm_records.append((rec, idx, small))
    face_idx = np.array([[_face_indices(nd, d, 1), _face_indices(nd, d, 0)] for d in range(nd)], dtype=np.int64)
    ic = InterfaceContainer()
    prim, sec, idim = (np.array(x, dtype=np.int64) for x in (prim, sec, idim))
    ic.neighbor_ids = np.asfortranarray(np.stack([prim + 1, sec + 1])) if prim.size else np.zeros((2, 0), dtype=np.int64)
    ic.node_indices = (np.asfortranarray(face_idx[idim].transpose(2, 1, 0)) if prim.size
                       else np.zeros((nd, 2, 0), dtype=np.int64))
    ic.orientations = np.zeros(prim.shape[0], dtype=np.int64)
    ic.ninterfaces = int(prim.shape[0])
    mc = P4estMortarContainer()
    mc.nmortars = len(m_ids)
    mc.neighbor_ids = (np.asfortranarray(np.array(m_ids, dtype=np.int64).T + 1) if m_ids
                       else np.zeros((npos + 1, 0), dtype=np.int64))
    mc.node_indices = (np.asfortranarray(np.array(m_idx, dtype=np.int64).transpose(2, 1, 0)) if m_ids
                       else np.zeros((nd, 2, 0), dtype=np.int64))
    bc = BoundaryContainer()
    counts = [len(b) for b in bnd]
    bc.neighbor_ids = np.array([e + 1 for b in bnd for e in b], dtype=np.int64)
    bc.node_indices = (np.asfortranarray(np.array([_face_indices(nd, dr // 2, dr % 2) for dr, b in enumerate(bnd) for _ in b],
                                                  dtype=np.int64).T.reshape(nd, -1)))
    bc.orientations = np.array([dr // 2 + 1 for dr, b in enumerate(bnd) for _ in b], dtype=np.int64)
    bc.neighbor_sides = np.array([2 if dr % 2 == 0 else 1 for dr, b in enumerate(bnd) for _ in b], dtype=np.int64)
    bc.node_coordinates = np.zeros((nd, 0))
    bc.n_boundaries_per_direction = np.array(counts + [0] * (6 - len(counts)), dtype=np.int64)
    bc.nboundaries = int(bc.neighbor_ids.shape[0])
    # MPI interfaces sorted by (neighbour rank, global interface id), then the mortar pieces merged in
    mi = _empty_mpi_interfaces(nd)
    if mpi_rows:
        rows = sorted(mpi_rows, key=lambda r: (r[3], r[4]))
        mi.local_neighbor_ids = np.array([r[0] + 1 for r in rows], dtype=np.int64)
        mi.local_sides = np.array([r[1] for r in rows], dtype=np.int64)
        mi.orientations = np.array([r[2] for r in rows], dtype=np.int64)
        mi.neighbor_ranks = np.array([r[3] for r in rows], dtype=np.int64)
        mi.global_interface_ids = np.array([r[4] for r in rows], dtype=np.int64)
        mi.node_indices = np.asfortranarray(np.array([r[5] for r in rows], dtype=np.int64).T.reshape(nd, -1))
        mi.nmpiinterfaces = len(rows)
    mi.is_mortar_piece = np.zeros(mi.nmpiinterfaces, dtype=np.int64)
    mm = MPIMortarContainer()
    mm.neighbor_ids = np.zeros((npos + 1, 0), dtype=np.int64)
    mm.node_indices = np.zeros((nd, 2, 0), dtype=np.int64)
    mm.small_elements_global = np.zeros((npos, 0), dtype=np.int64)
    if mm_records:
        where = merge_mpi_mortar_pieces(mi, pieces, nd)
        for (r, p), w in zip(slots, where):
            if mm_records[r][0][p] == 0:
                mm_records[r][0][p] = -(int(w) + 1)
        mm.neighbor_ids = np.asfortranarray(np.array([r[0] for r in mm_records], dtype=np.int64).T)
        mm.node_indices = np.asfortranarray(np.array([r[1] for r in mm_records], dtype=np.int64).transpose(2, 1, 0))
        mm.small_elements_global = np.array([r[2] for r in mm_records], dtype=np.int64).T
    mm.nmpimortars = len(mm_records)
    mesh._surfaces = (cache_key, (ic, mc, bc, mi, mm))
    return mesh._surfaces[1]


def init_mpi_mortar_normals_p4est(mesh, basis, mm):
    """``normal_directions`` of the MPI mortar container (dgsem_p4est/containers_parallel.jl:130-155,
    init_normal_directions! :557-620): the outward normals of the small elements at the mortar's face nodes,
    [ndims, n^(d-1), 2^(d-1), MM] -- the small elements may live on another rank, so the normals cannot be read from
    the local contravariant vectors."""
    nd, n = mesh.ndims, basis.nnodes
    nf, npos = n ** (nd - 1), 1 << (nd - 1)
    out = np.zeros((nd, nf, npos, mm.nmpimortars), order="F")
    if mm.nmpimortars == 0:
        return out
    elements = np.unique(mm.small_elements_global)
    pos_of = {int(e): k for k, e in enumerate(elements)}
    ja = {}
    for e in elements:  # one element at a time: exactly the arithmetic of the owner's init_elements_p4est
        el = init_elements_p4est(mesh, basis, int(e), int(e) + 1)
        ja[int(e)] = el.contravariant_vectors[..., 0].reshape((nd, nd) + (n,) * nd, order="F")
    del pos_of
    for m in range(mm.nmpimortars):
        sidx = mm.node_indices[:, 0, m]
        direction = [k for k in range(nd) if sidx[k] in (IDX_BEGIN, IDX_END)][0]
        sign = 1.0 if sidx[direction] == IDX_END else -1.0
        for p in range(npos):
            J = ja[int(mm.small_elements_global[p, m])]  # [nd (component), nd (which vector), n, n(, n)]
            face = np.take(J[:, direction], n - 1 if sidx[direction] == IDX_END else 0, axis=1 + direction)
            out[:, :, p, m] = sign * face.reshape(nd, nf, order="F")
    return out


def _empty_mpi_interfaces(nd):
    mi = MPIInterfaceContainer()
    mi.local_neighbor_ids = mi.local_sides = mi.orientations = mi.neighbor_ranks = np.zeros(0, dtype=np.int64)
    mi.global_interface_ids = np.zeros(0, dtype=np.int64)
    mi.node_indices = np.zeros((nd, 0), dtype=np.int64)
    mi.nmpiinterfaces = 0
    return mi


def init_mortars_p4est(mesh, first=0, last=None, world_size=1):
    """init_mortars! (dgsem_p4est/containers.jl:646-689).  Conforming forests have none."""
    if mesh.is_uniform:
        mc = P4estMortarContainer()
        mc.nmortars = 0
        mc.neighbor_ids = np.zeros(((1 << (mesh.ndims - 1)) + 1, 0), dtype=np.int64)
        mc.node_indices = np.zeros((mesh.ndims, 2, 0), dtype=np.int64)
        return mc
    return _init_surfaces_general(mesh, first, last, world_size)[1]


def init_mpi_mortars_p4est(mesh, basis, first=0, last=None, world_size=1):
    """init_mpi_mortars! (dgsem_p4est/containers_parallel.jl:130-210) incl. the small elements' normals."""
    from .containers import MPIMortarContainer
    if mesh.is_uniform:
        mm = MPIMortarContainer()
        mm.neighbor_ids = np.zeros(((1 << (mesh.ndims - 1)) + 1, 0), dtype=np.int64)
        mm.node_indices = np.zeros((mesh.ndims, 2, 0), dtype=np.int64)
        mm.normal_directions = np.zeros((mesh.ndims, 0))
        return mm
    mm = _init_surfaces_general(mesh, first, last, world_size)[4]
    mm.normal_directions = init_mpi_mortar_normals_p4est(mesh, basis, mm)
    return mm


def init_interfaces_p4est(mesh, first=0, last=None, world_size=1):
    """init_interfaces! (dgsem_p4est/containers.jl:264-300) for a conforming brick forest: one interface per
    interior (or periodic) face, primary = the element on the negative side.  Faces whose other element
    belongs to another rank become MPI interfaces (init_mpi_interfaces! containers_parallel.jl:85-127,
    ``local_neighbor_ids``/``local_sides``/``node_indices`` of the local side), sorted by
    (neighbour rank, global interface id) (init_mpi_neighbor_connectivity dg_parallel.jl:313-390).
    Returns (interfaces, mpi_interfaces)."""
    nd = mesh.ndims
    if not mesh.is_uniform:
        surf = _init_surfaces_general(mesh, first, last, world_size)
        return surf[0], surf[3]
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


def init_boundaries_p4est(mesh, first=0, last=None, world_size=1):
    """init_boundaries! (dgsem_p4est/containers.jl:302-345), sorted by boundary name
    :x_neg, :x_pos, :y_neg, ... (structured_boundary_names! p4est_mesh.jl:298-365); elements [first, last)."""
    nd = mesh.ndims
    if not mesh.is_uniform:
        return _init_surfaces_general(mesh, first, last, world_size)[2]
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
