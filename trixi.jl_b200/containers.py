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


class MPIInterfaceContainer:
    pass


def partition_cells(ncells, rank, world_size):
    """Contiguous chunk of the space-filling-curve element order owned by ``rank``: equal counts, the
    remainder goes to the first ranks (``partition!`` parallel_tree_mesh.jl:16-28; p4est
    ``global_first_quadrant`` dg_parallel.jl:290-296).  Returns (first, last) with last exclusive."""
    if world_size > ncells:
        raise ValueError("Too many ranks to properly partition the mesh!")
    base, rem = divmod(ncells, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def owner_of(cells, ncells, world_size):
    base, rem = divmod(ncells, world_size)
    split = rem * (base + 1)
    cells = np.asarray(cells)
    return np.where(cells < split, cells // (base + 1), rem + (cells - split) // max(base, 1))


def init_elements(mesh, basis, cells=None):
    """containers_3d.jl:87-133.  Returns inverse_jacobian[nelem], node_coordinates[ndims, n^d, nelem];
    ``cells`` restricts to the global cell indices owned by this rank."""
    nodes = basis.nodes
    n = basis.nnodes
    nd = mesh.ndims
    # integrate(one, nodes, basis): sequential sum of the weights (basis_lobatto_legendre.jl:142-150)
    reference_length = 0.0
    for w in basis.weights:
        reference_length += 1.0 * w
    reference_offset = (nodes[0] + nodes[-1]) / 2
    dx = mesh.length_at_level(mesh.levels if cells is None else mesh.levels[cells])
    jacobian = dx / reference_length
    el = ElementContainer()
    el.inverse_jacobian = 1.0 / jacobian
    centers = mesh.cell_coordinates(cells)  # [nd, nelem]
    nelem = centers.shape[1]
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


def init_interfaces(mesh, first=0, last=None, world_size=1):
    """containers_3d.jl:229-281 (count :195-227) for the elements [first, last) of this rank.  Faces whose
    other element lives on another rank become MPI interfaces (``init_mpi_interfaces!``
    containers_2d.jl:913-973; p4est ``local_neighbor_ids``/``local_sides`` dg_3d_parallel.jl:126-162):
    sorted by (neighbour rank, global interface id) so both sides enumerate a shared face identically.
    Returns (interfaces, mpi_interfaces)."""
    nd = mesh.ndims
    ncells = mesh.ncells
    last = ncells if last is None else last
    cells = None if (first == 0 and last == ncells) else np.arange(first, last, dtype=np.int64)
    nloc = last - first
    gid0 = np.arange(first, last, dtype=np.int64)
    nbp = np.empty((nloc, nd), dtype=np.int64)  # global neighbour in the positive directions
    nbm = np.empty((nloc, nd), dtype=np.int64)  # ... negative directions
    for d in range(nd):
        nbp[:, d] = mesh._face_neighbors(2 * d + 1, "same", cells)
        nbm[:, d] = mesh._face_neighbors(2 * d, "same", cells)
    orient_all = np.repeat(np.arange(1, nd + 1, dtype=np.int64)[None, :], nloc, axis=0)
    left_all = np.repeat(gid0[:, None], nd, axis=1)
    is_local_p = (nbp >= first) & (nbp < last)
    ic = InterfaceContainer()
    ic.neighbor_ids = np.asfortranarray(np.stack([left_all[is_local_p] - first + 1,
                                                  nbp[is_local_p] - first + 1]))  # [2, I], 1-based local
    ic.orientations = orient_all[is_local_p]
    ic.ninterfaces = ic.orientations.shape[0]

    mi = MPIInterfaceContainer()
    rem_p = (nbp >= 0) & ~is_local_p  # local element is the left one (local side 1)
    rem_m = (nbm >= 0) & ~((nbm >= first) & (nbm < last))  # local element is the right one (side 2)
    loc = np.concatenate([left_all[rem_p], left_all[rem_m]])
    remote = np.concatenate([nbp[rem_p], nbm[rem_m]])
    side = np.concatenate([np.ones(rem_p.sum(), dtype=np.int64), np.full(rem_m.sum(), 2, dtype=np.int64)])
    orient = np.concatenate([orient_all[rem_p], orient_all[rem_m]])
    # global interface id = (global id of the left element) * ndims + orientation - 1
    gif = np.where(side == 1, loc, remote) * nd + (orient - 1)
    peer = owner_of(remote, ncells, world_size) if remote.size else remote
    order = np.lexsort((gif, peer))
    mi.local_neighbor_ids = (loc[order] - first + 1).astype(np.int64)
    mi.local_sides = side[order]
    mi.orientations = orient[order]
    mi.neighbor_ranks = peer[order].astype(np.int64)
    mi.global_interface_ids = gif[order]
    mi.nmpiinterfaces = int(order.shape[0])
    return ic, mi


class MortarContainer:
    nmortars = 0


def init_mortars(mesh, first=0, last=None):
    """``init_mortars!`` (containers_2d.jl:706-810, containers_3d.jl:652-790): one L2 mortar per face of a
    large element whose same-level neighbour cell is refined; ``neighbor_ids[1..2^(d-1), m]`` are the small
    elements by position (2D: lower, upper; 3D: lower-left, lower-right, upper-left, upper-right), the last
    row is the large element; ``large_sides`` 1: large element on the negative side; ordered by
    (large element, direction) like the reference's loop."""
    nd = mesh.ndims
    last = mesh.ncells if last is None else last
    mc = MortarContainer()
    nsmall = 1 << (nd - 1)
    if not hasattr(mesh, "_fine_neighbors") or int(mesh.levels.min()) == int(mesh.levels.max()):
        mc.neighbor_ids = np.zeros((nsmall + 1, 0), dtype=np.int64)
        mc.large_sides = np.zeros(0, dtype=np.int64)
        mc.orientations = np.zeros(0, dtype=np.int64)
        return mc
    cells = None if (first == 0 and last == mesh.ncells) else np.arange(first, last, dtype=np.int64)
    large, small, dirs = [], [], []
    for direction in range(2 * nd):
        fine = mesh._fine_neighbors(direction, cells)
        has = (fine >= 0).all(axis=0)
        el = np.nonzero(has)[0]
        large.append(el + first)
        small.append(fine[:, el])
        dirs.append(np.full(el.shape[0], direction, dtype=np.int64))
    large, small, dirs = np.concatenate(large), np.concatenate(small, axis=1), np.concatenate(dirs)
    # mortars with an element on another rank are MPI mortars (init_mpi_mortars below)
    all_local = ((small >= first) & (small < last)).all(axis=0)
    large, small, dirs = large[all_local], small[:, all_local], dirs[all_local]
    order = np.lexsort((dirs, large))
    large, small, dirs = large[order], small[:, order], dirs[order]
    mc.neighbor_ids = np.asfortranarray(np.concatenate([small, large[None, :]]) - first + 1)
    mc.large_sides = np.where(dirs % 2 == 1, 1, 2).astype(np.int64)
    mc.orientations = (dirs // 2 + 1).astype(np.int64)
    mc.nmortars = int(large.shape[0])
    return mc


def init_boundaries(mesh, elements, basis, first=0, last=None):
    """containers_3d.jl:391-471 (for the elements [first, last) of this rank)."""
    nd = mesh.ndims
    n = basis.nnodes
    last = mesh.ncells if last is None else last
    cells = None if (first == 0 and last == mesh.ncells) else np.arange(first, last, dtype=np.int64)
    ids, orients, sides, coords, counts = [], [], [], [], []
    for direction in range(2 * nd):
        same = mesh._face_neighbors(direction, "same", cells)
        coarse = mesh._face_neighbors(direction, "coarse", cells)
        isb = (same < 0) & (coarse < 0)
        # a cell with a *refined* same-level neighbour has a non-leaf neighbour cell -> not a boundary
        if hasattr(mesh, "_has_refined_neighbor"):
            isb &= ~mesh._has_refined_neighbor(direction, cells)
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


class MPIMortarContainer:
    nmpimortars = 0


def merge_mpi_mortar_pieces(mi, pieces, nd):
    """Appends the (large element, small element) pairs of mortars that straddle ranks to the MPI interface list as
    exchange-only entries (``is_mortar_piece``): each side sends the face of its own element like for a conforming
    shared face (prolong2mpimortars! + start_mpi_send!, dg_2d_parallel.jl:600-698, dgsem_p4est/dg_3d_parallel.jl:275-380
    send the same faces inside one mortar buffer), but no interface flux is evaluated there.  ``pieces`` rows:
    (local element 0-based local id, local side, orientation, peer rank, key, node_indices or None).  Returns the
    position of every piece in the merged, (peer, key)-sorted list."""
    npc = len(pieces)
    old = mi.nmpiinterfaces
    has_idx = getattr(mi, "node_indices", None) is not None and (old > 0 or (npc and pieces[0][5] is not None))
    loc = np.concatenate([mi.local_neighbor_ids, np.array([q[0] + 1 for q in pieces], dtype=np.int64)])
    side = np.concatenate([mi.local_sides, np.array([q[1] for q in pieces], dtype=np.int64)])
    orient = np.concatenate([mi.orientations, np.array([q[2] for q in pieces], dtype=np.int64)])
    peer = np.concatenate([mi.neighbor_ranks, np.array([q[3] for q in pieces], dtype=np.int64)])
    key = np.concatenate([mi.global_interface_ids, np.array([q[4] for q in pieces], dtype=np.int64)])
    piece = np.concatenate([np.zeros(old, dtype=np.int64), np.ones(npc, dtype=np.int64)])
    if has_idx:
        idx_old = mi.node_indices if old else np.zeros((nd, 0), dtype=np.int64)
        idx_new = np.array([q[5] for q in pieces], dtype=np.int64).reshape(npc, nd).T
        node_indices = np.concatenate([idx_old, idx_new], axis=1)
    order = np.lexsort((key, peer))
    mi.local_neighbor_ids, mi.local_sides, mi.orientations = loc[order], side[order], orient[order]
    mi.neighbor_ranks, mi.global_interface_ids, mi.is_mortar_piece = peer[order], key[order], piece[order]
    if has_idx:
        mi.node_indices = np.asfortranarray(node_indices[:, order])
    mi.nmpiinterfaces = int(order.shape[0])
    where = np.empty(order.shape[0], dtype=np.int64)
    where[order] = np.arange(order.shape[0])
    return where[old:]


def init_mpi_mortars(mesh, mi, first, last, world_size):
    """``init_mpi_mortars!`` (containers_2d.jl:1131-1258; the reference has TreeMesh MPI in 2D only, here any
    dimension): mortars with elements on more than one rank.  ``neighbor_ids[p, m]`` > 0: local element (1-based),
    < 0: minus the 1-based position in the MPI interface list of the exchange entry that brings that element's face,
    0: a small element this rank neither owns nor needs (it does not own the large element).  Appends the exchange
    entries to ``mi``."""
    nd = mesh.ndims
    nsmall = 1 << (nd - 1)
    mm = MPIMortarContainer()
    mm.neighbor_ids = np.zeros((nsmall + 1, 0), dtype=np.int64)
    mm.large_sides = np.zeros(0, dtype=np.int64)
    mm.orientations = np.zeros(0, dtype=np.int64)
    if not hasattr(mi, "is_mortar_piece"):
        mi.is_mortar_piece = np.zeros(mi.nmpiinterfaces, dtype=np.int64)
    if world_size == 1 or not hasattr(mesh, "_fine_neighbors") or int(mesh.levels.min()) == int(mesh.levels.max()):
        return mm
    large, small, dirs = [], [], []
    for direction in range(2 * nd):
        fine = mesh._fine_neighbors(direction, None)
        el = np.nonzero((fine >= 0).all(axis=0))[0]
        large.append(el)
        small.append(fine[:, el])
        dirs.append(np.full(el.shape[0], direction, dtype=np.int64))
    large, small, dirs = np.concatenate(large), np.concatenate(small, axis=1), np.concatenate(dirs)
    order = np.lexsort((dirs, large))
    large, small, dirs = large[order], small[:, order], dirs[order]
    local = lambda e: (e >= first) & (e < last)  # noqa: E731
    ids_all = np.concatenate([small, large[None, :]])
    mixed = local(ids_all).any(axis=0) & ~local(ids_all).all(axis=0)
    base_key = mesh.ncells * nd  # above every conforming interface id
    pieces, slots, records = [], [], []
    for m in np.nonzero(mixed)[0]:
        L, direction = int(large[m]), int(dirs[m])
        o, large_side = direction // 2 + 1, 1 if direction % 2 == 1 else 2
        owner_L = int(owner_of(np.array([L]), mesh.ncells, world_size)[0])
        rec = [0] * (nsmall + 1)
        rec[nsmall] = L - first + 1 if local(L) else 0
        for p in range(nsmall):
            s = int(small[p, m])
            owner_s = int(owner_of(np.array([s]), mesh.ncells, world_size)[0])
            key = base_key + (int(m) * nsmall + p)
            if local(s):
                rec[p] = s - first + 1
            if owner_s == owner_L:
                continue
            if local(L):    # send the large face, receive the small face
                pieces.append((L - first, large_side, o, owner_s, key, None))
                slots.append((len(records), p))
            elif local(s):  # send the small face, receive the large face
                pieces.append((s - first, 3 - large_side, o, owner_L, key, None))
                slots.append((len(records), nsmall))
        records.append((rec, large_side, o))
    where = merge_mpi_mortar_pieces(mi, pieces, nd)
    for (r, p), w in zip(slots, where):
        if records[r][0][p] == 0:
            records[r][0][p] = -(int(w) + 1)
    if records:
        mm.neighbor_ids = np.asfortranarray(np.array([r[0] for r in records], dtype=np.int64).T)
        mm.large_sides = np.array([r[1] for r in records], dtype=np.int64)
        mm.orientations = np.array([r[2] for r in records], dtype=np.int64)
    mm.nmpimortars = len(records)
    return mm
