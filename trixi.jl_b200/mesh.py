"""Host-side meshes. Mesh construction is out of the timed hot path (BASELINE.json north_star:
"AMR and mesh construction remain host-side"); only the resulting connectivity arrays cross the
C-ABI boundary.

``TreeMesh`` mirrors ``src/meshes/tree_mesh.jl:128-186`` for a hypercube domain with uniform initial
refinement plus optional box refinement patches.  Instead of the reference's pointer-linked
``SerialTree`` (``src/meshes/serial_tree.jl``) the leaves are kept as integer (level, ix, iy, iz)
tuples sorted in the reference's depth-first order: children are inserted directly behind their
parent (``abstract_tree.jl:331``) in the child order x-fastest (``abstract_tree.jl:105-130``), so a
leaf's position is given by its Morton key -- what the reference's ``leaf_cells`` traversal yields.
All neighbour searches are vectorised key look-ups (2M+ elements build in seconds).
"""
from __future__ import annotations

import numpy as np


def _part1by1(x):
    x = x.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
    x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    x = (x | (x << np.uint64(2))) & np.uint64(0x3333333333333333)
    x = (x | (x << np.uint64(1))) & np.uint64(0x5555555555555555)
    return x


def _part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_key(coords, ndims):
    """Interleave integer coordinates, x in the lowest bit (child order of abstract_tree.jl:105-130)."""
    if ndims == 2:
        return _part1by1(coords[0]) | (_part1by1(coords[1]) << np.uint64(1))
    if ndims == 3:
        return (_part1by2(coords[0]) | (_part1by2(coords[1]) << np.uint64(1))
                | (_part1by2(coords[2]) << np.uint64(2)))
    if ndims == 1:
        return coords[0].astype(np.uint64)
    raise ValueError(ndims)


class TreeMesh:
    """Cartesian hypercube mesh refined as a 2^d-tree (``tree_mesh.jl:128-186``).

    Leaves are stored as ``levels[c]`` and integer coordinates ``icoords[d, c]`` (cell index at the
    cell's own level); ``cell order == reference leaf order``.
    """

    def __init__(self, coordinates_min, coordinates_max, initial_refinement_level,
                 periodicity=False, refinement_patches=(), n_cells_max=None):
        coordinates_min = tuple(float(c) for c in coordinates_min)
        coordinates_max = tuple(float(c) for c in coordinates_max)
        if len(coordinates_min) != len(coordinates_max):
            raise ValueError("coordinates_min and coordinates_max must have the same length")
        if not (isinstance(initial_refinement_level, (int, np.integer)) and initial_refinement_level >= 0):
            raise ValueError("`initial_refinement_level` must be a non-negative integer")
        self.ndims = len(coordinates_min)
        if any(a >= b for a, b in zip(coordinates_min, coordinates_max)):
            raise ValueError("coordinates_max must be larger than coordinates_min")
        self.center_level_0 = np.array([(a + b) / 2 for a, b in zip(coordinates_min, coordinates_max)])
        self.length_level_0 = coordinates_max[0] - coordinates_min[0]
        for d in range(1, self.ndims):
            if not np.isclose(coordinates_max[d] - coordinates_min[d], self.length_level_0):
                raise ValueError("The TreeMesh domain must be a hypercube")
        if isinstance(periodicity, bool):
            periodicity = (periodicity,) * self.ndims
        self.periodicity = tuple(bool(p) for p in periodicity)

        L = int(initial_refinement_level)
        n1 = 1 << L
        grids = np.meshgrid(*[np.arange(n1, dtype=np.int64)] * self.ndims, indexing="ij")
        ic = np.stack([g.ravel() for g in grids])  # [ndims, ncells], arbitrary order
        self.levels = np.full(ic.shape[1], L, dtype=np.int64)
        self.icoords = ic
        self._sort()
        for patch in refinement_patches:
            if patch["type"] != "box":
                raise NotImplementedError("only box refinement patches are supported")
            self.refine_box(patch["coordinates_min"], patch["coordinates_max"])

    # ---- ordering ------------------------------------------------------------------------------
    def _sort(self):
        lmax = int(self.levels.max())
        shift = (lmax - self.levels).astype(np.int64)
        fine = self.icoords << shift  # coordinates of the first descendant at lmax
        key = morton_key(fine, self.ndims)
        order = np.argsort(key, kind="stable")
        self.levels = self.levels[order]
        self.icoords = np.ascontiguousarray(self.icoords[:, order])
        self._lmax = lmax
        self._keys = key[order]

    @property
    def ncells(self):
        return self.levels.shape[0]

    def length_at_level(self, level):
        # abstract_tree.jl:56
        return self.length_level_0 / (1 << np.asarray(level)).astype(np.float64)

    def total_volume(self):
        # tree_mesh.jl:311-313
        return self.length_level_0 ** self.ndims

    def cell_coordinates(self, cells=None):
        """Cell midpoints, accumulated parent->child exactly like ``child_coordinates``
        (abstract_tree.jl:668-674): x_child = x_parent + sign * (parent_length/2) / 2."""
        levels = self.levels if cells is None else self.levels[cells]
        icoords = self.icoords if cells is None else self.icoords[:, cells]
        x = np.repeat(self.center_level_0[:, None], levels.shape[0], axis=1).copy()
        lmax = int(levels.max()) if levels.shape[0] else 0
        for l in range(1, lmax + 1):
            active = levels >= l
            # bit of the ancestor at level l
            bit = (icoords >> np.maximum(levels - l, 0)) & 1
            sign = np.where(bit == 1, 1.0, -1.0)
            child_length = self.length_level_0 / (1 << (l - 1)) / 2
            x = np.where(active[None, :], x + sign * child_length / 2, x)
        return x

    # ---- refinement ----------------------------------------------------------------------------
    def refine_cells(self, mask):
        """Replace every leaf in ``mask`` by its 2^d children, then 2:1-balance
        (``refine!`` abstract_tree.jl:367-403 refines neighbours recursively to keep balance)."""
        while mask.any():
            keep = ~mask
            pl = self.levels[mask]
            pc = self.icoords[:, mask]
            nchild = 1 << self.ndims
            cl = np.repeat(pl + 1, nchild)
            cc = np.repeat(pc * 2, nchild, axis=1)
            for d in range(self.ndims):
                cc[d] += np.tile((np.arange(nchild) >> d) & 1, pl.shape[0])
            self.levels = np.concatenate([self.levels[keep], cl])
            self.icoords = np.concatenate([self.icoords[:, keep], cc], axis=1)
            self._sort()
            mask = self._unbalanced_mask()

    def _unbalanced_mask(self):
        """Leaves that have a face neighbour two or more levels finer."""
        mask = np.zeros(self.ncells, dtype=bool)
        for direction in range(2 * self.ndims):
            nb_kind, nb_idx = self._face_neighbors(direction, want="coarse_of_fine")
            mask |= nb_kind
        return mask

    def refine_box(self, coordinates_min, coordinates_max):
        # refine_box! abstract_tree.jl:405-419: cells whose midpoint lies strictly inside the box
        if any(lo > hi for lo, hi in zip(coordinates_min, coordinates_max)):
            raise ValueError("coordinates_min must not exceed coordinates_max")  # coordinates_min_max_check
        x = self.cell_coordinates()
        inside = np.ones(self.ncells, dtype=bool)
        for d in range(self.ndims):
            inside &= (x[d] > coordinates_min[d]) & (x[d] < coordinates_max[d])
        self.refine_cells(inside)

    # ---- neighbour search -----------------------------------------------------------------------
    def _lookup(self, level, coords):
        """Index of the leaf with exactly (level, coords), or -1."""
        shift = (self._lmax - level).astype(np.int64)
        key = morton_key(coords << shift, self.ndims)
        pos = np.searchsorted(self._keys, key)
        pos_c = np.minimum(pos, self.ncells - 1)
        ok = (pos < self.ncells) & (self._keys[pos_c] == key) & (self.levels[pos_c] == level)
        return np.where(ok, pos_c, -1)

    def _shifted(self, direction, cells=None):
        d = direction // 2
        step = -1 if direction % 2 == 0 else 1
        levels = self.levels if cells is None else self.levels[cells]
        n_at_level = np.int64(1) << levels
        c = self.icoords.copy() if cells is None else self.icoords[:, cells].copy()
        c[d] += step
        outside = (c[d] < 0) | (c[d] >= n_at_level)
        if self.periodicity[d]:
            c[d] = np.mod(c[d], n_at_level)
            outside = np.zeros_like(outside)
        return c, outside, levels

    def _face_neighbors(self, direction, want, cells=None):
        """Global index of the face neighbour of every cell (or of the subset ``cells``) in
        ``direction`` (0-based: -x,+x,-y,...), -1 if there is none of the wanted kind."""
        c, outside, levels = self._shifted(direction, cells)
        c_safe = np.where(outside[None, :], 0, c)
        if want == "same":
            idx = self._lookup(levels, c_safe)
            return np.where(outside, -1, idx)
        if want == "coarse":
            lv = np.maximum(levels - 1, 0)
            idx = self._lookup(lv, c_safe >> 1)
            idx = np.where((levels == 0) | outside, -1, idx)
            return idx
        if want == "coarse_of_fine":
            # does a leaf two levels finer touch this face?  check the 2^(d-1) level+1 neighbour
            # slots: if such a slot is neither a leaf at level+1 nor covered by a coarser/same leaf,
            # it is refined further -> imbalance.
            d = direction // 2
            bad = np.zeros(levels.shape[0], dtype=bool)
            # covered: the neighbour slot is a leaf at this level or lies inside a coarser leaf (any number of levels
            # coarser: with nested refinement patches a cell can temporarily sit next to a leaf two levels coarser;
            # that imbalance is the coarse leaf's to resolve, not this cell's)
            covered = (self._lookup(levels, c_safe) >= 0) | outside
            for k in range(1, int(levels.max()) + 1):
                anc = self._lookup(np.maximum(levels - k, 0), c_safe >> k)
                covered |= (anc >= 0) & (levels >= k)
            others = [e for e in range(self.ndims) if e != d]
            for sub in range(1 << (self.ndims - 1)):
                cc = c_safe * 2
                # the face of the neighbour that touches us: if we step +, its low side
                cc[d] += 0 if direction % 2 == 1 else 1
                for b, e in enumerate(others):
                    cc[e] += (sub >> b) & 1
                lv1 = levels + 1
                ok_lv = lv1 <= self._lmax
                fine = self._lookup(np.minimum(lv1, self._lmax), np.where(ok_lv[None, :], cc, 0))
                fine = np.where(ok_lv, fine, -1)
                # slot exists at level+1 as a leaf -> fine; otherwise, if not covered, deeper
                deeper = (~covered) & (fine < 0)
                bad |= deeper
            return bad, None
        raise ValueError(want)

    def _fine_neighbors(self, direction, cells=None):
        """The 2^(d-1) leaves one level finer that touch the face ``direction`` of every cell, as
        [2^(d-1), ncells] (position = bits of the tangential coordinates, lowest dimension first: the
        lower/upper and left/right of containers_3d.jl:686-715), or -1 where the same-level neighbour cell is
        not refined."""
        d = direction // 2
        c, outside, levels = self._shifted(direction, cells)
        c_safe = np.where(outside[None, :], 0, c)
        others = [e for e in range(self.ndims) if e != d]
        out = np.empty((1 << (self.ndims - 1), levels.shape[0]), dtype=np.int64)
        lv1 = levels + 1
        ok_lv = (lv1 <= self._lmax) & ~outside
        for sub in range(1 << (self.ndims - 1)):
            cc = c_safe * 2
            cc[d] += 0 if direction % 2 == 1 else 1  # the children of the neighbour that touch this cell
            for b, e in enumerate(others):
                cc[e] += (sub >> b) & 1
            fine = self._lookup(np.minimum(lv1, self._lmax), np.where(ok_lv[None, :], cc, 0))
            out[sub] = np.where(ok_lv, fine, -1)
        return out

    def _has_refined_neighbor(self, direction, cells=None):
        if int(self.levels.min()) == int(self.levels.max()):  # uniform mesh (also CartesianBoxMesh)
            n = self.ncells if cells is None else len(cells)
            return np.zeros(n, dtype=bool)
        return (self._fine_neighbors(direction, cells) >= 0).all(axis=0)

    def __repr__(self):
        return f"TreeMesh{{{self.ndims}}} with {self.ncells} leaf cells"


class CartesianBoxMesh(TreeMesh):
    """Synthetic uniform Cartesian connectivity with an arbitrary number of cubic cells per direction
    (100^3 = 64 M DOF, 200^3 = 512 M DOF and the weak-scaling boxes 256x128x128, ... are
    not TreeMesh levels).  Cells are ordered along the same Morton curve as TreeMesh leaves, so a box of
    2^L cells per direction reproduces the TreeMesh of level L exactly and contiguous chunks of the
    ordering are compact blocks.  Everything downstream consumes the same Tree containers."""

    def __init__(self, coordinates_min, cell_length, cells_per_dimension, periodicity=True):
        self.ndims = len(coordinates_min)
        self.cells_per_dimension = tuple(int(c) for c in cells_per_dimension)
        if len(self.cells_per_dimension) != self.ndims:
            raise ValueError("cells_per_dimension must have one entry per dimension")
        self.coordinates_min = np.array(coordinates_min, dtype=np.float64)
        self.dx = float(cell_length)
        self.domain_lengths = self.dx * np.array(self.cells_per_dimension, dtype=np.float64)
        self.length_level_0 = float(self.domain_lengths.max())
        self.center_level_0 = self.coordinates_min + 0.5 * self.domain_lengths
        if isinstance(periodicity, bool):
            periodicity = (periodicity,) * self.ndims
        self.periodicity = tuple(periodicity)
        n = self.cells_per_dimension
        grids = np.meshgrid(*[np.arange(m, dtype=np.int64) for m in n], indexing="ij")
        ic = np.stack([g.ravel() for g in grids])
        key = morton_key(ic, self.ndims)
        order = np.argsort(key, kind="stable")
        self.icoords = np.ascontiguousarray(ic[:, order])
        self._keys = key[order]
        self.levels = np.zeros(self.icoords.shape[1], dtype=np.int64)

    def total_volume(self):
        return float(np.prod(self.domain_lengths))

    def length_at_level(self, level):
        return np.full(np.shape(level), self.dx)

    def cell_coordinates(self, cells=None):
        ic = self.icoords if cells is None else self.icoords[:, cells]
        return self.coordinates_min[:, None] + (ic + 0.5) * self.dx

    def _face_neighbors(self, direction, want, cells=None):
        ncells = self.ncells if cells is None else len(cells)
        if want != "same":
            return np.full(ncells, -1, dtype=np.int64)
        d = direction // 2
        step = -1 if direction % 2 == 0 else 1
        n = self.cells_per_dimension
        c = self.icoords.copy() if cells is None else self.icoords[:, cells].copy()
        c[d] += step
        outside = (c[d] < 0) | (c[d] >= n[d])
        if self.periodicity[d]:
            c[d] %= n[d]
            outside[:] = False
        c[d] = np.where(outside, 0, c[d])
        key = morton_key(c, self.ndims)
        pos = np.searchsorted(self._keys, key)
        return np.where(outside, -1, pos)
