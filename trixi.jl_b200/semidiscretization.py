"""``SemidiscretizationHyperbolic`` and the ``rhs_hyperbolic!`` wrapper.

Mirrors ``src/semidiscretization/semidiscretization_hyperbolic.jl:14-30,50-76,578-597`` and
``semidiscretization.jl:102-153,224-242``.  The constructor builds the containers on the host
(``create_cache``) and the C-ABI descriptor; the device handle is created on first use -- the
analogue of ``semidiscretize(...; storage_type=CuArray)`` adapting all containers once
(``semidiscretization.jl:115-126``).  There is no CPU fallback: every RHS evaluation goes through
``libtrixi_b200.so``.
"""
from __future__ import annotations

import time

import numpy as np

from . import _abi
from .containers import (BoundaryContainer, InterfaceContainer, MPIInterfaceContainer, init_boundaries,
                         init_elements, init_interfaces, init_mortars, init_mpi_mortars, partition_cells)
from .equations import (BC_DIRICHLET, BC_PERIODIC, BC_SLIP_WALL, IC_NONE, SRC_NONE,
                        BoundaryConditionDirichlet, boundary_condition_periodic, resolve_flux)
from .basis import LobattoLegendreMortarL2
from .mesh import TreeMesh
from .p4est import (P4estMesh, init_boundaries_p4est, init_elements_p4est, init_interfaces_p4est,
                    init_mortars_p4est, init_mpi_mortars_p4est)
from .structured import StructuredMesh, init_elements_structured

MESH_TREE, MESH_STRUCTURED, MESH_P4EST = 0, 1, 2

_DIRECTION_NAMES = ("x_neg", "x_pos", "y_neg", "y_pos", "z_neg", "z_pos")


class PerformanceCounter:
    """``PerformanceCounter`` (auxiliary/auxiliary.jl:22-41): accumulates synchronised RHS run time."""

    def __init__(self):
        self.ncalls_since_readout = 0
        self.runtime = 0.0

    def put(self, runtime_ns):
        self.ncalls_since_readout += 1
        self.runtime += runtime_ns

    def take(self):
        r = self.runtime / max(self.ncalls_since_readout, 1)
        self.ncalls_since_readout = 0
        self.runtime = 0.0
        return r


class Cache:
    pass


def create_cache(mesh, equations, solver, rank=0, world_size=1):
    """``create_cache`` for TreeMesh (dgsem_tree/dg_2d.jl:14-37; with ``world_size > 1`` the containers
    of this rank's contiguous chunk of the element order plus its MPI interfaces,
    dgsem_tree/dg_2d_parallel.jl:245-281)."""
    cache = Cache()
    if isinstance(mesh, TreeMesh):
        first, last = partition_cells(mesh.ncells, rank, world_size)
        cache.first_element, cache.last_element = first, last
        cells = None if world_size == 1 else np.arange(first, last, dtype=np.int64)
        cache.elements = init_elements(mesh, solver.basis, cells)
        cache.interfaces, cache.mpi_interfaces = init_interfaces(mesh, first, last, world_size)
        cache.boundaries = init_boundaries(mesh, cache.elements, solver.basis, first, last)
        cache.mortars = init_mortars(mesh, first, last)
        cache.mpi_mortars = init_mpi_mortars(mesh, cache.mpi_interfaces, first, last, world_size)
    elif isinstance(mesh, StructuredMesh):
        if world_size != 1:
            raise NotImplementedError("StructuredMesh runs on a single rank (the reference has no MPI path for it)")
        # create_cache dgsem_structured/dg.jl:11-25
        cache.first_element, cache.last_element = 0, mesh.ncells
        cache.elements = init_elements_structured(mesh, solver.basis)
        cache.interfaces = InterfaceContainer()
        cache.interfaces.ninterfaces = 0  # faces are found through left_neighbors (dg_3d.jl:657-689)
        cache.interfaces.neighbor_ids = np.zeros((2, 0), dtype=np.int64)
        cache.interfaces.orientations = np.zeros(0, dtype=np.int64)
        cache.mpi_interfaces = MPIInterfaceContainer()
        cache.mpi_interfaces.nmpiinterfaces = 0
        cache.boundaries = _structured_boundaries(mesh, cache.elements)
    elif isinstance(mesh, P4estMesh):
        # create_cache dgsem_p4est/dg.jl:13-70 (+ dg_parallel.jl:267-311 for a partition); the forest is balanced
        # first "in case someone has tampered with the p4est after creating the mesh" (dg.jl:13-16)
        if not mesh.is_uniform:
            mesh.balance()
        first, last = partition_cells(mesh.ncells, rank, world_size)
        cache.first_element, cache.last_element = first, last
        cache.elements = init_elements_p4est(mesh, solver.basis, first, last)
        cache.interfaces, cache.mpi_interfaces = init_interfaces_p4est(mesh, first, last, world_size)
        cache.boundaries = init_boundaries_p4est(mesh, first, last, world_size)
        cache.mortars = init_mortars_p4est(mesh, first, last, world_size)
        cache.mpi_mortars = init_mpi_mortars_p4est(mesh, solver.basis, first, last, world_size)
    else:
        raise TypeError(f"unsupported mesh type {type(mesh).__name__}")
    return cache


def _structured_boundaries(mesh, elements):
    """Domain-boundary faces of a StructuredMesh as a direction-sorted list (the reference loops over the
    boundary cells of each direction, dgsem_structured/dg_3d.jl:755-935)."""
    nd = mesh.ndims
    cells = mesh.cells_per_dimension
    lin = np.arange(1, mesh.ncells + 1, dtype=np.int64).reshape(cells, order="F")
    ids, counts = [], []
    for direction in range(2 * nd):
        d = direction // 2
        if mesh.periodicity[d]:
            counts.append(0)
            continue
        idx = [slice(None)] * nd
        idx[d] = 0 if direction % 2 == 0 else cells[d] - 1
        el = lin[tuple(idx)].ravel(order="F")
        ids.append(el)
        counts.append(el.shape[0])
    bc = BoundaryContainer()
    bc.neighbor_ids = np.concatenate(ids).astype(np.int64) if ids else np.zeros(0, dtype=np.int64)
    orient, side = [], []
    for direction, c in enumerate(counts):
        orient += [direction // 2 + 1] * c
        side += [2 if direction % 2 == 0 else 1] * c
    bc.orientations = np.array(orient, dtype=np.int64)
    bc.neighbor_sides = np.array(side, dtype=np.int64)
    bc.node_coordinates = np.zeros((nd, 0))
    bc.n_boundaries_per_direction = np.array(counts + [0] * (6 - len(counts)), dtype=np.int64)
    bc.nboundaries = int(bc.neighbor_ids.shape[0])
    return bc


def _digest_boundary_conditions(boundary_conditions, mesh):
    """semidiscretization_hyperbolic.jl:115-206: a single BC applies to all directions; a dict keyed
    x_neg, x_pos, ... gives one per direction; periodic BCs must match the mesh periodicity."""
    nd = mesh.ndims
    if isinstance(boundary_conditions, dict):
        names = _DIRECTION_NAMES[:2 * nd]
        if set(boundary_conditions) != set(names):
            raise ValueError(f"boundary_conditions must have exactly the keys {names}")
        bcs = [boundary_conditions[k] for k in names]
    elif isinstance(boundary_conditions, (tuple, list)):
        if len(boundary_conditions) != 2 * nd:
            raise ValueError("need one boundary condition per direction")
        bcs = list(boundary_conditions)
    else:
        bcs = [boundary_conditions] * (2 * nd)
    tags, ics = [0] * 6, [0] * 6
    for i, bc in enumerate(bcs):
        periodic_dim = mesh.periodicity[i // 2]
        if bc is boundary_condition_periodic:
            if not periodic_dim:
                raise ValueError(f"boundary_condition_periodic in non-periodic direction {_DIRECTION_NAMES[i]}")
            tags[i] = BC_PERIODIC
        else:
            if periodic_dim:
                raise ValueError(f"non-periodic boundary condition in periodic direction {_DIRECTION_NAMES[i]}")
            if isinstance(bc, BoundaryConditionDirichlet):
                tags[i], ics[i] = BC_DIRICHLET, bc.ic_id
            elif getattr(bc, "tag", None) == BC_SLIP_WALL:
                tags[i] = BC_SLIP_WALL
            else:
                raise TypeError(f"boundary condition {bc!r} is not in the libtrixi_b200 registry")
    return tags, ics


class SemidiscretizationHyperbolic:
    """``SemidiscretizationHyperbolic(mesh, equations, initial_condition, solver; source_terms,
    boundary_conditions)`` (semidiscretization_hyperbolic.jl:50-76)."""

    def __init__(self, mesh, equations, initial_condition, solver, source_terms=None,
                 boundary_conditions=boundary_condition_periodic, device=-1, rank=0, world_size=1, comm=None):
        if mesh.ndims != equations.ndims:
            raise ValueError("mesh and equations must have the same number of dimensions")
        self.mesh, self.equations, self.solver = mesh, equations, solver
        self.initial_condition = initial_condition
        self.source_terms = source_terms
        self.boundary_conditions = boundary_conditions
        self.rank, self.world_size = int(rank), int(world_size)
        self.comm = comm  # torch.distributed (plumbing: connection blobs, dt / error-norm reductions)
        self.cache = create_cache(mesh, equations, solver, self.rank, self.world_size)
        self.performance_counter = PerformanceCounter()
        self.device = device
        self._bc_tags, self._bc_ics = _digest_boundary_conditions(boundary_conditions, mesh)
        self._desc = None
        self._backend = None

    # ---- sizes -----------------------------------------------------------------------------------
    @property
    def nelements(self):
        return self.cache.elements.nelements

    def ndofs(self):
        """Number of DOFs = nodes (docs/src/performance.md:201-218, solvers/dg.jl:969-971); local to
        this rank."""
        return self.nelements * self.solver.nnodes ** self.mesh.ndims

    def ndofsglobal(self):
        return self.mesh.ncells * self.solver.nnodes ** self.mesh.ndims

    @property
    def is_curved(self):
        return isinstance(self.mesh, (StructuredMesh, P4estMesh))

    def u_shape(self):
        return (self.equations.nvars,) + (self.solver.nnodes,) * self.mesh.ndims + (self.nelements,)

    def u_length(self):
        return int(np.prod(self.u_shape()))

    # ---- descriptor -------------------------------------------------------------------------------
    def descriptor(self):
        if self._desc is not None:
            return self._desc
        h = _abi.DescHolder()
        d = h.desc
        eq, dg, cache = self.equations, self.solver, self.cache
        d.abi_version = _abi.ABI_VERSION
        d.device = self.device
        d.ndims, d.nvars, d.nnodes = self.mesh.ndims, eq.nvars, dg.nnodes
        d.mesh_kind = (MESH_STRUCTURED if isinstance(self.mesh, StructuredMesh)
                       else MESH_P4EST if isinstance(self.mesh, P4estMesh) else MESH_TREE)
        d.nelements = self.nelements
        d.equation = eq.eq_id
        d.volume_integral = dg.volume_integral.kind
        d.volume_flux = resolve_flux(dg.volume_integral.volume_flux)
        d.surface_flux = resolve_flux(dg.surface_integral.surface_flux)
        d.source_terms = SRC_NONE if self.source_terms is None else self.source_terms.tag
        for i in range(6):
            d.boundary_conditions[i] = self._bc_tags[i]
            d.boundary_ic[i] = self._bc_ics[i]
        for i, p in enumerate(eq.params()):
            d.eq_params[i] = p
        h.set_f64("derivative_split", dg.basis.derivative_split)
        h.set_f64("derivative_hat", dg.basis.derivative_hat)
        h.set_f64("inverse_weights", dg.basis.inverse_weights)
        if dg.volume_integral.kind == 3:  # VolumeIntegralPureLGLFiniteVolume
            d.volume_flux_fv = resolve_flux(dg.volume_integral.volume_flux_fv)
            if self.is_curved:
                from .structured import calc_normalvectors_subcell_fv
                if getattr(cache, "normal_vectors", None) is None:
                    cache.normal_vectors = calc_normalvectors_subcell_fv(cache.elements.contravariant_vectors, dg.basis)
                for a, nv in enumerate(cache.normal_vectors):
                    h.set_f64_item("subcell_normal_vectors", a, nv)
        if dg.volume_integral.kind == 2:  # VolumeIntegralShockCapturingHG
            ind = dg.volume_integral.indicator
            d.volume_flux_fv = resolve_flux(dg.volume_integral.volume_flux_fv)
            d.indicator_variable = ind.variable.var_id
            d.indicator_alpha_smooth = int(ind.alpha_smooth)
            d.indicator_alpha_max, d.indicator_alpha_min = ind.alpha_max, ind.alpha_min
            h.set_f64("inverse_vandermonde_legendre", dg.basis.inverse_vandermonde_legendre)
            if self.is_curved:
                # cache.normal_vectors (create_cache for VolumeIntegralShockCapturingHG on curved meshes,
                # dgsem_structured/dg.jl: NormalVectorContainer)
                from .structured import calc_normalvectors_subcell_fv
                if getattr(cache, "normal_vectors", None) is None:
                    cache.normal_vectors = calc_normalvectors_subcell_fv(cache.elements.contravariant_vectors, dg.basis)
                for a, nv in enumerate(cache.normal_vectors):
                    h.set_f64_item("subcell_normal_vectors", a, nv)
        h.set_f64("inverse_jacobian", cache.elements.inverse_jacobian)
        h.set_f64("node_coordinates", cache.elements.node_coordinates)
        if isinstance(self.mesh, StructuredMesh):
            h.set_f64("contravariant_vectors", cache.elements.contravariant_vectors)
            h.set_i64("left_neighbors", cache.elements.left_neighbors)
        if isinstance(self.mesh, P4estMesh):
            h.set_f64("contravariant_vectors", cache.elements.contravariant_vectors)
            h.set_i64("interface_node_indices", cache.interfaces.node_indices)
            if cache.boundaries.nboundaries:
                h.set_i64("boundary_node_indices", cache.boundaries.node_indices)
        d.ninterfaces = cache.interfaces.ninterfaces
        h.set_i64("interface_neighbor_ids", cache.interfaces.neighbor_ids)
        h.set_i64("interface_orientations", cache.interfaces.orientations)
        b = cache.boundaries
        d.nboundaries = b.nboundaries
        if b.nboundaries:
            h.set_i64("boundary_neighbor_ids", b.neighbor_ids)
            h.set_i64("boundary_orientations", b.orientations)
            h.set_i64("boundary_neighbor_sides", b.neighbor_sides)
            if b.node_coordinates.size:
                h.set_f64("boundary_node_coordinates", b.node_coordinates)
        for i in range(6):
            d.n_boundaries_per_direction[i] = int(b.n_boundaries_per_direction[i])
        mortars = getattr(cache, "mortars", None)
        d.nmortars = 0 if mortars is None else mortars.nmortars
        if d.nmortars:
            # L2 mortars (containers_3d.jl:495-510) and their operators (basis_lobatto_legendre.jl:159-206)
            l2 = LobattoLegendreMortarL2(dg.basis)
            h.set_i64("mortar_neighbor_ids", mortars.neighbor_ids)
            if isinstance(self.mesh, P4estMesh):
                h.set_i64("mortar_node_indices", mortars.node_indices)
            else:
                h.set_i64("mortar_large_sides", mortars.large_sides)
                h.set_i64("mortar_orientations", mortars.orientations)
            h.set_f64("mortar_forward_upper", l2.forward_upper)
            h.set_f64("mortar_forward_lower", l2.forward_lower)
            h.set_f64("mortar_reverse_upper", l2.reverse_upper)
            h.set_f64("mortar_reverse_lower", l2.reverse_lower)
        d.rank, d.world_size = self.rank, self.world_size
        mi = cache.mpi_interfaces
        d.nmpiinterfaces = mi.nmpiinterfaces
        if mi.nmpiinterfaces:
            h.set_i64("mpi_local_neighbor_ids", mi.local_neighbor_ids)
            h.set_i64("mpi_local_sides", mi.local_sides)
            h.set_i64("mpi_orientations", mi.orientations)
            h.set_i64("mpi_neighbor_ranks", mi.neighbor_ranks)
            if isinstance(self.mesh, P4estMesh):
                h.set_i64("mpi_node_indices", mi.node_indices)
            if getattr(mi, "is_mortar_piece", None) is not None and np.any(mi.is_mortar_piece):
                h.set_i64("mpi_is_mortar_piece", mi.is_mortar_piece)
        mm = getattr(cache, "mpi_mortars", None)
        d.nmpimortars = 0 if mm is None else mm.nmpimortars
        if d.nmpimortars:
            l2 = LobattoLegendreMortarL2(dg.basis)
            h.set_i64("mpi_mortar_neighbor_ids", mm.neighbor_ids)
            if isinstance(self.mesh, P4estMesh):
                h.set_i64("mpi_mortar_node_indices", mm.node_indices)
                h.set_f64("mpi_mortar_normal_directions", mm.normal_directions)
            else:
                h.set_i64("mpi_mortar_large_sides", mm.large_sides)
                h.set_i64("mpi_mortar_orientations", mm.orientations)
            if not d.nmortars:  # the operators travel with the local mortar fields
                h.set_f64("mortar_forward_upper", l2.forward_upper)
                h.set_f64("mortar_forward_lower", l2.forward_lower)
                h.set_f64("mortar_reverse_upper", l2.reverse_upper)
                h.set_f64("mortar_reverse_lower", l2.reverse_lower)
        self._desc = h
        return h

    # ---- device backend ---------------------------------------------------------------------------
    def backend(self):
        """The B200 backend object (``trixi_backend(u)`` analogue, auxiliary/containers.jl:278-285)."""
        if self._backend is None:
            from .lib import B200Backend
            self._backend = B200Backend(self.descriptor(), self.u_length())
            if self.world_size > 1 and self.comm is not None:
                self._backend.connect(self.comm)
        return self._backend

    def set_backend(self, backend):
        """Tests inject the CPU oracle here; the product never does."""
        self._backend = backend


def mesh_equations_solver_cache(semi):
    return semi.mesh, semi.equations, semi.solver, semi.cache


def compute_coefficients(t, semi):
    """``compute_coefficients(t, semi)`` (semidiscretization.jl:224-242): u0 = initial_condition(x, t)
    at every node, shape [nvars, n, n, (n,) nelements] Fortran-ordered (solvers/dg.jl:1200-1210)."""
    x = semi.cache.elements.node_coordinates
    u = semi.initial_condition(x, t, semi.equations)
    return np.asfortranarray(u, dtype=np.float64)


class ODEProblem:
    def __init__(self, f, u0, tspan, p):
        self.f, self.u0, self.tspan, self.p = f, u0, tspan, p


def semidiscretize(semi, tspan):
    """``semidiscretize(semi, tspan)`` (semidiscretization.jl:102-153)."""
    u0 = compute_coefficients(tspan[0], semi)
    return ODEProblem(rhs_hyperbolic, u0, (float(tspan[0]), float(tspan[1])), semi)


def rhs_hyperbolic(du_ode, u_ode, semi, t):
    """``rhs_hyperbolic!(du_ode, u_ode, semi, t)`` (semidiscretization_hyperbolic.jl:578-597): host
    buffers in and out; the run time (synchronised, copies included) feeds the PerformanceCounter."""
    t0 = time.perf_counter_ns()
    semi.backend().rhs_host(du_ode, u_ode, t)
    semi.performance_counter.put(time.perf_counter_ns() - t0)
    return None
