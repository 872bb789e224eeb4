# TrixiB200.jl -- the reference-side binding of libtrixi_b200.so.
#
# Trixi.jl selects its solver kernels by multiple dispatch on a `backend` argument
# (src/solvers/dgsem_tree/dg_2d.jl:113-120 `backend::Nothing`, src/solvers/dgsem_p4est/dg_2d_gpu.jl:8
# `backend::Backend`).  This file adds a backend type `B200` and the methods whose bodies `ccall` into
# the C ABI of include/trixi_b200.h -- the same mechanism as ext/TrixiCUDACoreExt.jl uses to plug in
# CUDA.jl, but without KernelAbstractions.  Julia is not installed in the build container or on the GPU
# box, so this file is exercised only by inspection; its Python twin (trixi.jl_b200/lib.py) binds the
# identical entry points and is what the test-suite drives (tests/c_abi_smoke.c drives them from plain C).
#
# Usage (the only change to an elixir is the `offload` line):
#
#     using Trixi, TrixiB200
#     semi = SemidiscretizationHyperbolic(mesh, equations, initial_condition, solver; ...)
#     ode  = semidiscretize(semi, tspan)
#     ode  = TrixiB200.offload(ode)             # create_cache -> trixi_b200_create; u0 becomes a B200Vector
#     sol  = Trixi.solve(ode, Trixi.CarpenterKennedy2N54(); dt = 1.0, callback = callbacks)
#
# From there on everything dispatches on the array type of the solution vector, exactly like `CuArray` does for
# the reference's KernelAbstractions path: `trixi_backend(u)` returns the `B200` object, `rhs_hyperbolic!`,
# `max_dt` and `step!` below take over, and the callbacks keep seeing a host array.
module TrixiB200

using Trixi
using Trixi: TreeMesh, StructuredMesh, P4estMesh, DG, DGSEM, SemidiscretizationHyperbolic, nvariables, nnodes,
             ndims, nelements, ninterfaces, nboundaries, nmortars, mesh_equations_solver_cache, True, False
import KernelAbstractions
import LoopVectorization
import SciMLBase

const libtrixi_b200 = get(ENV, "TRIXI_B200_LIBRARY", "libtrixi_b200.so")

# ---- enums of include/trixi_b200.h ------------------------------------------------------------------
const ABI_VERSION = Int32(5)
mesh_kind(::TreeMesh) = Cint(0)
mesh_kind(::StructuredMesh) = Cint(1)
mesh_kind(::P4estMesh) = Cint(2)
equation_id(::LinearScalarAdvectionEquation2D) = Cint(1)
equation_id(::LinearScalarAdvectionEquation3D) = Cint(5)
equation_id(::CompressibleEulerEquations2D) = Cint(2)
equation_id(::CompressibleEulerEquations3D) = Cint(3)
equation_id(::IdealGlmMhdEquations3D) = Cint(4)
equation_params(eq::IdealGlmMhdEquations3D) = (eq.gamma, eq.inv_gamma_minus_one, eq.c_h, 0.0, 0.0, 0.0, 0.0, 0.0)
equation_params(eq::LinearScalarAdvectionEquation2D) = (eq.advection_velocity..., 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
equation_params(eq::LinearScalarAdvectionEquation3D) = (eq.advection_velocity..., 0.0, 0.0, 0.0, 0.0, 0.0)
equation_params(eq::Union{CompressibleEulerEquations2D, CompressibleEulerEquations3D}) = (eq.gamma,
                                                                                          eq.inv_gamma_minus_one,
                                                                                          0.0, 0.0, 0.0, 0.0, 0.0,
                                                                                          0.0)
volume_integral_id(::VolumeIntegralWeakForm) = (Cint(0), Cint(0))
volume_integral_id(v::VolumeIntegralFluxDifferencing) = (Cint(1), flux_id(v.volume_flux))
# VolumeIntegralShockCapturingHG (solvers/dg.jl): volume_flux = volume_flux_dg; the FV flux and the indicator
# travel in the descriptor's trailing fields
volume_integral_id(v::VolumeIntegralShockCapturingHG) = (Cint(2), flux_id(v.volume_flux_dg))
volume_integral_id(v::VolumeIntegralPureLGLFiniteVolume) = (Cint(3), flux_id(v.volume_flux_fv))
indicator_variable_id(::typeof(density_pressure)) = Cint(0)
indicator_variable_id(::typeof(density)) = Cint(1)
indicator_variable_id(::typeof(pressure)) = Cint(2)
shock_capturing_fields(v, basis) = (Cint(0), Cint(0), Cint(0), 0.0, 0.0, Float64[])
# VolumeIntegralPureLGLFiniteVolume (solvers/dg.jl:559-583): only the subcell finite-volume flux
shock_capturing_fields(v::VolumeIntegralPureLGLFiniteVolume, basis) = (flux_id(v.volume_flux_fv), Cint(0), Cint(0), 0.0, 0.0,
                                                                       Float64[])
function shock_capturing_fields(v::VolumeIntegralShockCapturingHG, basis)
    ind = v.indicator
    return (flux_id(v.volume_flux_fv), indicator_variable_id(ind.variable), Cint(ind.alpha_smooth),
            Float64(ind.alpha_max), Float64(ind.alpha_min), Matrix(basis.inverse_vandermonde_legendre))
end
flux_id(::typeof(flux_central)) = Cint(0)
flux_id(::typeof(flux_ranocha)) = Cint(1)
flux_id(f::FluxLaxFriedrichs) = f.dissipation.max_abs_speed === max_abs_speed_naive ? Cint(3) : Cint(2)
# FluxHLL(min_max_speed_davis) = flux_hll, FluxHLL(min_max_speed_naive), FluxHLL(min_max_speed_einfeldt) = flux_hlle
flux_id(f::FluxHLL) = f.min_max_speed === min_max_speed_naive ? Cint(5) :
                      (f.min_max_speed === Trixi.min_max_speed_einfeldt ? Cint(17) : Cint(4))
flux_id(::typeof(flux_shima_etal)) = Cint(6)
flux_id(::typeof(flux_kennedy_gruber)) = Cint(7)
flux_id(::typeof(flux_chandrashekar)) = Cint(8)
flux_id(::typeof(flux_godunov)) = Cint(10)
flux_id(::typeof(flux_hllc)) = Cint(18)
flux_id(::typeof(flux_hindenlang_gassner)) = Cint(9)
flux_id(::typeof(Trixi.flux_ranocha_turbo)) = Cint(11)
# (conservative, nonconservative) tuples of the GLM-MHD elixirs (elixir_mhd_ec.jl:13-17)
flux_id(f::Tuple{typeof(flux_hindenlang_gassner), typeof(flux_nonconservative_powell)}) = Cint(13)
flux_id(f::Tuple{FluxLaxFriedrichs, typeof(flux_nonconservative_powell)}) =
    f[1].dissipation.max_abs_speed === max_abs_speed_naive ? Cint(14) : Cint(12)
flux_id(f::Tuple{FluxHLL, typeof(flux_nonconservative_powell)}) =
    f[1].min_max_speed === Trixi.min_max_speed_einfeldt ? Cint(15) :
    error("FluxHLL with the Powell term: only min_max_speed_einfeldt (flux_hlle) is in the libtrixi_b200 registry")
flux_id(f::Tuple{typeof(flux_central), typeof(flux_nonconservative_powell)}) = Cint(16)
flux_id(f) = error("numerical flux $f is not in the libtrixi_b200 registry")
source_id(::Nothing) = Cint(0)
source_id(::typeof(source_terms_convergence_test)) = Cint(1)
source_id(::typeof(Trixi.source_terms_eoc_test_euler)) = Cint(2)
source_id(::typeof(Trixi.source_terms_eoc_test_coupled_euler_gravity)) = Cint(3)
source_id(f) = error("source term $f is not in the libtrixi_b200 registry")
ic_id(::typeof(initial_condition_constant)) = Cint(1)
ic_id(::typeof(initial_condition_convergence_test)) = Cint(2)
ic_id(::typeof(initial_condition_weak_blast_wave)) = Cint(3)
ic_id(::typeof(Trixi.initial_condition_eoc_test_coupled_euler_gravity)) = Cint(4)
ic_id(f) = error("initial condition $f is not in the libtrixi_b200 registry of Dirichlet boundary states")
bc_id(::Trixi.BoundaryConditionPeriodic) = (Cint(0), Cint(0))
bc_id(bc::BoundaryConditionDirichlet) = (Cint(1), ic_id(bc.boundary_value_function))
bc_id(::typeof(boundary_condition_slip_wall)) = (Cint(2), Cint(0))

# ---- struct trixi_b200_desc (field order and types exactly as in the header) ---------------------------
struct Desc
    abi_version::Int32
    device::Int32
    ndims::Int32
    nvars::Int32
    nnodes::Int32
    mesh_kind::Int32
    nelements::Int64
    equation::Int32
    volume_integral::Int32
    volume_flux::Int32
    surface_flux::Int32
    source_terms::Int32
    boundary_conditions::NTuple{6, Int32}
    boundary_ic::NTuple{6, Int32}
    reserved0::Int32
    eq_params::NTuple{8, Float64}
    derivative_split::Ptr{Float64}
    derivative_hat::Ptr{Float64}
    inverse_weights::Ptr{Float64}
    inverse_jacobian::Ptr{Float64}
    node_coordinates::Ptr{Float64}
    contravariant_vectors::Ptr{Float64}
    ninterfaces::Int64
    interface_neighbor_ids::Ptr{Int64}
    interface_orientations::Ptr{Int64}
    interface_node_indices::Ptr{Int64}
    nboundaries::Int64
    boundary_neighbor_ids::Ptr{Int64}
    boundary_orientations::Ptr{Int64}
    boundary_neighbor_sides::Ptr{Int64}
    boundary_node_coordinates::Ptr{Float64}
    n_boundaries_per_direction::NTuple{6, Int64}
    nmortars::Int64
    mortar_neighbor_ids::Ptr{Int64}
    mortar_large_sides::Ptr{Int64}
    mortar_orientations::Ptr{Int64}
    mortar_forward_upper::Ptr{Float64}
    mortar_forward_lower::Ptr{Float64}
    mortar_reverse_upper::Ptr{Float64}
    mortar_reverse_lower::Ptr{Float64}
    left_neighbors::Ptr{Int64}
    rank::Int32
    world_size::Int32
    nmpiinterfaces::Int64
    mpi_local_neighbor_ids::Ptr{Int64}
    mpi_local_sides::Ptr{Int64}
    mpi_orientations::Ptr{Int64}
    mpi_neighbor_ranks::Ptr{Int64}
    boundary_node_indices::Ptr{Int64}
    mpi_node_indices::Ptr{Int64}
    volume_flux_fv::Int32
    indicator_variable::Int32
    indicator_alpha_smooth::Int32
    reserved1::Int32
    indicator_alpha_max::Float64
    indicator_alpha_min::Float64
    inverse_vandermonde_legendre::Ptr{Float64}
    mortar_node_indices::Ptr{Int64}
    nmpimortars::Int64
    mpi_mortar_neighbor_ids::Ptr{Int64}
    mpi_mortar_large_sides::Ptr{Int64}
    mpi_mortar_orientations::Ptr{Int64}
    mpi_mortar_node_indices::Ptr{Int64}
    mpi_mortar_normal_directions::Ptr{Float64}
    mpi_is_mortar_piece::Ptr{Int64}
    subcell_normal_vectors::NTuple{3, Ptr{Float64}}
end

# ---- the backend object ---------------------------------------------------------------------------------
# (deliberately NOT a subtype of KernelAbstractions.Backend: the reference's `rhs_hyperbolic!(backend::Backend, ...)`
# methods for P4estMesh, dgsem_p4est/dg_2d_gpu.jl:8-73, would be ambiguous with the one below)
mutable struct B200
    handle::Ptr{Cvoid}
    ulength::Int
    device_u_valid::Bool   # the handle's device-resident u equals the integrator's host u
    function B200(handle, ulength)
        b = new(handle, ulength, false)
        finalizer(x -> ccall((:trixi_b200_destroy, libtrixi_b200), Cvoid, (Ptr{Cvoid},), x.handle), b)
        return b
    end
end

function check(b, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:trixi_b200_last_error, libtrixi_b200), Cstring, (Ptr{Cvoid},),
                              b === nothing ? C_NULL : b.handle))
    error("libtrixi_b200 error $rc: $msg")   # there is no CPU fallback by design
end

# node_indices tuples of Symbols (dgsem_p4est/containers.jl:226-252) -> the integer code of the header
const INDEX_CODE = Dict(:begin => 0, :end => 1, :i_forward => 2, :i_backward => 3, :j_forward => 4,
                        :j_backward => 5)
encode(node_indices::AbstractArray{<:NTuple{N, Symbol}}) where {N} =
    Int64[INDEX_CODE[t[d]] for d in 1:N, t in node_indices]   # [ndims, size(node_indices)...]
ptr_or_null(a) = isempty(a) ? Ptr{eltype(a)}(C_NULL) : pointer(a)

# Called once after `create_cache` (dgsem_tree/dg_2d.jl:14-37, dgsem_structured/dg.jl:11-25,
# dgsem_p4est/dg.jl:13-70); the analogue of `semidiscretize(...; storage_type = CuArray)` adapting all
# containers (semidiscretization.jl:115-126).  Every array is handed over in the reference's own layout.
function B200(semi::SemidiscretizationHyperbolic; device = -1)
    mesh, equations, dg, cache = mesh_equations_solver_cache(semi)
    basis = dg.basis
    volint, volflux = volume_integral_id(dg.volume_integral)
    bcs = semi.boundary_conditions isa NamedTuple ? values(semi.boundary_conditions) :
          ntuple(_ -> semi.boundary_conditions, 2 * ndims(mesh))
    bc_tags = ntuple(i -> i <= length(bcs) ? bc_id(bcs[i])[1] : Cint(0), 6)
    bc_ics = ntuple(i -> i <= length(bcs) ? bc_id(bcs[i])[2] : Cint(0), 6)
    el = cache.elements
    D_split, D_hat = Matrix(basis.derivative_split), Matrix(basis.derivative_hat)
    inv_w = collect(basis.inverse_weights)
    # geometry and connectivity per mesh type
    contravariant = mesh isa TreeMesh ? Float64[] : el.contravariant_vectors
    left_neighbors = mesh isa StructuredMesh ? el.left_neighbors : Int64[]
    if mesh isa StructuredMesh   # faces are found through left_neighbors (dgsem_structured/dg_3d.jl:657-689)
        if_ids, if_orient, if_idx, n_if = Int64[], Int64[], Int64[], 0
        # domain-boundary faces as a direction-sorted list (the reference loops over the boundary cells of each
        # direction, dgsem_structured/dg_3d.jl:755-935)
        lin = LinearIndices(size(mesh))
        bd_ids, bd_orient, bd_sides, counts = Int64[], Int64[], Int64[], Int64[]
        for direction in 1:(2 * ndims(mesh))
            d = (direction + 1) ÷ 2
            if Trixi.isperiodic(mesh, d)
                push!(counts, 0)
                continue
            end
            cells = vec(selectdim(lin, d, isodd(direction) ? 1 : size(mesh, d)))
            append!(bd_ids, cells)
            append!(bd_orient, fill(d, length(cells)))
            append!(bd_sides, fill(isodd(direction) ? 2 : 1, length(cells)))
            push!(counts, length(cells))
        end
        bd_x, bd_idx, n_bd = Float64[], Int64[], length(bd_ids)
        nbd = ntuple(i -> i <= length(counts) ? counts[i] : Int64(0), 6)
    else
        ifc, bd = cache.interfaces, cache.boundaries
        if_ids, n_if = ifc.neighbor_ids, ninterfaces(dg, cache)
        bd_ids, n_bd = bd.neighbor_ids, nboundaries(dg, cache)
        if mesh isa P4estMesh
            if_orient, if_idx = Int64[], encode(ifc.node_indices)
            bd_orient, bd_sides, bd_x, bd_idx = Int64[], Int64[], Float64[], encode(bd.node_indices)
            # boundaries sorted by name = the order of the boundary-condition container
            # (UnstructuredSortedBoundaryTypes, dgsem_unstructured/sort_boundary_conditions.jl)
            counts = length.(semi.boundary_conditions.boundary_indices)
            nbd = ntuple(i -> i <= length(counts) ? Int64(counts[i]) : Int64(0), 6)
        else
            if_orient, if_idx = ifc.orientations, Int64[]
            bd_orient, bd_sides, bd_x, bd_idx = bd.orientations, bd.neighbor_sides, bd.node_coordinates, Int64[]
            nbd = ntuple(i -> i <= 2 * ndims(mesh) ? Int64(bd.n_boundaries_per_direction[i]) : Int64(0), 6)
        end
    end
    # L2 mortars (TreeMesh containers_3d.jl:495-510; P4estMesh dgsem_p4est/containers.jl:563-613; operators
    # basis_lobatto_legendre.jl:159-206)
    n_mo = mesh isa StructuredMesh ? 0 : nmortars(dg, cache)
    mo_ids = n_mo > 0 ? cache.mortars.neighbor_ids : Int64[]
    mo_sides = n_mo > 0 && mesh isa TreeMesh ? cache.mortars.large_sides : Int64[]
    mo_orient = n_mo > 0 && mesh isa TreeMesh ? cache.mortars.orientations : Int64[]
    mo_idx = n_mo > 0 && mesh isa P4estMesh ? encode(cache.mortars.node_indices) : Int64[]
    fu, fl, ru, rl = n_mo > 0 ? Matrix.((dg.mortar.forward_upper, dg.mortar.forward_lower,
                                         dg.mortar.reverse_upper, dg.mortar.reverse_lower)) :
                     ntuple(_ -> zeros(0, 0), 4)
    fv_flux, ind_var, ind_smooth, ind_max, ind_min, inv_vdm = shock_capturing_fields(dg.volume_integral, dg.basis)
    # shock capturing on curved meshes: cache.normal_vectors (NormalVectorContainer, dgsem_structured/containers_3d.jl:488-541)
    nvc = haskey(cache, :normal_vectors) ? cache.normal_vectors : nothing
    subcell_normals = nvc === nothing ? ntuple(_ -> Ptr{Float64}(C_NULL), 3) :
                      (pointer(nvc.normal_vectors_1), pointer(nvc.normal_vectors_2),
                       ndims(mesh) == 3 ? pointer(nvc.normal_vectors_3) : Ptr{Float64}(C_NULL))
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    # the library copies during `create` only
    GC.@preserve nvc inv_vdm D_split D_hat inv_w el contravariant left_neighbors if_ids if_orient if_idx bd_ids bd_orient bd_sides bd_x bd_idx mo_ids mo_sides mo_orient mo_idx fu fl ru rl begin
        desc = Desc(ABI_VERSION, device, ndims(mesh), nvariables(equations), nnodes(dg), mesh_kind(mesh),
                    nelements(dg, cache), equation_id(equations), volint, volflux,
                    flux_id(dg.surface_integral.surface_flux), source_id(semi.source_terms),
                    bc_tags, bc_ics, 0, equation_params(equations),
                    pointer(D_split), pointer(D_hat), pointer(inv_w),
                    pointer(el.inverse_jacobian), pointer(el.node_coordinates), ptr_or_null(contravariant),
                    n_if, ptr_or_null(if_ids), ptr_or_null(if_orient), ptr_or_null(if_idx),
                    n_bd, ptr_or_null(bd_ids), ptr_or_null(bd_orient), ptr_or_null(bd_sides), ptr_or_null(bd_x), nbd,
                    n_mo, ptr_or_null(mo_ids), ptr_or_null(mo_sides), ptr_or_null(mo_orient),
                    ptr_or_null(fu), ptr_or_null(fl), ptr_or_null(ru), ptr_or_null(rl),
                    ptr_or_null(left_neighbors),
                    0, 1, 0, C_NULL, C_NULL, C_NULL, C_NULL,   # single rank: no MPI interfaces
                    ptr_or_null(bd_idx), C_NULL,
                    fv_flux, ind_var, ind_smooth, 0, ind_max, ind_min, ptr_or_null(inv_vdm),
                    ptr_or_null(mo_idx),
                    0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL,   # single rank: no MPI mortars
                    subcell_normals)
        rc = ccall((:trixi_b200_create, libtrixi_b200), Cint, (Ref{Desc}, Ref{Ptr{Cvoid}}), desc, handle)
    end
    check(nothing, rc)
    return B200(handle[], nvariables(equations) * nnodes(dg)^ndims(mesh) * nelements(dg, cache))
end

# GlmSpeedCallback (glm_speed.jl:85-105) updates equations.c_h every step
set_c_h!(backend::B200, c_h) = check(backend, ccall((:trixi_b200_set_eq_param, libtrixi_b200), Cint,
                                                    (Ptr{Cvoid}, Cint, Float64), backend.handle, 2, c_h))

# ---- the methods Trixi dispatches to ---------------------------------------------------------------------
# rhs_hyperbolic!(backend, du, u, t, mesh, equations, boundary_conditions, source_terms, dg, cache)
# (dgsem_tree/dg_2d.jl:113-186), reached from rhs_hyperbolic!(du_ode, u_ode, semi, t)
# (semidiscretization_hyperbolic.jl:578-597) once `trixi_backend(u)` returns a B200.
function Trixi.rhs_hyperbolic!(backend::B200, du, u, t, mesh::Union{TreeMesh, StructuredMesh, P4estMesh}, equations,
                               boundary_conditions, source_terms, dg::DG, cache)
    sync_equation_params!(backend, equations)
    backend.device_u_valid = false   # rhs_host leaves the caller's `u` argument on the device, not the integrator's
    GC.@preserve du u begin
        check(backend, ccall((:trixi_b200_rhs_host, libtrixi_b200), Cint,
                             (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64),
                             backend.handle, pointer(du), pointer(u), t))
    end
    return nothing
end

# max_dt(u, t, mesh, constant_speed, equations, dg, cache) (stepsize_dg3d.jl:8-32): on the
# device-resident u of the integrator
function max_dt_device(backend::B200, t)
    dt = Ref{Float64}(NaN)
    check(backend, ccall((:trixi_b200_max_dt, libtrixi_b200), Cint, (Ptr{Cvoid}, Float64, Ref{Float64}),
                         backend.handle, t, dt))
    return dt[]
end

# step!(integrator::SimpleIntegrator2N) (methods_2N.jl:131-168): the stage loop :144-159 runs on the
# device with the stage update fused into the element kernel; u, du, u_tmp stay resident.
function step_2n!(backend::B200, t, dt, alg::Trixi.SimpleAlgorithm2N)
    a, b, c = collect(alg.a), collect(alg.b), collect(alg.c)
    GC.@preserve a b c begin
        check(backend, ccall((:trixi_b200_step_2n, libtrixi_b200), Cint,
                             (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                             backend.handle, t, dt, pointer(a), pointer(b), pointer(c), length(c)))
    end
    return nothing
end

# The same step on the integrator's host-resident u (methods_2N.jl:95-111 keeps u in a Vector): the first stage
# consumes u chunk-wise as it arrives over PCIe, the last stage returns it chunk-wise (trixi_b200_step_2n_host)
function step_2n_host!(backend::B200, u::Vector{Float64}, t, dt, alg::Trixi.SimpleAlgorithm2N)
    a, b, c = collect(alg.a), collect(alg.b), collect(alg.c)
    GC.@preserve u a b c begin
        check(backend, ccall((:trixi_b200_step_2n_host, libtrixi_b200), Cint,
                             (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                             backend.handle, pointer(u), t, dt, pointer(a), pointer(b), pointer(c), length(c)))
    end
    return nothing
end

# step!(integrator::SimpleIntegrator3Sstar) stage loop (methods_3Sstar.jl:186-207)
function step_3sstar!(backend::B200, t, dt, alg::Trixi.SimpleAlgorithm3Sstar)
    g1, g2, g3, be, de, c = collect.((alg.gamma1, alg.gamma2, alg.gamma3, alg.beta, alg.delta, alg.c))
    GC.@preserve g1 g2 g3 be de c begin
        check(backend, ccall((:trixi_b200_step_3sstar, libtrixi_b200), Cint,
                             (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                              Ptr{Float64}, Ptr{Float64}, Cint),
                             backend.handle, t, dt, pointer(g1), pointer(g2), pointer(g3), pointer(be), pointer(de),
                             pointer(c), length(c)))
    end
    return nothing
end

# step!(integrator::SimpleIntegratorSSP) stage loop without stage callbacks (methods_SSP.jl:185-202)
function step_ssp!(backend::B200, t, dt, alg::Trixi.SimpleAlgorithmSSP)
    na, nb, den, c = collect.((alg.numerator_a, alg.numerator_b, alg.denominator, alg.c))
    GC.@preserve na nb den c begin
        check(backend, ccall((:trixi_b200_step_ssp, libtrixi_b200), Cint,
                             (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                             backend.handle, t, dt, pointer(na), pointer(nb), pointer(den), pointer(c), length(c)))
    end
    return nothing
end

# indicator.cache.alpha for the SaveSolutionCallback's :indicator_shock_capturing (indicators.jl:20-24)
function indicator_alpha(backend::B200, nelements)
    alpha = Vector{Float64}(undef, nelements)
    GC.@preserve alpha check(backend, ccall((:trixi_b200_calc_indicator, libtrixi_b200), Cint,
                                            (Ptr{Cvoid}, Ptr{Float64}), backend.handle, pointer(alpha)))
    return alpha
end

upload!(backend::B200, which, host::Vector{Float64}) = GC.@preserve host check(backend,
    ccall((:trixi_b200_upload, libtrixi_b200), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), backend.handle, which,
          pointer(host)))
# Array(u) for the AnalysisCallback / SaveSolutionCallback (analysis_dg3d.jl:172-177)
download!(host::Vector{Float64}, backend::B200, which) = GC.@preserve host check(backend,
    ccall((:trixi_b200_download, libtrixi_b200), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), backend.handle, which,
          pointer(host)))

# ---- the array type that carries the backend ----------------------------------------------------------------
# The reference selects its accelerator path through the array type of the solution vector: `storage_type`
# (src/auxiliary/containers.jl:250-268, ext/TrixiCUDACoreExt.jl:9-11) and `trixi_backend(u) = get_backend(u)`
# (containers.jl:278-285).  `B200Array` is a host-resident `Array` tagged with the `B200` handle: callbacks, I/O and
# `Array(u)` (analysis_dg3d.jl:172-177) see ordinary host memory; the hot-path methods below see the tag.
struct B200Array{T, N} <: DenseArray{T, N}
    data::Array{T, N}
    backend::B200
end
const B200Vector{T} = B200Array{T, 1}
Base.size(a::B200Array) = size(a.data)
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
Base.@propagate_inbounds Base.getindex(a::B200Array, i::Int) = a.data[i]
Base.@propagate_inbounds function Base.setindex!(a::B200Array, v, i::Int)
    a.backend.device_u_valid = false   # a host-side write (callback, limiter, restart) invalidates the device copy
    a.data[i] = v
    return a
end
Base.similar(a::B200Array, ::Type{T}, dims::Dims{N}) where {T, N} = B200Array{T, N}(Array{T, N}(undef, dims), a.backend)
Base.copy(a::B200Array{T, N}) where {T, N} = B200Array{T, N}(copy(a.data), a.backend)
Base.pointer(a::B200Array) = pointer(a.data)
Base.unsafe_convert(::Type{Ptr{T}}, a::B200Array{T}) where {T} = pointer(a.data)
Base.strides(a::B200Array) = strides(a.data)
Base.Array(a::B200Array) = a.data
Base.resize!(::B200Array, n) = error("libtrixi_b200: a change of the element count (AMR) needs a new handle")

Trixi.storage_type(::Type{<:B200Array}) = B200Array
KernelAbstractions.get_backend(a::B200Array) = a.backend
# `@trixi_timeit_ext backend ...` (auxiliary.jl:94-103) synchronises the backend when debug timings are on
KernelAbstractions.synchronize(b::B200) = check(b, ccall((:trixi_b200_synchronize, libtrixi_b200), Cint, (Ptr{Cvoid},),
                                                         b.handle))
# keep the tagged vector off the Polyester/PtrArray path (dg.jl:1189-1204) ...
LoopVectorization.check_args(::B200Array) = false
# ... and carry the tag through `wrap_array` (dg.jl:1205-1211 would `unsafe_wrap(ArrayType{...}, ptr, dims)`,
# which cannot know the handle)
@inline function Trixi.wrap_array(u_ode::B200Array{T, 1}, mesh::Trixi.AbstractMesh, equations, dg::DGSEM,
                                  cache) where {T}
    dims = (nvariables(equations), ntuple(_ -> nnodes(dg), ndims(mesh))..., nelements(dg, cache))
    return B200Array{T, ndims(mesh) + 2}(unsafe_wrap(Array{T, ndims(mesh) + 2}, pointer(u_ode.data), dims),
                                         u_ode.backend)
end
Trixi.wrap_array_native(u_ode::B200Array{T, 1}, mesh::Trixi.AbstractMesh, equations, dg::DG, cache) where {T} =
    Trixi.wrap_array_native(u_ode.data, mesh, equations, dg, cache)

"""
    offload(ode::ODEProblem; device = -1)

Create the `libtrixi_b200` handle for `ode.p` (a `SemidiscretizationHyperbolic`) and return the same problem with
`u0` wrapped in a [`B200Array`](@ref).  `Trixi.solve`/`init` (methods_2N.jl:113-129) then build `u`, `du` and
`u_tmp` with `copy`/`similar`, so the integrator's vectors all carry the backend.
"""
function offload(ode::SciMLBase.ODEProblem; device = -1)
    backend = B200(ode.p; device = device)
    return SciMLBase.remake(ode; u0 = B200Array{Float64, 1}(copy(ode.u0), backend))
end
offload(semi::SemidiscretizationHyperbolic, tspan; kwargs...) = offload(Trixi.semidiscretize(semi, tspan); kwargs...)

# GlmSpeedCallback (glm_speed.jl:85-105) mutates equations.c_h between steps
sync_equation_params!(backend::B200, equations) = nothing
sync_equation_params!(backend::B200, equations::IdealGlmMhdEquations3D) = set_c_h!(backend, equations.c_h)

function ensure_device_u!(u::B200Array)
    if !u.backend.device_u_valid
        GC.@preserve u check(u.backend, ccall((:trixi_b200_upload, libtrixi_b200), Cint,
                                               (Ptr{Cvoid}, Cint, Ptr{Float64}), u.backend.handle, 0, pointer(u.data)))
        u.backend.device_u_valid = true
    end
    return nothing
end

# max_dt(u, t, mesh, constant_speed, equations, dg, cache) (stepsize_dg2d.jl:8-75, stepsize_dg3d.jl:8-123), called by
# calculate_dt (stepsize.jl:146-154) with the wrapped u.  One method per reference method so that the argument
# lists stay comparable (more specific in `u` only): no ambiguities.  After a `step!` below the device copy is
# current and the CFL maxima were already reduced by the last Runge-Kutta stage (TRIXI_B200_OPT_FUSED_CFL).
for (M, C) in ((:(TreeMesh{2}), :False), (:(TreeMesh{2}), :True), (:(TreeMesh{3}), :False), (:(TreeMesh{3}), :True),
               (:(StructuredMesh{2}), :Any), (:(StructuredMesh{3}), :Any), (:(P4estMesh{2}), :Any),
               (:(P4estMesh{3}), :Any))
    @eval function Trixi.max_dt(u::B200Array, t, mesh::$M, constant_speed::$C, equations, dg::DG, cache)
        sync_equation_params!(u.backend, equations)
        ensure_device_u!(u)
        return max_dt_device(u.backend, t)
    end
end

# step!(integrator::SimpleIntegrator2N) (methods_2N.jl:131-168) for integrators whose vectors are B200Arrays: the
# bookkeeping is the reference's, the stage loop :144-159 is one call (u travels in chunks that overlap the first
# and the last stage; afterwards the host and the device copies of u agree).
function Trixi.step!(integrator::Trixi.SimpleIntegrator2N{<:Real, <:B200Array})
    prob = integrator.sol.prob
    alg = integrator.alg
    t_end = last(prob.tspan)
    callbacks = integrator.opts.callback

    @assert !integrator.finalstep
    if isnan(integrator.dt)
        error("time step size `dt` is NaN")
    end
    Trixi.limit_dt!(integrator, t_end)

    backend = integrator.u.backend
    _, equations, _, _ = mesh_equations_solver_cache(prob.p)
    sync_equation_params!(backend, equations)
    check(backend, ccall((:trixi_b200_set_option, libtrixi_b200), Cint, (Ptr{Cvoid}, Cint, Cint), backend.handle, 1, 1))
    step_2n_host!(backend, integrator.u.data, integrator.t, integrator.dt, alg)
    backend.device_u_valid = true

    integrator.iter += 1
    integrator.t += integrator.dt
    Trixi.@trixi_timeit Trixi.timer() "Step-Callbacks" Trixi.handle_callbacks!(callbacks, integrator)
    Trixi.check_max_iter!(integrator)
    return nothing
end

end # module
