# TrixiB200.jl -- the reference-side binding of libtrixi_b200.so.
#
# Trixi.jl selects its solver kernels by multiple dispatch on a `backend` argument
# (src/solvers/dgsem_tree/dg_2d.jl:113-120 `backend::Nothing`, src/solvers/dgsem_p4est/dg_2d_gpu.jl:8
# `backend::Backend`).  This file adds a backend type `B200` and the methods whose bodies `ccall` into
# the C ABI of include/trixi_b200.h -- the same mechanism as ext/TrixiCUDACoreExt.jl uses to plug in
# CUDA.jl, but without KernelAbstractions.  Julia is not installed in the build container or on the GPU
# box, so this file is exercised only by inspection; its Python twin (trixi.jl_b200/lib.py) binds the
# identical entry points and is what the test-suite drives.
#
# Usage (the only change to an elixir):
#
#     using Trixi, TrixiB200
#     semi = SemidiscretizationHyperbolic(mesh, equations, initial_condition, solver; ...)
#     semi = TrixiB200.offload(semi)            # uploads the cache once (create_cache -> trixi_b200_create)
#     ode  = semidiscretize(semi, tspan)
#     sol  = Trixi.solve(ode, Trixi.CarpenterKennedy2N54(); dt = 1.0, callback = callbacks)
module TrixiB200

using Trixi
using Trixi: TreeMesh, DG, DGSEM, SemidiscretizationHyperbolic, nvariables, nnodes, ndims,
             nelements, ninterfaces, nboundaries, mesh_equations_solver_cache

const libtrixi_b200 = get(ENV, "TRIXI_B200_LIBRARY", "libtrixi_b200.so")

# ---- enums of include/trixi_b200.h ------------------------------------------------------------------
const MESH_TREE = Cint(0)
equation_id(::LinearScalarAdvectionEquation2D) = Cint(1)
equation_id(::CompressibleEulerEquations2D) = Cint(2)
equation_id(::CompressibleEulerEquations3D) = Cint(3)
equation_params(eq::LinearScalarAdvectionEquation2D) = (eq.advection_velocity..., 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
equation_params(eq::Union{CompressibleEulerEquations2D, CompressibleEulerEquations3D}) = (eq.gamma,
                                                                                          eq.inv_gamma_minus_one,
                                                                                          0.0, 0.0, 0.0, 0.0, 0.0,
                                                                                          0.0)
volume_integral_id(::VolumeIntegralWeakForm) = (Cint(0), Cint(0))
volume_integral_id(v::VolumeIntegralFluxDifferencing) = (Cint(1), flux_id(v.volume_flux))
flux_id(::typeof(flux_central)) = Cint(0)
flux_id(::typeof(flux_ranocha)) = Cint(1)
flux_id(f::FluxLaxFriedrichs) = f.dissipation.max_abs_speed === max_abs_speed_naive ? Cint(3) : Cint(2)
flux_id(f::FluxHLL) = f.min_max_speed === min_max_speed_naive ? Cint(5) : Cint(4)
flux_id(::typeof(flux_shima_etal)) = Cint(6)
flux_id(::typeof(flux_kennedy_gruber)) = Cint(7)
flux_id(::typeof(flux_chandrashekar)) = Cint(8)
flux_id(::typeof(flux_godunov)) = Cint(10)
flux_id(f) = error("numerical flux $f is not in the libtrixi_b200 registry")
source_id(::Nothing) = Cint(0)
source_id(::typeof(source_terms_convergence_test)) = Cint(1)
source_id(::typeof(Trixi.source_terms_eoc_test_euler)) = Cint(2)
source_id(f) = error("source term $f is not in the libtrixi_b200 registry")
ic_id(::typeof(initial_condition_constant)) = Cint(1)
ic_id(::typeof(initial_condition_convergence_test)) = Cint(2)
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
end

# ---- the backend object ---------------------------------------------------------------------------------
mutable struct B200
    handle::Ptr{Cvoid}
    ulength::Int
    function B200(handle, ulength)
        b = new(handle, ulength)
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

# Called once after `create_cache` (dgsem_tree/dg_2d.jl:14-37); the analogue of
# `semidiscretize(...; storage_type = CuArray)` adapting all containers (semidiscretization.jl:115-126).
function B200(semi::SemidiscretizationHyperbolic; device = -1)
    mesh, equations, dg, cache = mesh_equations_solver_cache(semi)
    mesh isa TreeMesh || error("this build of libtrixi_b200 accelerates TreeMesh")
    @assert Trixi.nmortars(dg, cache) == 0
    basis = dg.basis
    volint, volflux = volume_integral_id(dg.volume_integral)
    bcs = semi.boundary_conditions isa NamedTuple ? values(semi.boundary_conditions) :
          ntuple(_ -> semi.boundary_conditions, 2 * ndims(mesh))
    bc_tags = ntuple(i -> i <= length(bcs) ? bc_id(bcs[i])[1] : Cint(0), 6)
    bc_ics = ntuple(i -> i <= length(bcs) ? bc_id(bcs[i])[2] : Cint(0), 6)
    el, ifc, bd = cache.elements, cache.interfaces, cache.boundaries
    D_split, D_hat = Matrix(basis.derivative_split), Matrix(basis.derivative_hat)
    inv_w = collect(basis.inverse_weights)
    nbd = ntuple(i -> i <= 2 * ndims(mesh) ? Int64(bd.n_boundaries_per_direction[i]) : Int64(0), 6)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve D_split D_hat inv_w el ifc bd begin   # the library copies during `create` only
        desc = Desc(2, device, ndims(mesh), nvariables(equations), nnodes(dg), MESH_TREE,
                    nelements(dg, cache), equation_id(equations), volint, volflux,
                    flux_id(dg.surface_integral.surface_flux), source_id(semi.source_terms),
                    bc_tags, bc_ics, 0, equation_params(equations),
                    pointer(D_split), pointer(D_hat), pointer(inv_w),
                    pointer(el.inverse_jacobian), pointer(el.node_coordinates), C_NULL,
                    ninterfaces(dg, cache), pointer(ifc.neighbor_ids), pointer(ifc.orientations), C_NULL,
                    nboundaries(dg, cache), pointer(bd.neighbor_ids), pointer(bd.orientations),
                    pointer(bd.neighbor_sides), pointer(bd.node_coordinates), nbd,
                    0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL,
                    0, 1, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)
        rc = ccall((:trixi_b200_create, libtrixi_b200), Cint, (Ref{Desc}, Ref{Ptr{Cvoid}}), desc, handle)
    end
    check(nothing, rc)
    return B200(handle[], nvariables(equations) * nnodes(dg)^ndims(mesh) * nelements(dg, cache))
end

# ---- the methods Trixi dispatches to ---------------------------------------------------------------------
# rhs_hyperbolic!(backend, du, u, t, mesh, equations, boundary_conditions, source_terms, dg, cache)
# (dgsem_tree/dg_2d.jl:113-186), reached from rhs_hyperbolic!(du_ode, u_ode, semi, t)
# (semidiscretization_hyperbolic.jl:578-597) once `trixi_backend(u)` returns a B200.
function Trixi.rhs_hyperbolic!(backend::B200, du, u, t, mesh::TreeMesh, equations, boundary_conditions,
                               source_terms, dg::DG, cache)
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

upload!(backend::B200, which, host::Vector{Float64}) = GC.@preserve host check(backend,
    ccall((:trixi_b200_upload, libtrixi_b200), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), backend.handle, which,
          pointer(host)))
# Array(u) for the AnalysisCallback / SaveSolutionCallback (analysis_dg3d.jl:172-177)
download!(host::Vector{Float64}, backend::B200, which) = GC.@preserve host check(backend,
    ccall((:trixi_b200_download, libtrixi_b200), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), backend.handle, which,
          pointer(host)))

end # module
