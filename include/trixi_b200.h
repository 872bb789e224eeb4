/* libtrixi_b200 -- C ABI of the B200-native DGSEM right-hand side / 2N Runge-Kutta / CFL path.
 *
 * The reference (Trixi.jl v0.17.4-DEV) has no FFI for this path: it is selected by Julia multiple
 * dispatch on the `backend` argument.  Each entry point below is what the Julia
 * method named in its comment forwards to with `ccall` (see INTEGRATION.md / julia/TrixiB200.jl).
 * Plain pointers and sizes only; all arrays use the reference's own layouts: column-major,
 * variable index fastest, `u[v, i, j, (k,) element]` (src/solvers/dg.jl:1169-1212), index arrays are
 * Julia `Int` = int64 and 1-based.
 *
 * Conventions: the caller owns every host array; `create` copies what it needs and
 * the library owns all device memory.  No call aborts: every function returns 0 on success or a
 * negative TRIXI_B200_E* code, and `trixi_b200_last_error` gives the message.  NaNs propagate like in
 * the reference (math.jl:89-97,137-146).  One handle per GPU, calls from one thread at a time.
 * There is NO CPU fallback: without a CUDA device `create` fails with TRIXI_B200_ENODEVICE.
 */
#ifndef TRIXI_B200_H
#define TRIXI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRIXI_B200_ABI_VERSION 5

#if defined(__GNUC__)
#define TRIXI_B200_API __attribute__((visibility("default")))
#else
#define TRIXI_B200_API
#endif

/* status codes */
#define TRIXI_B200_OK 0
#define TRIXI_B200_EINVAL (-1)      /* bad descriptor / unsupported combination */
#define TRIXI_B200_ENODEVICE (-2)   /* no usable CUDA device */
#define TRIXI_B200_ECUDA (-3)       /* CUDA runtime error (message in last_error) */
#define TRIXI_B200_ENOMEM (-4)
#define TRIXI_B200_ECOMM (-5)       /* halo exchange / NCCL error */

/* mesh kinds (dispatch of rhs_hyperbolic!: dgsem_tree/dg_2d.jl:113-120, dgsem_structured/dg.jl:41-94,
 * dgsem_p4est/dg_3d_parallel.jl:8-13) */
enum { TRIXI_B200_MESH_TREE = 0, TRIXI_B200_MESH_STRUCTURED = 1, TRIXI_B200_MESH_P4EST = 2 };

/* equations (src/equations/) */
enum {
    TRIXI_B200_EQ_ADVECTION_2D = 1, /* linear_scalar_advection_2d.jl; params: a1, a2 */
    TRIXI_B200_EQ_EULER_2D = 2,     /* compressible_euler_2d.jl; params: gamma, inv_gamma_minus_one */
    TRIXI_B200_EQ_EULER_3D = 3,     /* compressible_euler_3d.jl:45-54; params: gamma, inv_gamma_minus_one */
    TRIXI_B200_EQ_MHD_3D = 4,       /* ideal_glm_mhd_3d.jl:49-59; params: gamma, inv_gamma_minus_one, c_h */
    TRIXI_B200_EQ_ADVECTION_3D = 5  /* linear_scalar_advection_3d.jl; params: a1, a2, a3 */
};

/* volume integral types (src/solvers/dg.jl:105,135-141) */
enum {
    TRIXI_B200_VOLINT_WEAK_FORM = 0,
    TRIXI_B200_VOLINT_FLUX_DIFFERENCING = 1,
    /* VolumeIntegralShockCapturingHG(indicator; volume_flux_dg, volume_flux_fv) (solvers/dg.jl; driver
     * dgsem/calc_volume_integral.jl:231-272): per element a blend (1 - alpha) flux differencing + alpha first-order
     * subcell finite volumes (fv_kernel! dg_3d.jl:268-306), alpha from IndicatorHennemannGassner.  Compressible
     * Euler and ideal GLM-MHD (calcflux_fv! with nonconservative terms, dg_3d.jl:391-452) on TreeMesh, P4estMesh
     * (across ranks the neighbour's alpha travels with the halo exchange) and StructuredMesh; curved meshes need
     * subcell_normal_vectors below. */
    TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG = 2,
    /* VolumeIntegralPureLGLFiniteVolume(volume_flux_fv) (solvers/dg.jl:559-583; calc_volume_integral.jl: fv_kernel! on
     * every element with alpha = true): first-order finite volumes on the LGL subcells, flux `volume_flux_fv` */
    TRIXI_B200_VOLINT_PURE_LGL_FV = 3
};

/* indicator variables of IndicatorHennemannGassner (compressible_euler_3d.jl:1945-1956 and 2D analogues) */
enum { TRIXI_B200_INDVAR_DENSITY_PRESSURE = 0, TRIXI_B200_INDVAR_DENSITY = 1, TRIXI_B200_INDVAR_PRESSURE = 2 };

/* numerical fluxes (src/equations/numerical_fluxes.jl, compressible_euler_3d.jl) */
enum {
    TRIXI_B200_FLUX_CENTRAL = 0,            /* numerical_fluxes.jl:17-25 */
    TRIXI_B200_FLUX_RANOCHA = 1,            /* compressible_euler_3d.jl:746-828 */
    TRIXI_B200_FLUX_LLF = 2,                /* FluxLaxFriedrichs(max_abs_speed)  :229-253, euler :1156-1199 */
    TRIXI_B200_FLUX_LLF_NAIVE = 3,          /* FluxLaxFriedrichs(max_abs_speed_naive) euler :1112-1153 */
    TRIXI_B200_FLUX_HLL_DAVIS = 4,          /* FluxHLL(min_max_speed_davis) :422-449, euler :1240-1285 */
    TRIXI_B200_FLUX_HLL_NAIVE = 5,          /* FluxHLL(min_max_speed_naive) */
    TRIXI_B200_FLUX_SHIMA_ETAL = 6,         /* compressible_euler_3d.jl:473-547 */
    TRIXI_B200_FLUX_KENNEDY_GRUBER = 7,     /* :560-627 */
    TRIXI_B200_FLUX_CHANDRASHEKAR = 8,      /* :639-733 */
    TRIXI_B200_FLUX_HINDENLANG_GASSNER = 9, /* ideal_glm_mhd_3d.jl:680-855 */
    TRIXI_B200_FLUX_GODUNOV = 10,           /* linear_scalar_advection_2d.jl:248-275 */
    TRIXI_B200_FLUX_RANOCHA_TURBO = 11,     /* dg_3d_compressible_euler.jl:265-617: hoisted logs */
    TRIXI_B200_FLUX_LLF_MHD_POWELL = 12,            /* (flux_lax_friedrichs, flux_nonconservative_powell) */
    TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL = 13, /* (flux_hindenlang_gassner, flux_nonconservative_powell)
                                                       ideal_glm_mhd_3d.jl:295-340,680-779 */
    TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL = 14,      /* (FluxLaxFriedrichs(max_abs_speed_naive), powell) */
    TRIXI_B200_FLUX_HLLE_MHD_POWELL = 15,           /* (flux_hlle, powell): FluxHLL(min_max_speed_einfeldt)
                                                       numerical_fluxes.jl:422-457, ideal_glm_mhd_3d.jl:1094-1130,
                                                       Roe averages :1415-1491 */
    TRIXI_B200_FLUX_CENTRAL_MHD_POWELL = 16,        /* (flux_central, flux_nonconservative_powell) */
    TRIXI_B200_FLUX_HLLE = 17,              /* flux_hlle = FluxHLL(min_max_speed_einfeldt) numerical_fluxes.jl:457, compressible
                                             * Euler: compressible_euler_3d.jl:1662-1771, compressible_euler_2d.jl:1925-2025 */
    TRIXI_B200_FLUX_HLLC = 18               /* flux_hllc compressible_euler_3d.jl:1423-1665, compressible_euler_2d.jl:1720-1925 */
};

/* source terms (calc_sources! dg_3d.jl:1417-1437 calls an arbitrary closure; here: registry) */
enum {
    TRIXI_B200_SRC_NONE = 0,
    TRIXI_B200_SRC_CONVERGENCE_TEST = 1,              /* compressible_euler_3d.jl:127-153 / _2d */
    TRIXI_B200_SRC_EOC_TEST_EULER = 2,                /* compressible_euler_3d.jl:265-284 */
    TRIXI_B200_SRC_EOC_TEST_COUPLED_EULER_GRAVITY = 3 /* :228-249 */
};

/* initial conditions usable as Dirichlet boundary value functions */
enum {
    TRIXI_B200_IC_NONE = 0,
    TRIXI_B200_IC_CONSTANT = 1,         /* compressible_euler_3d.jl:78-86 */
    TRIXI_B200_IC_CONVERGENCE_TEST = 2, /* :94-111 */
    TRIXI_B200_IC_WEAK_BLAST_WAVE = 3,  /* :163-184 */
    TRIXI_B200_IC_EOC_TEST_COUPLED_EULER_GRAVITY = 4
};

/* boundary conditions, per direction -x,+x,-y,+y,-z,+z (semidiscretization_hyperbolic.jl:158-164) */
enum {
    TRIXI_B200_BC_PERIODIC = 0, /* basic_types.jl:58-127: no boundary faces in that direction */
    TRIXI_B200_BC_DIRICHLET = 1, /* BoundaryConditionDirichlet equations.jl:159-183 */
    TRIXI_B200_BC_SLIP_WALL = 2  /* compressible_euler_3d.jl:315-417 */
};

/* Descriptor: everything `create_cache` (dgsem_tree/dg_2d.jl:14-37) and the DG/equation structs hold
 * that the hot path reads.  All pointers are host pointers, read during `create` only. */
typedef struct trixi_b200_desc {
    int32_t abi_version; /* TRIXI_B200_ABI_VERSION */
    int32_t device;      /* CUDA device ordinal, -1 = current */
    int32_t ndims, nvars, nnodes, mesh_kind;
    int64_t nelements;

    int32_t equation;
    int32_t volume_integral, volume_flux, surface_flux;
    int32_t source_terms;
    int32_t boundary_conditions[6]; /* TRIXI_B200_BC_* per direction */
    int32_t boundary_ic[6];         /* TRIXI_B200_IC_* for Dirichlet directions */
    int32_t reserved0;
    double eq_params[8];

    /* LobattoLegendreBasis (basis_lobatto_legendre.jl:17-31), column-major [n, n] */
    const double *derivative_split;
    const double *derivative_hat;
    const double *inverse_weights; /* [n]; surface lifting uses inverse_weights[1] (dg_3d.jl:1349) */

    /* element container (containers_3d.jl:9-18 / dgsem_structured/containers.jl:8-34) */
    const double *inverse_jacobian;      /* Tree: [nelements]; curved: [n^d, nelements] */
    const double *node_coordinates;      /* [ndims, n^d, nelements] */
    const double *contravariant_vectors; /* curved only: [ndims, ndims, n^d, nelements], else NULL */

    /* interface container (containers_3d.jl:136-144), conforming faces */
    int64_t ninterfaces;
    const int64_t *interface_neighbor_ids;  /* [2, ninterfaces] 1-based, 1 = left/-, 2 = right/+ */
    const int64_t *interface_orientations;  /* [ninterfaces] in 1..ndims */
    const int64_t *interface_node_indices;  /* P4est only: [ndims, 2, ninterfaces] node_indices tuples
                                               (dgsem_p4est/containers.jl:226-252) encoded :begin 0, :end 1,
                                               :i_forward 2, :i_backward 3, :j_forward 4, :j_backward 5 */

    /* boundary container (containers_3d.jl:284-295), sorted by direction */
    int64_t nboundaries;
    const int64_t *boundary_neighbor_ids;     /* [nboundaries] */
    const int64_t *boundary_orientations;     /* [nboundaries] */
    const int64_t *boundary_neighbor_sides;   /* [nboundaries] 1: element on the - side */
    const double *boundary_node_coordinates;  /* [ndims, n^(d-1), nboundaries] */
    int64_t n_boundaries_per_direction[6];

    /* L2 mortar container (containers_3d.jl:495-510); nmortars = 0 on conforming meshes */
    int64_t nmortars;
    const int64_t *mortar_neighbor_ids; /* [2^(d-1)+1, nmortars] */
    const int64_t *mortar_large_sides;  /* [nmortars] */
    const int64_t *mortar_orientations; /* [nmortars] */
    const double *mortar_forward_upper, *mortar_forward_lower; /* [n, n] */
    const double *mortar_reverse_upper, *mortar_reverse_lower; /* [n, n] */

    /* StructuredMesh: left_neighbors [ndims, nelements], 0 = domain boundary (containers.jl:8-34) */
    const int64_t *left_neighbors;

    /* distributed run (replaces P4estMPICache dg_parallel.jl:8-20): faces shared with other ranks */
    int32_t rank, world_size;
    int64_t nmpiinterfaces;
    const int64_t *mpi_local_neighbor_ids; /* [nmpiinterfaces] local element, 1-based */
    const int64_t *mpi_local_sides;        /* [nmpiinterfaces] 1: local element is left/-, 2: right/+ */
    const int64_t *mpi_orientations;       /* [nmpiinterfaces] */
    const int64_t *mpi_neighbor_ranks;     /* [nmpiinterfaces] peer rank of each face; faces are sorted by
                                              (peer rank, global interface id) on both sides */

    /* P4est boundary container (dgsem_p4est/containers.jl:302-345): node_indices [ndims, nboundaries], same
     * encoding as interface_node_indices; boundaries sorted by boundary name = direction */
    const int64_t *boundary_node_indices;

    /* P4est MPI interface container (dgsem_p4est/containers_parallel.jl:8-28): node_indices of the local side
     * [ndims, nmpiinterfaces], same encoding; the exchanged face states are aligned at the primary element */
    const int64_t *mpi_node_indices;

    /* VolumeIntegralShockCapturingHG: volume_flux above is volume_flux_dg; IndicatorHennemannGassner
     * (dgsem/indicators.jl:48-70,114-148; dgsem_tree/indicators_3d.jl:41-131, indicators_2d.jl:26-98) */
    int32_t volume_flux_fv;          /* TRIXI_B200_FLUX_* of the subcell finite volume fluxes */
    int32_t indicator_variable;      /* TRIXI_B200_INDVAR_* */
    int32_t indicator_alpha_smooth;  /* apply_smoothing! over interfaces and mortars (indicators_3d.jl:133-186) */
    int32_t reserved1;
    double indicator_alpha_max, indicator_alpha_min;
    const double *inverse_vandermonde_legendre; /* [n, n] column-major (basis_lobatto_legendre.jl:711-724) */

    /* P4est L2 mortar container (dgsem_p4est/containers.jl:563-613): node_indices [ndims, 2, nmortars], 1: the small
     * side, 2: the large side, same encoding as interface_node_indices; mortar_neighbor_ids holds the small elements
     * by position and the large element last; mortar_large_sides / mortar_orientations are TreeMesh-only */
    const int64_t *mortar_node_indices;

    /* MPI mortars (MPIL2MortarContainer dgsem_tree/containers_2d.jl:1003-1070, P4estMPIMortarContainer
     * dgsem_p4est/containers_parallel.jl:130-210): mortars whose elements live on more than one rank.
     * mpi_mortar_neighbor_ids [2^(d-1)+1, nmpimortars], small elements by position, the large element last:
     *   > 0  element of this rank (1-based);
     *   < 0  minus the 1-based position of the MPI-interface entry through which that element's face state arrives
     *        (its remote side); those entries have mpi_is_mortar_piece = 1: both ranks send the face of their own
     *        element there as for a conforming shared face, but no interface flux is evaluated;
     *   = 0  a small element this rank neither owns nor needs (it does not own the large element).
     * Every rank evaluates the mortar fluxes of the positions it has both sides of and stores them for its own
     * elements only (calc_mpi_mortar_flux! + mpi_mortar_fluxes_to_elements!, dg_2d_parallel.jl:742-860,
     * dgsem_p4est/dg_3d_parallel.jl:382-560). */
    int64_t nmpimortars;
    const int64_t *mpi_mortar_neighbor_ids;
    const int64_t *mpi_mortar_large_sides;        /* TreeMesh: [nmpimortars] */
    const int64_t *mpi_mortar_orientations;       /* TreeMesh: [nmpimortars] */
    const int64_t *mpi_mortar_node_indices;       /* P4est: [ndims, 2, nmpimortars], 1: small side, 2: large side */
    const double *mpi_mortar_normal_directions;   /* P4est: [ndims, n^(d-1), 2^(d-1), nmpimortars] outward normals of
                                                     the small elements (they may be remote) */
    const int64_t *mpi_is_mortar_piece;           /* [nmpiinterfaces], NULL = all zero */

    /* VolumeIntegralShockCapturingHG on curved meshes: NormalVectorContainer (dgsem_structured/containers_3d.jl:488-541),
     * the free-stream preserving normals of the subcell interfaces used by calcflux_fv! (dgsem_structured/
     * dg_3d.jl:377-436): direction a -> [ndims, n .. (n - 1 along a) .., nelements]; NULL otherwise */
    const double *subcell_normal_vectors[3];
} trixi_b200_desc;

typedef struct trixi_b200_handle trixi_b200_handle;

/* ---- life cycle --------------------------------------------------------------------------------
 * Called once after `create_cache(mesh, equations, dg, RealT, uEltype)` (dgsem_tree/dg_2d.jl:14-37);
 * mirrors what `trixi_adapt`/`semidiscretize(...; storage_type)` does for the reference GPU path
 * (semidiscretization.jl:115-126): all containers are uploaded once. */
TRIXI_B200_API int trixi_b200_create(const trixi_b200_desc *desc, trixi_b200_handle **out);
TRIXI_B200_API void trixi_b200_destroy(trixi_b200_handle *h);
TRIXI_B200_API const char *trixi_b200_last_error(const trixi_b200_handle *h); /* h may be NULL: last create error */
TRIXI_B200_API int trixi_b200_abi_version(void);

/* Device-resident solution vectors owned by the handle: u, du, u_tmp (methods_2N.jl:95-111).
 * `which`: 0 = u, 1 = du, 2 = u_tmp.  Host arrays have nvars*n^d*nelements doubles. */
TRIXI_B200_API int trixi_b200_upload(trixi_b200_handle *h, int which, const double *host);
TRIXI_B200_API int trixi_b200_download(trixi_b200_handle *h, int which, double *host); /* Array(u) analysis_dg3d.jl:172-177 */
TRIXI_B200_API void *trixi_b200_device_ptr(trixi_b200_handle *h, int which);           /* raw device pointer (for DLPack/torch views) */
TRIXI_B200_API int trixi_b200_synchronize(trixi_b200_handle *h);
TRIXI_B200_API void *trixi_b200_stream(trixi_b200_handle *h); /* the library-owned cudaStream_t */

/* ---- hot path -----------------------------------------------------------------------------------
 * rhs_hyperbolic!(backend, du, u, t, mesh, equations, boundary_conditions, source_terms, dg, cache)
 * (dgsem_tree/dg_2d.jl:113-186; structured dgsem_structured/dg.jl:41-94): du <- rhs(u, t), host
 * buffers: copies u host->device, runs the kernels, copies du device->host, synchronises. */
TRIXI_B200_API int trixi_b200_rhs_host(trixi_b200_handle *h, double *du_host, const double *u_host, double t);
/* On conforming TreeMeshes without physical boundaries (single rank) the host-buffer calls are pipelined: u is
 * uploaded in element chunks along a sweep of the last coordinate axis, every interface / element kernel starts as
 * soon as its neighbours are resident, and finished chunks are downloaded while later ones are still in flight,
 * so both PCIe directions and the kernels overlap (pinned host buffers required for the overlap, not for
 * correctness).  Results are bit-identical to the unpipelined path.  See TRIXI_B200_OPT_HOST_PIPELINE_CHUNK. */
/* same on the device-resident vectors (asynchronous on the handle's stream) */
TRIXI_B200_API int trixi_b200_rhs(trixi_b200_handle *h, double t);

/* max_dt(u, t, mesh, constant_speed, equations, dg, cache) (stepsize_dg3d.jl:8-32, stepsize_dg2d.jl):
 * returns 2 / (nnodes * max_e invJ_e * sum_d max_nodes lambda_d) over the device-resident u;
 * the caller multiplies by cfl(t) (stepsize.jl:146-154).  With world_size > 1 this is the rank-local
 * value; the caller takes the minimum over ranks exactly where the reference calls
 * MPI.Allreduce!(dt, min) (stepsize_dg3d.jl:264-279).  Synchronises. */
TRIXI_B200_API int trixi_b200_max_dt(trixi_b200_handle *h, double t, double *dt_out);

/* step!(integrator::SimpleIntegrator2N) stage loop (methods_2N.jl:144-159): for every stage
 * du <- rhs(u, t + c_s dt); u_tmp <- du - a_s u_tmp; u <- u + (b_s dt) u_tmp, with the stage update
 * fused into the last RHS kernel.  u_tmp is zeroed first.  Asynchronous. */
TRIXI_B200_API int trixi_b200_step_2n(trixi_b200_handle *h, double t, double dt, const double *a, const double *b,
                       const double *c, int nstages);
/* The same step on a host-resident u (read and overwritten in place): the integrator's u lives in a Julia Vector
 * (methods_2N.jl:95-111).  The first stage consumes u chunk-wise as it arrives, the last stage returns it
 * chunk-wise as it is finished.  Synchronises. */
TRIXI_B200_API int trixi_b200_step_2n_host(trixi_b200_handle *h, double *u_host, double t, double dt, const double *a,
                                            const double *b, const double *c, int nstages);
/* step!(integrator::SimpleIntegrator3Sstar) stage loop (methods_3Sstar.jl:186-207; tableaus
 * ParsaniKetchesonDeconinck3Sstar94/32 :63-133): u_tmp1 <- 0, u_tmp2 <- u, then per stage du <- rhs(u, t + c_s dt),
 * u_tmp1 += delta_s u, u <- gamma1_s u + gamma2_s u_tmp1 + gamma3_s u_tmp2 + beta_s dt du.  Asynchronous. */
TRIXI_B200_API int trixi_b200_step_3sstar(trixi_b200_handle *h, double t, double dt, const double *gamma1, const double *gamma2,
                                           const double *gamma3, const double *beta, const double *delta, const double *c,
                                           int nstages);
/* step!(integrator::SimpleIntegratorSSP) stage loop without stage callbacks (methods_SSP.jl:185-202, SimpleSSPRK33
 * :23-52): u_tmp <- u, then per stage du <- rhs(u, t + c_s dt), u <- u + dt du,
 * u <- (numerator_a_s u_tmp + numerator_b_s u) / denominator_s.  Asynchronous. */
TRIXI_B200_API int trixi_b200_step_ssp(trixi_b200_handle *h, double t, double dt, const double *numerator_a,
                                        const double *numerator_b, const double *denominator, const double *c, int nstages);
/* `nsteps` steps with the CFL step size recomputed on the device after every step
 * (StepsizeCallback interval = 1, stepsize.jl:93-126) and the final step clipped to t_end
 * (time_integration.jl:46-55); no host round trip inside.  Returns the number of steps taken, the
 * final time and the last dt through the out-pointers.  Synchronises at the end. */
TRIXI_B200_API int trixi_b200_solve_2n(trixi_b200_handle *h, double t0, double t_end, double cfl, int64_t max_steps,
                        const double *a, const double *b, const double *c, int nstages,
                        int64_t *steps_out, double *t_out, double *dt_out);

/* calc_error_norms (callbacks_step/analysis_dg3d.jl:123-216, analysis_dg2d.jl:133-215) on the device-resident u:
 * u, x (and the Jacobian of curved meshes) are interpolated to the n_analysis^d analysis nodes with the
 * [n_analysis, nnodes] column-major Vandermonde matrix (SolutionAnalyzer, basis_lobatto_legendre.jl:274-290);
 * the exact solution is the registered initial condition at time t.  Returns, per variable, the quadrature sum
 * of the squared errors (not yet divided by the volume, no square root: distributed callers add these over ranks
 * like analysis_dg3d.jl:218-275) and the maximum absolute error, plus the quadrature of the volume. */
TRIXI_B200_API int trixi_b200_calc_error_norms(trixi_b200_handle *h, double t, int initial_condition, int n_analysis,
                                               const double *vandermonde, const double *weights, double *l2_sums,
                                               double *linf, double *volume);

/* integrate_via_indices (callbacks_step/analysis_dg3d.jl:364-473, analysis_dg2d.jl analogues; the reference has a
 * device form of it, :405-458) for the integrands of the AnalysisCallback's `analysis_integrals` (analysis.jl:680-760)
 * on the device-resident u: the quadrature sum over this rank's elements, NOT normalised (distributed callers add the
 * sums and the volumes over ranks, then divide), plus the quadrature of the volume.  `integral` has nvars entries for
 * TRIXI_B200_INTEGRAL_CONS, one entry otherwise.  TRIXI_B200_INTEGRAL_ENTROPY_TIMEDERIVATIVE is
 * analyze(entropy_timederivative, du, u, ...) (analysis_dg3d.jl:506-517): sum of cons2entropy(u) . du with the
 * device-resident du (e.g. of the last trixi_b200_rhs).  Entropy and energies: compressible Euler. */
enum {
    TRIXI_B200_INTEGRAL_CONS = 0,                   /* integrate(u): conservation (cons2cons) */
    TRIXI_B200_INTEGRAL_ENTROPY = 1,                /* entropy = entropy_math (compressible_euler_3d.jl:1970-2009) */
    TRIXI_B200_INTEGRAL_ENERGY_TOTAL = 2,           /* :2012 */
    TRIXI_B200_INTEGRAL_ENERGY_KINETIC = 3,         /* :2015-2018 */
    TRIXI_B200_INTEGRAL_ENERGY_INTERNAL = 4,        /* :2021-2023 */
    TRIXI_B200_INTEGRAL_ENTROPY_TIMEDERIVATIVE = 5  /* cons2entropy :1796-1817 */
};
TRIXI_B200_API int trixi_b200_integrate(trixi_b200_handle *h, int quantity, double *integral, double *volume);

/* Tuning knobs (the analogue of the reference's compile-time Preferences, src/Trixi.jl:18-23).
 * TRIXI_B200_OPT_KERNEL_PATH: 0 = tuned kernels where one exists (default), 1 = generic kernels only,
 *   2 = the previous generation of the tuned headline kernel (kept for A/B measurements).
 * TRIXI_B200_OPT_FUSED_CFL: 1 = the last stage of trixi_b200_step_2n also reduces the CFL wave speeds of the
 *   state it writes, so the trixi_b200_max_dt that follows (StepsizeCallback, stepsize.jl:75-117) costs no pass
 *   over u.  The cached value is dropped by trixi_b200_upload(0), trixi_b200_rhs_host and
 *   trixi_b200_device_ptr(0); a caller that keeps a raw pointer to u and writes through it must leave this at
 *   0 (default). */
#define TRIXI_B200_OPT_KERNEL_PATH 0
#define TRIXI_B200_OPT_FUSED_CFL 1
#define TRIXI_B200_OPT_PREFETCH_DISTANCE 2 /* tuned element kernel: L2 prefetch distance in elements (0 = off) */
#define TRIXI_B200_OPT_HOST_PIPELINE_CHUNK 3 /* host-buffer calls: elements per chunk; -1 = auto (16-32 MiB, only
                                                for >= 32 chunks), 0 = one copy each way, no overlap */
/* TRIXI_B200_OPT_RK_REDUCE_UPDATE (default 1): the tuned 2N stage kernels apply u += (b dt) u_tmp as a bulk
 * reduce-add onto u in L2 instead of reading u back into the SM; the product is rounded before the addition in
 * both forms, so the option does not change results. */
#define TRIXI_B200_OPT_RK_REDUCE_UPDATE 4
/* TRIXI_B200_OPT_SINGLE_FACE_FLUX (default 1): on TreeMeshes with conservative equations both neighbours of a
 * conforming interface receive the same flux (dg_3d.jl:581-597); where the tuned element kernel supports it the
 * interface kernel writes it once (the left element's + face) and the right element fetches it from there.  Does not
 * change results; trixi_b200_download_surface_flux_values always returns the reference's two-copy layout. */
#define TRIXI_B200_OPT_SINGLE_FACE_FLUX 5
/* TRIXI_B200_OPT_L2_HINTS: L2 eviction priorities on the bulk copies of the tuned headline kernel (u evict_last until
 * its reduce-add, everything else evict_first).  Performance only. */
#define TRIXI_B200_OPT_L2_HINTS 6
/* TRIXI_B200_OPT_FUSED_STAGE (default 1): trixi_b200_step_3sstar / trixi_b200_step_ssp apply their stage updates
 * (methods_3Sstar.jl:195-205, methods_SSP.jl:192-201) in the epilogue of the element kernel, as trixi_b200_step_2n
 * always does; 0 = rhs! into du followed by a pointwise stage kernel.  Same operations, bit-identical results. */
#define TRIXI_B200_OPT_FUSED_STAGE 7
TRIXI_B200_API int trixi_b200_set_option(trixi_b200_handle *h, int option, int value);

/* GlmSpeedCallback (glm_speed.jl:85-105) mutates equations.c_h every step */
TRIXI_B200_API int trixi_b200_set_eq_param(trixi_b200_handle *h, int index, double value);

/* ---- stage-level entry points (parity tests against the reference's stage functions) ------------
 * calc_volume_integral! (calc_volume_integral.jl:180-191): du <- volume terms only (set_zero! included) */
TRIXI_B200_API int trixi_b200_calc_volume_integral(trixi_b200_handle *h);
/* prolong2interfaces! + calc_interface_flux! + prolong2boundaries! + calc_boundary_flux!
 * (dg_3d.jl:530-602,651-768) -> surface_flux_values[nvars, n^(d-1), 2*ndims, nelements] */
TRIXI_B200_API int trixi_b200_calc_surface_fluxes(trixi_b200_handle *h, double t);
TRIXI_B200_API int trixi_b200_download_surface_flux_values(trixi_b200_handle *h, double *host);
/* the blending factors alpha [nelements] of IndicatorHennemannGassner for the device-resident u (what the
 * reference exposes as element variable :indicator_shock_capturing, indicators.jl:20-24); runs the indicator */
TRIXI_B200_API int trixi_b200_calc_indicator(trixi_b200_handle *h, double *alpha_host);

/* ---- distributed halo exchange (replaces the MPI Isend/Irecv of dg_parallel.jl:66-182) ----------------
 * One handle per rank/GPU.  Every handle with world_size > 1 owns a receive buffer that its neighbour
 * ranks map (CUDA IPC across processes, plain pointers inside one process); during an RHS evaluation the
 * pack kernel of rank A stores A's face states directly into B's buffer over NVLink and raises a
 * sequence flag, B's MPI-interface kernel waits on it.  The host process group (torch.distributed / MPI)
 * only moves the opaque connection blobs once: all-gather `comm_info` of every rank, pass the
 * concatenation (rank order) to `comm_connect`.  All ranks must then evaluate the RHS collectively. */
TRIXI_B200_API int64_t trixi_b200_comm_info_size(void);
TRIXI_B200_API int trixi_b200_comm_info(trixi_b200_handle *h, void *blob_out);
TRIXI_B200_API int trixi_b200_comm_connect(trixi_b200_handle *h, const void *blobs, int world);

/* ---- measurement helpers --------------------------------------------------------------------- */
/* number of kernel launches issued by this handle since creation */
TRIXI_B200_API int64_t trixi_b200_launch_count(const trixi_b200_handle *h);
/* elapsed device milliseconds of the most recent trixi_b200_rhs/step_2n call, measured with CUDA
 * events on the handle's stream (feeds the PerformanceCounter, semidiscretization_hyperbolic.jl:586-594) */
TRIXI_B200_API int trixi_b200_last_elapsed_ms(trixi_b200_handle *h, float *ms_out);
/* CUDA-event stopwatch on the handle's stream (bench.py brackets its timed region with it) */
TRIXI_B200_API int trixi_b200_timer_start(trixi_b200_handle *h);
TRIXI_B200_API int trixi_b200_timer_stop(trixi_b200_handle *h, float *ms_out); /* synchronises */
/* Dependency-free DFMA / streaming-copy microbenchmarks on the handle's device: the FP64 roofline
 * denominator is not in MEASURED_PEAKS.json and must be measured on the box. */
TRIXI_B200_API int trixi_b200_measure_fp64_peak(trixi_b200_handle *h, double *tflops_out);
TRIXI_B200_API int trixi_b200_measure_copy_bandwidth(trixi_b200_handle *h, double *gbs_out);
/* per-kernel-class accumulated device time (ms) and launch counts since the last reset; classes:
 * 0 = surface-flux kernel, 1 = element kernel (volume+surface+jacobian+source+RK), 2 = max_dt,
 * 3 = halo pack + signal + MPI interface flux, 4 = halo wait (spinning on the neighbours' flags: skew between
 * ranks shows up here).  Enabling costs two event records per launch. */
TRIXI_B200_API int trixi_b200_profile_enable(trixi_b200_handle *h, int on);
TRIXI_B200_API int trixi_b200_profile_read(trixi_b200_handle *h, int kernel_class, double *ms_out, int64_t *launches_out);

#ifdef __cplusplus
}
#endif
#endif /* TRIXI_B200_H */
