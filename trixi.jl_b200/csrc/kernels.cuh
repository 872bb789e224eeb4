// Generic DGSEM kernels for Cartesian (TreeMesh) elements, templated on <equation, nnodes>.
//
// Launch structure of one RHS evaluation:
//   1. k_interface_flux   prolong2interfaces! + calc_interface_flux! fused (dg_3d.jl:530-602): gathers
//                         both face states straight from u, writes surface_flux_values of both elements.
//   2. k_boundary_flux    prolong2boundaries! + calc_boundary_flux! fused (dg_3d.jl:651-768).
//   3. k_element          set_zero! + calc_volume_integral! + calc_surface_integral! + apply_jacobian!
//                         + calc_sources! fused (dg_3d.jl:133-214,1337-1437), and optionally the 2N
//                         Runge-Kutta stage update (methods_2N.jl:150-158) and the per-element CFL
//                         maxima of max_dt (stepsize_dg3d.jl:8-32) in the same pass.
// Because the surface fluxes only depend on u, they are computed first; the element kernel then
// produces the finished du (or the updated u) in a single sweep: u and du are each touched once.
#pragma once
#include "physics.cuh"

namespace tb {

constexpr int kMaxNodes = 8;

struct KParams {
    // sizes
    long long nelements, ninterfaces, nboundaries;
    // solution vectors
    const double *u;  // [nv, n^d, nelem]
    double *du;
    double *u_tmp;
    double *u_out;  // RK mode: updated u (may alias u)
    double *sfv;    // surface_flux_values [nv, n^(d-1), 2d, nelem]
    // operators (column-major [n, n]) live in constant-like global memory, staged to smem per block
    const double *dsplit, *dhat;
    double dsplit_c[kMaxNodes * kMaxNodes];  // same matrix in the kernel-parameter constant bank: a DFMA
                                             // can take it as an operand without a load or a register
    double dsplit_h[16], dsplit_q[16], dsplit_e[16];  // nnodes = 4 only: D_split / 2, / 4, / 8 (exact), for the tuned
                                             // kernels whose two-point fluxes come scaled by powers of two
    double inv_weight0;
    // geometry
    const double *inverse_jacobian;       // Tree: [nelem]; curved: [n^d, nelem]
    const double *node_coordinates;       // [nd, n^d, nelem]
    const double *contravariant_vectors;  // curved: [nd (dim), nd (index), n^d, nelem]
    int curved;                           // 0: Cartesian TreeMesh kernels, 1: curved (Structured/P4est) kernels
    int p4est;                            // 1: P4est conventions (outward normals, "+" surface integral, node_indices)
    const long long *if_node_indices;     // P4est: [nd, 2, I] encoded symbols
    const long long *bd_node_indices;     // P4est: [nd, B]
    // connectivity (1-based int64 as uploaded)
    const long long *if_neighbors;  // [2, I]
    const long long *if_orient;     // [I]
    const long long *bd_neighbor, *bd_orient, *bd_side;
    const double *bd_coords;  // [nd, n^(d-1), B]
    const int *bd_direction;  // [B] 1-based direction of each boundary (from n_boundaries_per_direction)
    int bc[6], bc_ic[6];
    // physics
    EqParams eq;
    int volume_integral, volume_flux, surface_flux, source_terms;
    double t;
    // RK stage fused into the element kernels.  mode 0: write du.
    // mode 1, 2N (methods_2N.jl:152-158): u_tmp = du - u_tmp * a; u = u + u_tmp * b_dt
    // mode 2, 3S* (methods_3Sstar.jl:195-205): u_tmp (= u_tmp1) += delta * u;
    //         u = gamma1 * u + gamma2 * u_tmp + gamma3 * u_tmp2 + (beta dt) * du;  rk_k = {delta, gamma1, gamma2, gamma3},
    //         rk_b_dt = beta dt
    // mode 3, SSP (methods_SSP.jl:192-201): u = (numerator_a * u_tmp + numerator_b * (u + dt du)) / denominator;
    //         rk_k = {numerator_a, numerator_b, denominator}, rk_b_dt = dt
    // rk_read_tmp = 0: the kernel takes the known value of u_tmp instead of reading it (2N: first stage, a = 0; 3S*:
    // first stage, u_tmp1 = 0; SSP: first stage, u_tmp = u, which the kernel then also writes: no copy pass).
    // rk_write_tmp = 0: u_tmp is left alone (SSP after the first stage).
    int mode;
    int rk_read_tmp, rk_write_tmp;
    double rk_a, rk_b_dt, rk_k[4];
    const double *u_tmp2;  // 3S*: the third register (read only)
    // CFL fused output: per-block max of invJ * sum_d max_nodes lambda_d, encoded as ordered uint64
    unsigned long long *cfl_key;  // kCflSlots partial maxima (spread same-address atomics over L2 slices)
    int want_cfl;                 // RK stage kernels that support it also reduce the CFL speed of the updated u
    int kernel_path;              // 0: tuned kernels where available, 1: generic kernels only
    int prefetch_distance;        // tuned element kernels: L2 prefetch this many elements ahead (0: off, < 0: one
                                  // wave of resident CTAs, resolved at launch)
    int sm_count;
    int rk_reduce_update;         // tuned kernels: u += b dt u_tmp as a bulk reduce-add in L2 (TRIXI_B200_OPT_RK_REDUCE_UPDATE)
    // Single-copy interface fluxes (TreeMesh, conservative equations: both neighbours of a conforming interface get
    // the same flux, dg_3d.jl:581-597): the interface kernel only writes the left element's + face, and the element
    // kernel fetches its - faces from its left neighbours' + faces.  minus_nb[ndims, nelements]: 0-based element
    // across the - face in each direction, or -1 where that face is a boundary, a mortar or shared with another
    // rank (those kernels keep writing the element's own slot).  sfv_single is decided per RHS evaluation.
    const int *minus_nb;
    int sfv_single;
    int l2_hints;                 // tuned headline kernel: L2 eviction priorities on its bulk copies (TRIXI_B200_OPT_L2_HINTS)
    long long elem_begin, elem_end;  // TreeMesh element kernels work on [elem_begin, elem_end) (pipelined rhs_host)
    // VolumeIntegralShockCapturingHG: blending factors of IndicatorHennemannGassner
    double *alpha;       // [nelem] after smoothing (ordered-bits atomicMax target)
    double *alpha_raw;   // [nelem] before smoothing
    const double *inv_vdm;  // inverse_vandermonde_legendre [n, n] column-major
    const double *subcell_normals[3];  // curved meshes: normal vectors of the subcell interfaces per direction
    double inv_weights_c[kMaxNodes];
    double weights_c[kMaxNodes];  // quadrature weights (1 / inverse_weights), for integrate_via_indices
    double ind_alpha_max, ind_alpha_min;
    int volume_flux_fv, ind_var, ind_smooth;
    // distributed: faces shared with other ranks (replaces mpi_interfaces, dg_2d_parallel.jl / dg_parallel.jl)
    long long nmpi;
    const long long *mpi_local, *mpi_side, *mpi_orient;  // [nmpi] 1-based local element, local side, orientation
    const long long *mpi_node_indices;                   // P4est: [ndims, nmpi] node_indices of the local side
    // L2 mortars (TreeMesh): neighbor_ids [2^(d-1)+1, M] 1-based (small elements by position, then the large
    // one), large_sides, orientations; forward (interpolation) and reverse (projection) operators [n, n]
    long long nmortars;
    const long long *mortar_ids, *mortar_large_sides, *mortar_orient;
    const long long *mortar_node_indices;         // P4est: [nd, 2, M], 1 = small side, 2 = large side
    // MPI mortars: ids > 0 local element, < 0 minus the 1-based MPI-interface entry whose received face is that
    // element's, 0 not available (include/trixi_b200.h); the exchange-only entries are flagged in mpi_is_piece
    long long nmpimortars;
    const long long *mpi_mortar_ids, *mpi_mortar_large_sides, *mpi_mortar_orient, *mpi_mortar_node_indices;
    const double *mpi_mortar_normals;             // P4est: [nd, NF, 2^(d-1), MM]
    const long long *mpi_is_piece;                // [nmpi] or nullptr
    const double *mortar_fwd[2], *mortar_rev[2];  // [0] lower, [1] upper
    const int *mpi_peer_slot;                            // [nmpi] index into the peer tables
    const long long *mpi_remote_index;                   // [nmpi] slot of this face in the peer's receive buffer
    double *const *peer_recv;                            // [npeers] peer receive buffers (current parity), NVLink-mapped
    const double *recv;                                  // my receive buffer (current parity) [nv, nf, nmpi]
    // shock capturing across ranks: the neighbour element's unsmoothed blending factor travels with its face
    // ([nmpi] doubles behind the face states of every receive buffer)
    const double *recv_alpha;                            // my receive buffer's alpha part (current parity) [nmpi]
    const long long *mpi_peer_nmpi;                      // [npeers] faces in that peer's receive buffer
};

__host__ __device__ constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

// volume node (0-based linear) of face node fn on layer s normal to orientation o
// (face node order x:(j,k) y:(i,k) z:(i,j), dg_3d.jl:540-563)
template <int ND, int N>
TB_DEV int face_to_volume_node(int o, int s, int fn) {
    if constexpr (ND == 2) {
        return o == 0 ? s + N * fn : fn + N * s;
    } else {
        const int a = fn % N, b = fn / N;
        return o == 0 ? s + N * (a + N * b) : (o == 1 ? a + N * (s + N * b) : a + N * (b + N * s));
    }
}

constexpr int kCflSlots = 1024;

TB_DEV unsigned long long cfl_encode(double v) {
    // positive doubles order like their bit patterns; NaN must win the max (Base.max propagates NaN,
    // stepsize_dg3d.jl:24-28)
    if (isnan(v)) return 0x7ff8000000000000ull;
    return (unsigned long long)__double_as_longlong(v);
}

// Fused stage updates of the 3S* and SSP integrators for one value (P.mode 2 and 3; see KParams): d = du, x = u_tmp
// (not looked at when !P.rk_read_tmp), u, u2 = u_tmp2 (3S* only).  Returns the new u; x_new is what u_tmp receives when
// P.rk_write_tmp.  The operations are those of the unfused k_stage_3sstar / k_stage_ssp, so both paths give the same bits.
TB_DEV double rk_stage_3s_ssp(const KParams &P, double d, double x, double u, double u2, double &x_new) {
    if (P.mode == 2) {
        // u_tmp1 += delta * u; u = gamma1 u + gamma2 u_tmp1 + gamma3 u_tmp2 + beta dt du (methods_3Sstar.jl:195-205)
        const double t1 = P.rk_read_tmp ? fma(P.rk_k[0], u, x) : P.rk_k[0] * u;
        x_new = t1;
        return fma(P.rk_b_dt, d, fma(P.rk_k[3], u2, fma(P.rk_k[2], t1, P.rk_k[1] * u)));
    }
    // u = (numerator_a u_tmp + numerator_b (u + dt du)) / denominator (methods_SSP.jl:192-201); u_tmp = u in stage 1
    const double xx = P.rk_read_tmp ? x : u;
    x_new = xx;
    const double ue = fma(P.rk_b_dt, d, u);
    return fma(P.rk_k[0], xx, P.rk_k[1] * ue) / P.rk_k[2];
}

// ---- 1. interfaces ------------------------------------------------------------------------------
template <class EQ>
struct HasFastRanocha {
    static constexpr bool value = false;
};
template <int ND>
struct HasFastRanocha<Euler<ND>> {
    static constexpr bool value = true;
};
template <int ND>
struct HasFastRanocha<EulerAllFluxes<ND>> {
    static constexpr bool value = true;
};

// surface flux of one face node; FAST selects the fast-division flux_ranocha (tuned path)
// FAST: 0 = the registry's generic numflux, 1 = flux_ranocha on fast divisions, 2 = FluxLaxFriedrichs on fast divisions
// (compile-time: a run-time choice between the two fast fluxes cost the flux_ranocha interface kernel 22%)
template <class EQ, int FAST>
TB_DEV void surface_numflux(const EQ &eq, int id, const double (&ul)[EQ::NVARS], const double (&ur)[EQ::NVARS], int o,
                            double (&f)[EQ::NVARS]) {
    if constexpr (FAST == 1 && HasFastRanocha<EQ>::value) {
        eq.flux_ranocha_fast(ul, ur, o, f);
    } else if constexpr (FAST == 2 && HasFastRanocha<EQ>::value) {
        eq.flux_llf_fast(id, ul, ur, o, f);
    } else {
        eq.numflux(id, ul, ur, o, f);
    }
}

// Equation systems with nonconservative terms (GLM-MHD): one compiled body of the surface flux for every interface
// kernel.  N ranks must reproduce one rank bit for bit (the MPI interface kernel and the local interface kernels
// evaluate the same face), which needs the same FMA contraction of these long expressions at every call site;
// inlined copies in different kernels are not guaranteed to get it.  fl / fr: flux + 0.5 * noncons for the left /
// right element (calc_interface_flux! with nonconservative terms, dg_3d.jl:604-649).
template <class EQ>
__device__ __noinline__ void surface_flux_noncons(const EQ eq, int id, const double *ul, const double *ur, int o,
                                                  double *fl, double *fr) {
    constexpr int NV = EQ::NVARS;
    double a[NV], b[NV], f[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        a[v] = ul[v];
        b[v] = ur[v];
    }
    eq.numflux(id, a, b, o, f);
    if constexpr (EQ::kHasNoncons) {
        if (EQ::has_noncons(id)) {
            double gl[NV], gr[NV];
            eq.noncons(a, b, o, gl);
            eq.noncons(b, a, o, gr);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                fl[v] = f[v] + 0.5 * gl[v];
                fr[v] = f[v] + 0.5 * gr[v];
            }
            return;
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) fl[v] = fr[v] = f[v];
}

template <class EQ, int N, int FAST = 0>
__global__ void __launch_bounds__(256) k_interface_flux(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / NF;
    const int fn = (int)(gid % NF);
    if (I >= P.ninterfaces) return;
    const EQ eq(P.eq);
    const long long left = P.if_neighbors[2 * I] - 1, right = P.if_neighbors[2 * I + 1] - 1;
    const int o = (int)P.if_orient[I] - 1;
    const int nl = face_to_volume_node<ND, N>(o, N - 1, fn), nr = face_to_volume_node<ND, N>(o, 0, fn);
    double ul[NV], ur[NV], f[NV];
    const double *pl = P.u + (left * NN + nl) * NV, *pr = P.u + (right * NN + nr) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        ul[v] = pl[v];
        ur[v] = pr[v];
    }
    // left element: direction 2*orientation (1-based) = index 2o+1; right element: 2o (dg_3d.jl:581-597)
    double *sl = P.sfv + ((left * (2 * ND) + (2 * o + 1)) * NF + fn) * NV;
    double *sr = P.sfv + ((right * (2 * ND) + (2 * o)) * NF + fn) * NV;
    if constexpr (EQ::kHasNoncons) {
        double fl[NV], fr[NV];
        surface_flux_noncons<EQ>(eq, P.surface_flux, ul, ur, o, fl, fr);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            sl[v] = fl[v];
            sr[v] = fr[v];
        }
        return;
    }
    surface_numflux<EQ, FAST>(eq, P.surface_flux, ul, ur, o, f);
#pragma unroll
    for (int v = 0; v < NV; ++v) sl[v] = f[v];
    if (!P.sfv_single) {
#pragma unroll
        for (int v = 0; v < NV; ++v) sr[v] = f[v];
    }
}

// Same stage with the face data staged through shared memory: a warp owns 32 face nodes (32 / NF whole
// interfaces), gathers both sides' face values with lane-consecutive addresses (u is AoS, so a face is made of
// contiguous runs of NV, N*NV or N*N*NV doubles), computes one flux per lane and writes the two
// surface_flux_values faces (contiguous NF*NV doubles each) back fully coalesced. Used when NF divides 32.
// CURVED (StructuredMesh, dgsem_structured/dg_3d.jl:619-753): the flux is taken along the contravariant vector of
// the right element's first node layer times sign(inverse_jacobian), and stored with that sign on both sides.
template <class EQ, int N, int FAST = 0, bool CURVED = false>
__global__ void __launch_bounds__(256, (CURVED || EQ::kHasNoncons) ? 3 : (FAST == 1 ? 6 : (FAST == 2 ? 5 : 4)))
    k_interface_flux_staged(const KParams P) {  // (resident blocks: what ptxas chose unprompted, 5 for the LLF form)
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    constexpr int G = 32 / NF;   // interfaces per warp
    constexpr int FV = NF * NV;  // doubles per face
    constexpr int WPB = 8;       // warps per block
    static_assert(32 % NF == 0, "staged interface kernel needs NF | 32");
    __shared__ double s_all[WPB][2 * G * FV];
    __shared__ long long s_elem[WPB][2 * G];
    __shared__ int s_orient[WPB][G];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *s = s_all[warp];
    const long long I0 = ((long long)blockIdx.x * WPB + warp) * G;
    if (I0 >= P.ninterfaces) return;  // the whole warp leaves together
    const int nvalid = (int)min((long long)G, P.ninterfaces - I0);
    if (lane < 2 * nvalid) s_elem[warp][lane] = P.if_neighbors[2 * I0 + lane] - 1;
    if (lane < nvalid) s_orient[warp][lane] = (int)P.if_orient[I0 + lane] - 1;
    __syncwarp();
    // 1. gather: slot c = 2 * g + side (side 0: left element, its +face; side 1: right element, its -face).
    // A slot is warp-uniform, so the element, the orientation and the node strides are uniform values and a
    // lane only adds its own (a, b, v) offset: face node fn = a + N b sits at volume node
    // idx S + a SA + b SB with (S, SA, SB) = (1, N, N^2), (N, 1, N^2), (N^2, 1, N) for orientation 0, 1, 2.
    constexpr int PASSES = (FV + 31) / 32;
    int la[PASSES], lb[PASSES], lv[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
        const int q = lane + 32 * p;
        const int fnq = q / NV;
        lv[p] = q - fnq * NV;
        lb[p] = fnq / N;
        la[p] = fnq - lb[p] * N;
    }
#pragma unroll
    for (int c = 0; c < 2 * G; ++c) {
        if (c < 2 * nvalid) {
            const int o = s_orient[warp][c >> 1];
            const int S = o == 0 ? 1 : (o == 1 ? N : N * N);
            const int SA = o == 0 ? N : 1;
            const int SB = (ND == 3 && o == 2) ? N : N * N;
            const double *base = P.u + (s_elem[warp][c] * NN + ((c & 1) ? 0 : (N - 1) * S)) * NV;
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                const int q = lane + 32 * p;
                if (q < FV) s[c * FV + q] = base[(la[p] * SA + lb[p] * SB) * NV + lv[p]];
            }
        }
    }
    __syncwarp();
    // 2. one face node per lane
    const int g = lane / NF, fn = lane - g * NF;
    double fl[NV], fr[NV];
    if (g < nvalid) {
        const EQ eq(P.eq);
        const int o = s_orient[warp][g];
        double ul[NV], ur[NV], f[NV];
        constexpr bool kFromMemory = FAST == 2 && !CURVED && HasFastRanocha<EQ>::value;
        if constexpr (!kFromMemory) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                ul[v] = s[(2 * g) * FV + fn * NV + v];
                ur[v] = s[(2 * g + 1) * FV + fn * NV + v];
            }
        }
        bool done = false;
        if constexpr (CURVED) {
            const int SA = o == 0 ? N : 1;
            const int SB = (ND == 3 && o == 2) ? N : N * N;
            const int b = fn / N, a = fn - b * N;
            const long long nr = s_elem[warp][2 * g + 1] * NN + (a * SA + b * SB);
            const double ij = P.inverse_jacobian[nr];
            const double sign_jacobian = ij > 0 ? 1.0 : (ij < 0 ? -1.0 : 0.0);
            const double *ja = P.contravariant_vectors + (nr * ND + o) * ND;
            double nrm[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) nrm[d] = ja[d] * sign_jacobian;
            eq.numflux_normal(P.surface_flux, ul, ur, nrm, f);
#pragma unroll
            for (int v = 0; v < NV; ++v) fl[v] = fr[v] = sign_jacobian * f[v];
            done = true;
        } else if constexpr (kFromMemory) {
            eq.flux_llf_fast_mem(P.surface_flux, s + (2 * g) * FV + fn * NV, s + (2 * g + 1) * FV + fn * NV, o, f);
        } else if constexpr (EQ::kHasNoncons) {
            surface_flux_noncons<EQ>(eq, P.surface_flux, ul, ur, o, fl, fr);
            done = true;
        } else {
            surface_numflux<EQ, FAST>(eq, P.surface_flux, ul, ur, o, f);
        }
        if (!done) {
#pragma unroll
            for (int v = 0; v < NV; ++v) fl[v] = fr[v] = f[v];
        }
    }
    __syncwarp();
    if (g < nvalid) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            s[(2 * g) * FV + fn * NV + v] = fl[v];
            s[(2 * g + 1) * FV + fn * NV + v] = fr[v];
        }
    }
    __syncwarp();
    // 3. scatter: left element's direction 2o+1, right element's 2o (0-based; dg_3d.jl:581-597)
#pragma unroll
    for (int c = 0; c < 2 * G; ++c) {
        if (c < 2 * nvalid && !((c & 1) && !CURVED && P.sfv_single)) {  // (single copy: the left element's + face only)
            const int o = s_orient[warp][c >> 1];
            const int dir = (c & 1) ? 2 * o : 2 * o + 1;
            double *dst = P.sfv + ((s_elem[warp][c] * (2 * ND) + dir) * NF) * NV;
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                const int q = lane + 32 * p;
                if (q < FV) dst[q] = s[c * FV + q];
            }
        }
    }
}

// surface_flux_values in the reference's layout after a single-copy evaluation: the right element's - face gets
// the left element's + face (only the stage-level download needs it)
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_sfv_fill_right(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), FV = NF * NV;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / FV;
    const int q = (int)(gid - I * FV);
    if (I >= P.ninterfaces) return;
    const long long left = P.if_neighbors[2 * I] - 1, right = P.if_neighbors[2 * I + 1] - 1;
    const int o = (int)P.if_orient[I] - 1;
    P.sfv[(right * (2 * ND) + 2 * o) * FV + q] = P.sfv[(left * (2 * ND) + 2 * o + 1) * FV + q];
}

// ---- 2. boundaries ------------------------------------------------------------------------------
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_boundary_flux(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long B = gid / NF;
    const int fn = (int)(gid % NF);
    if (B >= P.nboundaries) return;
    const EQ eq(P.eq);
    const long long element = P.bd_neighbor[B] - 1;
    const int o = (int)P.bd_orient[B] - 1;
    const int side = (int)P.bd_side[B];
    const int direction = P.bd_direction[B];  // 1-based
    const int vn = face_to_volume_node<ND, N>(o, side == 1 ? N - 1 : 0, fn);
    double ui[NV], f[NV], x[ND];
    const double *pu = P.u + (element * NN + vn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) ui[v] = pu[v];
#pragma unroll
    for (int d = 0; d < ND; ++d) x[d] = P.bd_coords[(B * NF + fn) * ND + d];
    boundary_flux(eq, P.bc[direction - 1], P.bc_ic[direction - 1], P.surface_flux, ui, o, direction, x, P.t, f);
    double *s = P.sfv + ((element * (2 * ND) + (direction - 1)) * NF + fn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = f[v];
}

// The point flux of a TreeMesh mortar, shared (out of line) by the local and the MPI mortar kernel so that N ranks
// evaluate exactly the arithmetic of one rank (inlined copies may contract to different FMAs).  fprim goes to the
// small element, fsec is projected to the large one; they differ by the nonconservative terms only
// (dg_3d.jl:1012-1233: 0.5 nonconservative_flux(u_large, u_small) and 0.5 nonconservative_flux(u_small, u_large)).
template <class EQ>
__device__ __noinline__ void mortar_point_flux(const EQ &eq, int flux_id, const double *ularge, const double *usmall, int o,
                                               bool large_left, double *fprim, double *fsec) {
    constexpr int NV = EQ::NVARS;
    double up[NV], us[NV], f[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        up[v] = ularge[v];
        us[v] = usmall[v];
    }
    if (large_left)
        eq.numflux(flux_id, up, us, o, f);
    else
        eq.numflux(flux_id, us, up, o, f);
#pragma unroll
    for (int v = 0; v < NV; ++v) fprim[v] = fsec[v] = f[v];
    if constexpr (EQ::kHasNoncons) {
        if (EQ::has_noncons(flux_id)) {
            double np_[NV], ns_[NV];
            eq.noncons(up, us, o, np_);
            eq.noncons(us, up, o, ns_);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                fprim[v] = f[v] + 0.5 * np_[v];
                fsec[v] = f[v] + 0.5 * ns_[v];
            }
        }
    }
}

// ---- 2a. L2 mortars ---------------------------------------------------------------------------------
// prolong2mortars! + calc_mortar_flux! + mortar_fluxes_to_elements! fused (dg_2d.jl:899-1243,
// dg_3d.jl:770-1335), conservative equations.  One block per mortar, one thread per (sub-face, face node);
// the large face, the interpolation temporaries and the 2^(d-1) flux faces live in shared memory.  Position
// p: bit 0 = upper half along the first face coordinate, bit 1 = upper half along the second.
template <class EQ, int N>
__global__ void __launch_bounds__((1 << (EQ::NDIMS - 1)) * ipow(N, EQ::NDIMS - 1)) k_mortar_flux(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND), NP = 1 << (ND - 1);
    __shared__ double s_large[NV * NF];
    __shared__ double s_tmp[NP][NV * NF];
    __shared__ double s_f[NP][NV * NF];
    const long long m = blockIdx.x;
    const int tid = threadIdx.x;
    const int p = tid / NF, fn = tid - p * NF;
    const int a = fn % N, b = fn / N;
    const EQ eq(P.eq);
    const long long *ids = P.mortar_ids + (NP + 1) * m;
    const long long large = ids[NP] - 1, small = ids[p] - 1;
    const int o = (int)P.mortar_orient[m] - 1;
    const bool large_left = P.mortar_large_sides[m] == 1;  // large element on the negative side
    const double *fwd1 = P.mortar_fwd[p & 1], *fwd2 = P.mortar_fwd[(p >> 1) & 1];
    const double *rev1 = P.mortar_rev[p & 1];
    // face of the large element that touches the mortar
    for (int q = tid; q < NV * NF; q += NP * NF) {
        const int f = q / NV, v = q - f * NV;
        const int vn = face_to_volume_node<ND, N>(o, large_left ? N - 1 : 0, f);
        s_large[q] = P.u[(large * NN + vn) * NV + v];
    }
    __syncthreads();
    // element_solutions_to_mortars!: interpolate to sub-face p, first face coordinate first
    double up[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double acc = 0.0;
        for (int q = 0; q < N; ++q) acc += fwd1[a + N * q] * s_large[v + NV * (q + N * b)];
        up[v] = acc;
    }
    if constexpr (ND == 3) {
#pragma unroll
        for (int v = 0; v < NV; ++v) s_tmp[p][v + NV * fn] = up[v];
        __syncthreads();
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double acc = 0.0;
            for (int q = 0; q < N; ++q) acc += fwd2[b + N * q] * s_tmp[p][v + NV * (a + N * q)];
            up[v] = acc;
        }
        __syncthreads();  // s_tmp is reused by the projection below
    }
    // calc_fstar!: the left state is the side on the negative side of the mortar
    double us[NV], f[NV];
    {
        const int vn = face_to_volume_node<ND, N>(o, large_left ? 0 : N - 1, fn);
        const double *pu = P.u + (small * NN + vn) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) us[v] = pu[v];
    }
    // the small element takes its (primary) flux as it is (its face towards the large element)
    {
        double fsec[NV];
        mortar_point_flux<EQ>(eq, P.surface_flux, up, us, o, large_left, f, fsec);
        double *dst = P.sfv + ((small * (2 * ND) + (large_left ? 2 * o : 2 * o + 1)) * NF + fn) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            dst[v] = f[v];
            s_f[p][v + NV * fn] = fsec[v];
        }
    }
    __syncthreads();
    // L2 projection onto the large face
    double *out = P.sfv + ((large * (2 * ND) + (large_left ? 2 * o + 1 : 2 * o)) * NF) * NV;
    if constexpr (ND == 2) {
        // multiply_dimensionwise!(out, reverse_upper, f_upper, reverse_lower, f_lower) (dg_2d.jl:1238-1240)
        if (tid < N) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double acc = 0.0;
                for (int q = 0; q < N; ++q)
                    acc += P.mortar_rev[1][tid + N * q] * s_f[1][v + NV * q] + P.mortar_rev[0][tid + N * q] * s_f[0][v + NV * q];
                out[v + NV * tid] = acc;
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double acc = 0.0;
            for (int q = 0; q < N; ++q) acc += rev1[a + N * q] * s_f[p][v + NV * (q + N * b)];
            s_tmp[p][v + NV * fn] = acc;
        }
        __syncthreads();
        if (tid < NF) {
            // upper_left, upper_right, lower_left, lower_right in this order (dg_3d.jl:1314-1331)
            const int order[4] = {2, 3, 0, 1};
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double res = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pp = order[k];
                    const double *r2 = P.mortar_rev[(pp >> 1) & 1];
                    double acc = 0.0;
                    for (int q = 0; q < N; ++q) acc += r2[b + N * q] * s_tmp[pp][v + NV * (a + N * q)];
                    res = k == 0 ? acc : res + acc;
                }
                out[v + NV * fn] = res;
            }
        }
    }
}

// ---- 2b. faces shared with other ranks -------------------------------------------------------------
// prolong2mpiinterfaces! + start_mpi_send! fused (dg_2d_parallel.jl:565-598, dg_parallel.jl:66-132): the
// local face state is stored straight into the neighbour rank's receive buffer over NVLink/PCIe peer
// mapping -- no send staging, no separate copy.
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_mpi_pack(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / NF;
    const int fn = (int)(gid % NF);
    if (I >= P.nmpi) return;
    const long long element = P.mpi_local[I] - 1;
    const int o = (int)P.mpi_orient[I] - 1;
    const int side = (int)P.mpi_side[I];
    const int vn = face_to_volume_node<ND, N>(o, side == 1 ? N - 1 : 0, fn);
    const double *pu = P.u + (element * NN + vn) * NV;
    const int slot = P.mpi_peer_slot[I];
    double *dst = P.peer_recv[slot] + (P.mpi_remote_index[I] * NF + fn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) dst[v] = pu[v];
    if (fn == 0 && P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG)
        P.peer_recv[slot][P.mpi_peer_nmpi[slot] * NF * NV + P.mpi_remote_index[I]] = P.alpha_raw[element];
    // (no fence here: the signal kernel that follows in stream order issues one system-scope fence per peer before
    // it raises the flag, and fences are cumulative over everything that happened before the kernel started)
}

// The same with the faces staged through shared memory (NF | 32): a warp gathers 32 / NF local faces with
// lane-consecutive addresses and then writes each face record (NV * NF contiguous doubles in the neighbour's receive
// buffer) as 16-byte stores from consecutive lanes.  The one-thread-per-face-node form above issues 8-byte stores
// 40 bytes apart; local HBM merges those in L2, but over NVLink every one of them travels as its own small packet
// (measured: 100 GB/s effective for the exchange on links that carry 770 GB/s).
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_mpi_pack_staged(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    constexpr int G = 32 / NF, FV = NF * NV, WPB = 8;
    static_assert(32 % NF == 0 && FV % 2 == 0, "staged pack kernel needs NF | 32 and an even face record");
    __shared__ __align__(16) double s_all[WPB][G * FV];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *s = s_all[warp];
    const long long I0 = ((long long)blockIdx.x * WPB + warp) * G;
    if (I0 >= P.nmpi) return;
    const int nvalid = (int)min((long long)G, P.nmpi - I0);
    constexpr int PASSES = (FV + 31) / 32;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g < nvalid) {
            const long long I = I0 + g;
            const int o = (int)P.mpi_orient[I] - 1, side = (int)P.mpi_side[I];
            const int S = o == 0 ? 1 : (o == 1 ? N : N * N);
            const int SA = o == 0 ? N : 1;
            const int SB = (ND == 3 && o == 2) ? N : N * N;
            const double *base = P.u + ((P.mpi_local[I] - 1) * NN + (side == 1 ? (N - 1) * S : 0)) * NV;
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                const int q = lane + 32 * p;
                if (q < FV) {
                    const int fnq = q / NV, v = q - fnq * NV, b = fnq / N, a = fnq - b * N;
                    s[g * FV + q] = base[(a * SA + b * SB) * NV + v];
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g < nvalid) {
            const long long I = I0 + g;
            const int slot = P.mpi_peer_slot[I];
            double2 *dst = reinterpret_cast<double2 *>(P.peer_recv[slot] + P.mpi_remote_index[I] * FV);
            const double2 *src = reinterpret_cast<const double2 *>(s + g * FV);
            for (int q = lane; q < FV / 2; q += 32) dst[q] = src[q];
            if (lane == 0 && P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG)
                P.peer_recv[slot][P.mpi_peer_nmpi[slot] * FV + P.mpi_remote_index[I]] = P.alpha_raw[P.mpi_local[I] - 1];
        }
    }
}

// calc_mpi_interface_flux! (dg_2d_parallel.jl:700-740, dg_3d_parallel.jl:167-242): the shared flux is
// computed on both ranks with identical operands; only the local element's storage is written.
template <class EQ, int N, int FAST = 0>
__global__ void __launch_bounds__(256) k_mpi_interface_flux(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / NF;
    const int fn = (int)(gid % NF);
    if (I >= P.nmpi) return;
    if (P.mpi_is_piece && P.mpi_is_piece[I]) return;  // exchange-only entry of an MPI mortar
    const EQ eq(P.eq);
    const long long element = P.mpi_local[I] - 1;
    const int o = (int)P.mpi_orient[I] - 1;
    const int side = (int)P.mpi_side[I];
    const int vn = face_to_volume_node<ND, N>(o, side == 1 ? N - 1 : 0, fn);
    const double *pl = P.u + (element * NN + vn) * NV;
    const double *pr = P.recv + (I * NF + fn) * NV;
    double ul[NV], ur[NV], f[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const double a = pl[v], b = pr[v];
        ul[v] = side == 1 ? a : b;
        ur[v] = side == 1 ? b : a;
    }
    const int direction0 = side == 1 ? 2 * o + 1 : 2 * o;
    double *s = P.sfv + ((element * (2 * ND) + direction0) * NF + fn) * NV;
    if constexpr (EQ::kHasNoncons) {
        double fl[NV], fr[NV];
        surface_flux_noncons<EQ>(eq, P.surface_flux, ul, ur, o, fl, fr);
#pragma unroll
        for (int v = 0; v < NV; ++v) s[v] = side == 1 ? fl[v] : fr[v];
        return;
    }
    surface_numflux<EQ, FAST>(eq, P.surface_flux, ul, ur, o, f);
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = f[v];
}

// The same stage in the form of k_interface_flux_staged: a warp owns 32 / NF shared faces, gathers the local side
// from u with lane-consecutive addresses, copies the neighbour rank's side from the receive buffer (already one
// contiguous [NV, NF] record per face), computes one flux per lane with exactly the arithmetic of the local
// interface kernel (N ranks must reproduce one rank bit for bit) and writes the local element's face coalesced.
template <class EQ, int N, int FAST = 0>
__global__ void __launch_bounds__(256, EQ::kHasNoncons ? 3 : (FAST == 1 ? 6 : (FAST == 2 ? 5 : 4)))
    k_mpi_interface_flux_staged(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    constexpr int G = 32 / NF, FV = NF * NV, WPB = 8;
    static_assert(32 % NF == 0, "staged interface kernel needs NF | 32");
    __shared__ double s_all[WPB][2 * G * FV];
    __shared__ long long s_elem[WPB][G];
    __shared__ int s_orient[WPB][G], s_side[WPB][G];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *s = s_all[warp];
    const long long I0 = ((long long)blockIdx.x * WPB + warp) * G;
    if (I0 >= P.nmpi) return;  // the whole warp leaves together
    const int nvalid = (int)min((long long)G, P.nmpi - I0);
    if (lane < nvalid) {
        s_elem[warp][lane] = P.mpi_local[I0 + lane] - 1;
        s_orient[warp][lane] = (int)P.mpi_orient[I0 + lane] - 1;
        s_side[warp][lane] = (int)P.mpi_side[I0 + lane];
    }
    __syncwarp();
    constexpr int PASSES = (FV + 31) / 32;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g < nvalid) {
            const int o = s_orient[warp][g], side = s_side[warp][g];
            const int S = o == 0 ? 1 : (o == 1 ? N : N * N);
            const int SA = o == 0 ? N : 1;
            const int SB = (ND == 3 && o == 2) ? N : N * N;
            // local element on the left/- side (side 1): its +face; on the right/+ side: its -face
            const double *base = P.u + (s_elem[warp][g] * NN + (side == 1 ? (N - 1) * S : 0)) * NV;
            const double *remote = P.recv + (I0 + g) * FV;
            double *sl = s + (2 * g + (side == 1 ? 0 : 1)) * FV, *sr = s + (2 * g + (side == 1 ? 1 : 0)) * FV;
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                const int q = lane + 32 * p;
                if (q < FV) {
                    const int fnq = q / NV, v = q - fnq * NV, b = fnq / N, a = fnq - b * N;
                    sl[q] = base[(a * SA + b * SB) * NV + v];
                    sr[q] = remote[q];
                }
            }
        }
    }
    __syncwarp();
    const int g = lane / NF, fn = lane - g * NF;
    double fl[NV], fr[NV];
    if (g < nvalid) {
        const EQ eq(P.eq);
        const int o = s_orient[warp][g];
        double f[NV];
        constexpr bool kFromMemory = FAST == 2 && HasFastRanocha<EQ>::value;
        if constexpr (kFromMemory) {
            eq.flux_llf_fast_mem(P.surface_flux, s + (2 * g) * FV + fn * NV, s + (2 * g + 1) * FV + fn * NV, o, f);
#pragma unroll
            for (int v = 0; v < NV; ++v) fl[v] = fr[v] = f[v];
        } else {
            double ul[NV], ur[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                ul[v] = s[(2 * g) * FV + fn * NV + v];
                ur[v] = s[(2 * g + 1) * FV + fn * NV + v];
            }
            if constexpr (EQ::kHasNoncons) {
                surface_flux_noncons<EQ>(eq, P.surface_flux, ul, ur, o, fl, fr);
            } else {
                surface_numflux<EQ, FAST>(eq, P.surface_flux, ul, ur, o, f);
#pragma unroll
                for (int v = 0; v < NV; ++v) fl[v] = fr[v] = f[v];
            }
        }
    }
    __syncwarp();
    if (g < nvalid) {
        const bool left = s_side[warp][g] == 1;  // the local element takes the flux of its own side
#pragma unroll
        for (int v = 0; v < NV; ++v) s[g * FV + fn * NV + v] = left ? fl[v] : fr[v];
    }
    __syncwarp();
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) {
        if (g2 < nvalid && !(P.mpi_is_piece && P.mpi_is_piece[I0 + g2])) {  // (exchange-only entries of MPI mortars)
            const int o = s_orient[warp][g2];
            const int dir = s_side[warp][g2] == 1 ? 2 * o + 1 : 2 * o;
            double *dst = P.sfv + ((s_elem[warp][g2] * (2 * ND) + dir) * NF) * NV;
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                const int q = lane + 32 * p;
                if (q < FV) dst[q] = s[g2 * FV + q];
            }
        }
    }
}

// ---- 3. element kernel ----------------------------------------------------------------------------
// One thread per node, EPB elements per block.  Shared memory: the block's u tile (coalesced flat copy
// of EPB contiguous element records), the D matrix, and for the weak form the nodal fluxes.
template <class EQ, int N>
struct ElemCfg {
    static constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NN = ipow(N, ND);
    static constexpr int EPB = NN >= 128 ? 1 : 128 / NN;  // elements per block
    static constexpr int THREADS = EPB * NN;
    // node record stride in shared memory (odd number of doubles -> conflict-free 64-bit access)
    static constexpr int US = (NV % 2 == 0) ? NV + 1 : NV;
};

template <class EQ, int N, int VOLINT, bool WITH_SURFACE>
__global__ void __launch_bounds__(ElemCfg<EQ, N>::THREADS) k_element(const KParams P) {
    using C = ElemCfg<EQ, N>;
    constexpr int ND = C::ND, NV = C::NV, NN = C::NN, EPB = C::EPB, US = C::US, NF = ipow(N, ND - 1);
    extern __shared__ double smem[];
    double *s_u = smem;                     // [EPB*NN][US]
    double *s_D = s_u + EPB * NN * US;      // [N*N] column-major
    double *s_f = s_D + N * N;              // weak form: [ND][EPB*NN][US]
    __shared__ double s_red[C::THREADS / 32 > 0 ? C::THREADS / 32 : 1];

    const EQ eq(P.eq);
    const int tid = threadIdx.x;
    const long long e0 = P.elem_begin + (long long)blockIdx.x * EPB;
    const int nel = (int)min((long long)EPB, P.elem_end - e0);

    // stage D and the u tile
    const double *Dsrc = VOLINT == TRIXI_B200_VOLINT_WEAK_FORM ? P.dhat : P.dsplit;
    for (int q = tid; q < N * N; q += C::THREADS) s_D[q] = Dsrc[q];
    {
        const double *src = P.u + e0 * NN * NV;
        const int total = nel * NN * NV;
        for (int q = tid; q < total; q += C::THREADS) {
            const int node = q / NV, v = q - node * NV;
            s_u[node * US + v] = src[q];
        }
    }
    __syncthreads();

    const int le = tid / NN;        // local element
    const int node = tid - le * NN; // node within element
    const bool active = le < nel;
    const long long e = e0 + le;
    int idx[3];
    idx[0] = node % N;
    idx[1] = (node / N) % N;
    idx[2] = ND == 3 ? node / (N * N) : 0;
    const int stride[3] = {1, N, N * N};
    const double *ue = s_u + le * NN * US;

    double un[NV], acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        un[v] = active ? ue[node * US + v] : 0.0;
        acc[v] = 0.0;
    }

    if constexpr (VOLINT == TRIXI_B200_VOLINT_WEAK_FORM) {
        // weak_form_kernel! (dg_3d.jl:133-164): du[:, ii, j, k] += Dhat[ii, i] * flux1(u[i, j, k]) ...
        if (active) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                double f[NV];
                eq.flux(un, d, f);
#pragma unroll
                for (int v = 0; v < NV; ++v) s_f[(d * EPB * NN + le * NN + node) * US + v] = f[v];
            }
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const int base = node - idx[d] * stride[d];
#pragma unroll
                for (int l = 0; l < N; ++l) {
                    const double w = s_D[idx[d] + N * l];  // Dhat[idx_d, l]
                    const double *f = s_f + (d * EPB * NN + le * NN + base + l * stride[d]) * US;
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(w, f[v], acc[v]);
                }
            }
        }
    } else {
        // flux_differencing_kernel! (dg_3d.jl:166-214): the reference evaluates each symmetric pair once
        // as volume_flux(u_lower, u_upper); every node here evaluates its own partners with the same
        // argument order, so both ends see the bit-identical flux.
        if (active) {
            // VolumeIntegralShockCapturingHG (calc_volume_integral.jl:231-272): pure DG where alpha is (almost)
            // zero, otherwise (1 - alpha) flux differencing + alpha subcell finite volumes
            double w_dg = 1.0, w_fv = 0.0;
            constexpr bool kPureFv = VOLINT == TRIXI_B200_VOLINT_PURE_LGL_FV;  // fv_kernel! alone, alpha = true
            constexpr bool kHasFv = kPureFv || VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG;
            if constexpr (VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG) {
                const double alpha = P.alpha[e];
                if (!(fabs(alpha) <= 1.8189894035458565e-12)) {  // isapprox(alpha, 0, atol = max(100 eps, eps^0.75))
                    w_dg = 1 - alpha;
                    w_fv = alpha;
                }
            }
            if constexpr (kPureFv) w_fv = 1.0;
#pragma unroll
            for (int d = 0; d < (kPureFv ? 0 : ND); ++d) {
                const int base = node - idx[d] * stride[d];
#pragma unroll 1
                for (int l = 0; l < N; ++l) {
                    if (l == idx[d]) continue;
                    double up[NV], f[NV];
                    const double *pu = ue + (base + l * stride[d]) * US;
#pragma unroll
                    for (int v = 0; v < NV; ++v) up[v] = pu[v];
                    if (l > idx[d])
                        eq.numflux(P.volume_flux, un, up, d, f);
                    else
                        eq.numflux(P.volume_flux, up, un, d, f);
                    double w = s_D[idx[d] + N * l];  // Dsplit[idx_d, l]
                    if constexpr (VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG) w = w_dg * w;
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(w, f[v], acc[v]);
                }
            }
            if constexpr (kHasFv) {
                // fv_kernel! (dg_3d.jl:268-306): du += alpha sum_d inverse_weights[idx_d] (fstar_L[idx_d + 1] -
                // fstar_R[idx_d]), fstar = volume_flux_fv of neighbouring subcells, zero on the element boundary
                if (w_fv != 0.0) {
                    double sum[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) sum[v] = 0.0;
#pragma unroll 1
                    for (int d = 0; d < ND; ++d) {
                        double fl[NV], fr[NV], up[NV];
#pragma unroll
                        for (int v = 0; v < NV; ++v) fl[v] = fr[v] = 0.0;
                        if (idx[d] > 0) {
                            const double *pu = ue + (node - stride[d]) * US;
#pragma unroll
                            for (int v = 0; v < NV; ++v) up[v] = pu[v];
                            eq.numflux(P.volume_flux_fv, up, un, d, fl);
                            if constexpr (EQ::kHasNoncons) {
                                // calcflux_fv! with nonconservative terms (dg_3d.jl:391-452): fstar_R = flux + 0.5
                                // g(u_rr, u_ll), fstar_L = flux + 0.5 g(u_ll, u_rr) -- the node's own state first
                                if (EQ::has_noncons(P.volume_flux_fv)) {
                                    double g[NV];
                                    eq.noncons(un, up, d, g);
#pragma unroll
                                    for (int v = 0; v < NV; ++v) fl[v] = fl[v] + 0.5 * g[v];
                                }
                            }
                        }
                        if (idx[d] < N - 1) {
                            const double *pu = ue + (node + stride[d]) * US;
#pragma unroll
                            for (int v = 0; v < NV; ++v) up[v] = pu[v];
                            eq.numflux(P.volume_flux_fv, un, up, d, fr);
                            if constexpr (EQ::kHasNoncons) {
                                if (EQ::has_noncons(P.volume_flux_fv)) {
                                    double g[NV];
                                    eq.noncons(un, up, d, g);
#pragma unroll
                                    for (int v = 0; v < NV; ++v) fr[v] = fr[v] + 0.5 * g[v];
                                }
                            }
                        }
                        const double iw = P.inv_weights_c[idx[d]];
#pragma unroll
                        for (int v = 0; v < NV; ++v) sum[v] = fma(iw, fr[v] - fl[v], sum[v]);
                    }
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(w_fv, sum[v], acc[v]);
                }
            }
            if constexpr (EQ::kHasNoncons && !kPureFv) {
                // nonconservative volume terms (dg_3d.jl:216-266): 0.5 * sum_l Dsplit[idx_d, l] g(u, u_l, d)
                if (EQ::has_noncons(P.volume_flux)) {
                    double ic[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) ic[v] = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const int base = node - idx[d] * stride[d];
#pragma unroll 1
                        for (int l = 0; l < N; ++l) {
                            double up[NV], g[NV];
                            const double *pu = ue + (base + l * stride[d]) * US;
#pragma unroll
                            for (int v = 0; v < NV; ++v) up[v] = pu[v];
                            eq.noncons(un, up, d, g);
                            const double w = s_D[idx[d] + N * l];
#pragma unroll
                            for (int v = 0; v < NV; ++v) ic[v] = fma(w, g[v], ic[v]);
                        }
                    }
                    // multiply_add_to_node_vars!(du, alpha * 0.5, integral_contribution, ...) (dg_3d.jl:259-262; alpha =
                    // 1 - blending factor under shock capturing, else 1)
                    const double half = VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG ? w_dg * 0.5 : 0.5;
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(half, ic[v], acc[v]);
                }
            }
        }
    }

    if (!active) return;

    if constexpr (WITH_SURFACE) {
        // calc_surface_integral! (dg_3d.jl:1337-1394): - on the negative faces, + on the positive ones
        const double *sf = P.sfv + e * (2 * ND) * NF * NV;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            int fn;
            if constexpr (ND == 2)
                fn = d == 0 ? idx[1] : idx[0];
            else
                fn = d == 0 ? idx[1] + N * idx[2] : (d == 1 ? idx[0] + N * idx[2] : idx[0] + N * idx[1]);
            if (idx[d] == 0) {
                const double *s = sf + ((2 * d) * NF + fn) * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] = acc[v] - s[v] * P.inv_weight0;
            }
            if (idx[d] == N - 1) {
                const double *s = sf + ((2 * d + 1) * NF + fn) * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] = acc[v] + s[v] * P.inv_weight0;
            }
        }
        // apply_jacobian! (dg_3d.jl:1396-1414)
        const double factor = -P.inverse_jacobian[e];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] *= factor;
        // calc_sources! (dg_3d.jl:1417-1437)
        if (P.source_terms != TRIXI_B200_SRC_NONE) {
            double x[ND], s[NV];
#pragma unroll
            for (int d = 0; d < ND; ++d) x[d] = P.node_coordinates[(e * NN + node) * ND + d];
            eq.source_terms(P.source_terms, un, x, P.t, s);
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] += s[v];
        }
    }

    const long long off = (e * NN + node) * NV;
    if (P.mode == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) P.du[off + v] = acc[v];
    } else if (P.mode == 1) {
        // 2N stage (methods_2N.jl:152-158): u_tmp = du - u_tmp * a; u += u_tmp * (b * dt)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            // first stage: a = 0 and u_tmp = 0 (methods_2N.jl:144), so du - 0 * 0 == du exactly; skipping
            // the read also removes the `u_tmp .= 0` sweep
            const double tmp = P.rk_read_tmp ? acc[v] - P.u_tmp[off + v] * P.rk_a : acc[v];
            P.u_tmp[off + v] = tmp;
            un[v] = un[v] + tmp * P.rk_b_dt;
            P.u_out[off + v] = un[v];
        }
    } else {
        // 3S* / SSP stage
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double xn;
            un[v] = rk_stage_3s_ssp(P, acc[v], P.rk_read_tmp ? P.u_tmp[off + v] : 0.0, un[v],
                                    P.mode == 2 ? P.u_tmp2[off + v] : 0.0, xn);
            if (P.rk_write_tmp) P.u_tmp[off + v] = xn;
            P.u_out[off + v] = un[v];
        }
    }
    (void)s_red;
}

// ---- IndicatorHennemannGassner ------------------------------------------------------------------------
// calc_indicator_hennemann_gassner! (dgsem_tree/indicators_3d.jl:41-131, indicators_2d.jl:26-98): one thread
// per node evaluates the indicator variable and takes part in the dimension-by-dimension nodal -> modal
// transform in shared memory (multiply_scalar_dimensionwise!, interpolation.jl:207-234,348-389); the first
// thread of each element sums the modal energies in the reference's order and applies the logistic map.
template <class EQ, int N>
__global__ void __launch_bounds__(ElemCfg<EQ, N>::THREADS) k_indicator_hg(const KParams P, double threshold,
                                                                          double parameter_s) {
    using C = ElemCfg<EQ, N>;
    constexpr int ND = C::ND, NV = C::NV, NN = C::NN, EPB = C::EPB;
    __shared__ double s_a[EPB * NN], s_b[EPB * NN], s_V[N * N];
    const EQ eq(P.eq);
    const int tid = threadIdx.x;
    const long long e0 = (long long)blockIdx.x * EPB;
    const int le = tid / NN, node = tid - le * NN;
    const long long e = e0 + le;
    const bool active = e < P.nelements;
    for (int q = tid; q < N * N; q += C::THREADS) s_V[q] = P.inv_vdm[q];
    if (active) {
        double un[NV];
        const double *pu = P.u + (e * NN + node) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) un[v] = pu[v];
        s_a[tid] = eq.indicator_variable(P.ind_var, un);
    }
    __syncthreads();
    const int i = node % N, j = (node / N) % N, k = ND == 3 ? node / (N * N) : 0;
    double *a = s_a + le * NN, *b = s_b + le * NN;
    if (active) {
        double res = 0.0;
#pragma unroll
        for (int ii = 0; ii < N; ++ii) res = fma(s_V[i + N * ii], a[ii + N * (j + N * k)], res);
        b[node] = res;
    }
    __syncthreads();
    if (active) {
        double res = 0.0;
#pragma unroll
        for (int jj = 0; jj < N; ++jj) res = fma(s_V[j + N * jj], b[i + N * (jj + N * k)], res);
        a[node] = res;
    }
    __syncthreads();
    const double *modal = a;
    if constexpr (ND == 3) {
        if (active) {
            double res = 0.0;
#pragma unroll
            for (int kk = 0; kk < N; ++kk) res = fma(s_V[k + N * kk], a[i + N * (j + N * kk)], res);
            b[node] = res;
        }
        __syncthreads();
        modal = b;
    }
    if (!active || node != 0) return;
    auto sq = [&](int x, int y, int z) {
        const double m = modal[x + N * (y + N * z)];
        return m * m;
    };
    double clip2 = 0.0, clip1, total;
    if constexpr (ND == 3) {
        for (int z = 0; z < N - 2; ++z)
            for (int y = 0; y < N - 2; ++y)
                for (int x = 0; x < N - 2; ++x) clip2 += sq(x, y, z);
        clip1 = clip2;
        for (int y = 0; y < N - 1; ++y)
            for (int x = 0; x < N - 1; ++x) clip1 += sq(x, y, N - 2);
        for (int z = 0; z < N - 2; ++z)
            for (int x = 0; x < N - 1; ++x) clip1 += sq(x, N - 2, z);
        for (int z = 0; z < N - 2; ++z)
            for (int y = 0; y < N - 2; ++y) clip1 += sq(N - 2, y, z);
        total = clip1;
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) total += sq(x, y, N - 1);
        for (int z = 0; z < N - 1; ++z)
            for (int x = 0; x < N; ++x) total += sq(x, N - 1, z);
        for (int z = 0; z < N - 1; ++z)
            for (int y = 0; y < N - 1; ++y) total += sq(N - 1, y, z);
    } else {
        for (int y = 0; y < N - 2; ++y)
            for (int x = 0; x < N - 2; ++x) clip2 += sq(x, y, 0);
        clip1 = clip2;
        for (int x = 0; x < N - 1; ++x) clip1 += sq(x, N - 2, 0);
        for (int y = 0; y < N - 2; ++y) clip1 += sq(N - 2, y, 0);
        total = clip1;
        for (int x = 0; x < N; ++x) total += sq(x, N - 1, 0);
        for (int y = 0; y < N - 1; ++y) total += sq(N - 1, y, 0);
    }
    const double frac1 = total != 0.0 ? (total - clip1) / total : 0.0;
    const double frac2 = clip1 != 0.0 ? (clip1 - clip2) / clip1 : 0.0;
    const double energy = fmax(frac1, frac2);
    double alpha = 1 / (1 + exp(-parameter_s / threshold * (energy - threshold)));
    if (alpha < P.ind_alpha_min) alpha = 0.0;
    if (alpha > 1 - P.ind_alpha_min) alpha = 1.0;
    alpha = fmin(P.ind_alpha_max, alpha);
    P.alpha_raw[e] = alpha;
    P.alpha[e] = alpha;
}

// The same indicator for 3D, polydeg 3 with a warp per element (8 elements per block): lanes read the element record
// coalesced (two nodes each), the modal transform runs along x, y, z in two 64-entry shared tiles with the inverse
// Vandermonde matrix in registers-by-broadcast, and the three energies are warp reductions (the one-thread-per-node
// kernel above leaves the 64 + 37 + 8 squared coefficients to one thread).  Summation order differs from the
// reference's sequential sums in the last bits.
template <class EQ>
__global__ void __launch_bounds__(256) k_indicator_hg_3d_p3(const KParams P, double threshold, double parameter_s) {
    constexpr int NV = EQ::NVARS;
    __shared__ double s_a[8][64], s_b[8][64], s_V[16];
    const EQ eq(P.eq);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long e = (long long)blockIdx.x * 8 + warp;
    if (threadIdx.x < 16) s_V[threadIdx.x] = P.inv_vdm[threadIdx.x];
    __syncthreads();
    if (e >= P.nelements) return;  // (whole warps leave; no block barrier below)
    double *sa = s_a[warp], *sb = s_b[warp];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = lane + 32 * r;
        double un[NV];
        const double *pu = P.u + (e * 64 + n) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) un[v] = pu[v];
        sa[n] = eq.indicator_variable(P.ind_var, un);
    }
    __syncwarp();
    double m2[2] = {0.0, 0.0};
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const double *src = pass == 1 ? sb : sa;
        double *dst = pass == 1 ? sa : sb;
        const int st = pass == 0 ? 1 : (pass == 1 ? 4 : 16);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int n = lane + 32 * r;
            const int c = (n / st) & 3, b0 = n - c * st;
            double res = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) res = fma(s_V[c + 4 * q], src[b0 + q * st], res);
            if (pass < 2)
                dst[n] = res;
            else
                m2[r] = res * res;
        }
        if (pass < 2) __syncwarp();
    }
    double total = 0.0, clip1 = 0.0, clip2 = 0.0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = lane + 32 * r;
        const int mx = max(n & 3, max((n >> 2) & 3, n >> 4));  // highest mode index of this coefficient
        total += m2[r];
        if (mx <= 2) clip1 += m2[r];
        if (mx <= 1) clip2 += m2[r];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        total += __shfl_xor_sync(0xffffffffu, total, off);
        clip1 += __shfl_xor_sync(0xffffffffu, clip1, off);
        clip2 += __shfl_xor_sync(0xffffffffu, clip2, off);
    }
    if (lane == 0) {
        const double frac1 = total != 0.0 ? (total - clip1) / total : 0.0;
        const double frac2 = clip1 != 0.0 ? (clip1 - clip2) / clip1 : 0.0;
        const double energy = fmax(frac1, frac2);
        double alpha = 1 / (1 + exp(-parameter_s / threshold * (energy - threshold)));
        if (alpha < P.ind_alpha_min) alpha = 0.0;
        if (alpha > 1 - P.ind_alpha_min) alpha = 1.0;
        alpha = fmin(P.ind_alpha_max, alpha);
        P.alpha_raw[e] = alpha;
        P.alpha[e] = alpha;
    }
}

// apply_smoothing! (indicators_3d.jl:133-186): alpha[e] = max(alpha_raw[e], 0.5 alpha_raw[neighbours]); the
// reference's sequential loop is a pure max, so the order does not matter; non-negative doubles order like
// their bit patterns
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_indicator_smooth(const KParams P) {
    constexpr int NS = 1 << (EQ::NDIMS - 1);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long *key = reinterpret_cast<unsigned long long *>(P.alpha);
    if (gid < P.ninterfaces) {
        const long long l = P.if_neighbors[2 * gid] - 1, r = P.if_neighbors[2 * gid + 1] - 1;
        atomicMax(key + l, (unsigned long long)__double_as_longlong(0.5 * P.alpha_raw[r]));
        atomicMax(key + r, (unsigned long long)__double_as_longlong(0.5 * P.alpha_raw[l]));
    } else if (gid - P.ninterfaces < P.nmortars) {
        const long long *ids = P.mortar_ids + (NS + 1) * (gid - P.ninterfaces);
        const long long large = ids[NS] - 1;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            const long long sm = ids[q] - 1;
            atomicMax(key + sm, (unsigned long long)__double_as_longlong(0.5 * P.alpha_raw[large]));
            atomicMax(key + large, (unsigned long long)__double_as_longlong(0.5 * P.alpha_raw[sm]));
        }
    } else if (P.recv_alpha && gid - P.ninterfaces - P.nmortars < P.nmpi) {
        // faces shared with other ranks (apply_smoothing! over mpi_interfaces, indicators_2d.jl:140-185 in the
        // reference's parallel TreeMesh): the neighbour's alpha arrived with its face state
        const long long I = gid - P.ninterfaces - P.nmortars;
        atomicMax(key + (P.mpi_local[I] - 1), (unsigned long long)__double_as_longlong(0.5 * P.recv_alpha[I]));
    }
}

// ---- max_dt ------------------------------------------------------------------------------------------
// stepsize_dg3d.jl:8-32: per element, maxima over nodes of each directional speed separately, summed,
// times inverse_jacobian; global max via atomicMax on the ordered bit pattern.
template <class EQ, int N>
__global__ void __launch_bounds__(ElemCfg<EQ, N>::THREADS) k_max_dt(const KParams P) {
    using C = ElemCfg<EQ, N>;
    constexpr int ND = C::ND, NV = C::NV, NN = C::NN, EPB = C::EPB;
    __shared__ unsigned long long s_lam[ND][EPB];
    const EQ eq(P.eq);
    const int tid = threadIdx.x;
    const long long e0 = (long long)blockIdx.x * EPB;
    const int le = tid / NN, node = tid - le * NN;
    const long long e = e0 + le;
    if (tid < ND * EPB) (&s_lam[0][0])[tid] = 0ull;
    __syncthreads();
    if (e < P.nelements) {
        double un[NV], lam[ND];
        const double *pu = P.u + (e * NN + node) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) un[v] = pu[v];
        eq.max_abs_speeds(un, lam);
#pragma unroll
        for (int d = 0; d < ND; ++d) atomicMax(&s_lam[d][le], cfl_encode(lam[d]));
    }
    __syncthreads();
    if (tid < EPB && e0 + tid < P.nelements) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) s += __longlong_as_double((long long)s_lam[d][tid]);
        const double val = P.inverse_jacobian[e0 + tid] * s;
        atomicMax(P.cfl_key + (blockIdx.x & (kCflSlots - 1)), cfl_encode(val));
    }
}


// ---- analysis: L2 / Linf errors on the device -------------------------------------------------------
// calc_error_norms (analysis_dg3d.jl:123-216, analysis_dg2d.jl:133-215): u and x (and 1/inverse_jacobian on
// curved meshes) are interpolated to the analysis nodes with the Vandermonde matrix, the registered initial
// condition gives the exact solution, and sum_nodes w J diff^2, max |diff| and sum_nodes w J are reduced.
// One block per element; every thread evaluates analysis nodes by direct tensor-product sums.
struct NormParams {
    int na, ic;
    double t;
    const double *vandermonde;  // [na, n] column-major
    const double *weights;      // [na]
    double *sums;               // [nvars + 1]: squared errors, volume
    unsigned long long *linf;   // [nvars] ordered bit patterns
};
constexpr int kMaxAnalysisNodes = 16;

template <class EQ, int N>
__global__ void __launch_bounds__(128) k_error_norms(const KParams P, const NormParams Q) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NN = ipow(N, ND);
    extern __shared__ __align__(16) double s_dyn[];  // (NV + ND + 1) * NN doubles
    double *s_u = s_dyn, *s_x = s_u + NV * NN, *s_j = s_x + ND * NN;
    __shared__ double s_V[kMaxAnalysisNodes * N], s_w[kMaxAnalysisNodes];
    __shared__ double s_red[4][NV + 1];
    __shared__ unsigned long long s_max[4][NV];
    const long long e = blockIdx.x;
    const int tid = threadIdx.x, na = Q.na;
    for (int q = tid; q < NV * NN; q += 128) s_u[q] = P.u[e * NV * NN + q];
    for (int q = tid; q < ND * NN; q += 128) s_x[q] = P.node_coordinates[e * ND * NN + q];
    if (P.curved)
        for (int q = tid; q < NN; q += 128) s_j[q] = 1.0 / P.inverse_jacobian[e * NN + q];
    for (int q = tid; q < na * N; q += 128) s_V[q] = Q.vandermonde[q];
    for (int q = tid; q < na; q += 128) s_w[q] = Q.weights[q];
    __syncthreads();
    const EQ eq(P.eq);
    double volume_jacobian = 1.0;
    if (!P.curved) {  // volume_jacobian (dgsem_tree/dg.jl:8-10)
        const double j1 = 1.0 / P.inverse_jacobian[e];
#pragma unroll
        for (int d = 0; d < ND; ++d) volume_jacobian *= j1;
    }
    double l2[NV], vol = 0.0;
    unsigned long long mx[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        l2[v] = 0.0;
        mx[v] = 0ull;
    }
    const int total = ND == 2 ? na * na : na * na * na;
    for (int a = tid; a < total; a += 128) {
        const int a0 = a % na, a1 = (a / na) % na, a2 = ND == 3 ? a / (na * na) : 0;
        double ua[NV], xa[ND], ja = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) ua[v] = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) xa[d] = 0.0;
        for (int k = 0; k < (ND == 3 ? N : 1); ++k) {
            const double ck = ND == 3 ? s_V[a2 + na * k] : 1.0;
            for (int j = 0; j < N; ++j) {
                const double cjk = ck * s_V[a1 + na * j];
                for (int i = 0; i < N; ++i) {
                    const double c = cjk * s_V[a0 + na * i];
                    const int node = i + N * (j + N * k);
#pragma unroll
                    for (int v = 0; v < NV; ++v) ua[v] = fma(c, s_u[node * NV + v], ua[v]);
#pragma unroll
                    for (int d = 0; d < ND; ++d) xa[d] = fma(c, s_x[node * ND + d], xa[d]);
                    if (P.curved) ja = fma(c, s_j[node], ja);
                }
            }
        }
        double uex[NV];
        eq.initial_condition(Q.ic, xa, Q.t, uex);
        double w = s_w[a0] * s_w[a1] * (ND == 3 ? s_w[a2] : 1.0);
        w *= P.curved ? fabs(ja) : volume_jacobian;
        vol += w;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double diff = uex[v] - ua[v];
            l2[v] = fma(diff * diff, w, l2[v]);
            mx[v] = max(mx[v], cfl_encode(fabs(diff)));
        }
    }
    // block reduction: warp shuffles, then the four warps through shared memory
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            l2[v] += __shfl_xor_sync(0xffffffffu, l2[v], off);
            mx[v] = max(mx[v], __shfl_xor_sync(0xffffffffu, mx[v], off));
        }
        vol += __shfl_xor_sync(0xffffffffu, vol, off);
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            s_red[warp][v] = l2[v];
            s_max[warp][v] = mx[v];
        }
        s_red[warp][NV] = vol;
    }
    __syncthreads();
    if (tid <= NV) {
        const double sum = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];
        atomicAdd(Q.sums + tid, sum);
        if (tid < NV) atomicMax(Q.linf + tid, max(max(s_max[0][tid], s_max[1][tid]), max(s_max[2][tid], s_max[3][tid])));
    }
}

// integrate_via_indices (analysis_dg3d.jl:364-473): sum over nodes of w_i w_j (w_k) |J| func(u, node), one thread per
// node; the integrands are the AnalysisCallback's analysis_integrals.  sums: [nvals + 1] (values, volume).
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_integrate(const KParams P, int quantity, double *sums) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NN = ipow(N, ND);
    __shared__ double s_red[8][NV + 1];
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gid / NN;
    const int node = (int)(gid - e * NN);
    double val[NV], w = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) val[v] = 0.0;
    if (e < P.nelements) {
        const EQ eq(P.eq);
        const int i = node % N, j = (node / N) % N, k = ND == 3 ? node / (N * N) : 0;
        w = P.weights_c[i] * P.weights_c[j] * (ND == 3 ? P.weights_c[k] : 1.0);
        if (P.curved) {
            w *= fabs(1.0 / P.inverse_jacobian[e * NN + node]);
        } else {  // volume_jacobian (dgsem_tree/dg.jl:8-10)
            const double j1 = 1.0 / P.inverse_jacobian[e];
            double vj = 1.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) vj *= j1;
            w *= vj;
        }
        double un[NV];
        const double *pu = P.u + (e * NN + node) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) un[v] = pu[v];
        if (quantity == TRIXI_B200_INTEGRAL_CONS) {
#pragma unroll
            for (int v = 0; v < NV; ++v) val[v] = w * un[v];
        } else {
            double dun[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) dun[v] = 0.0;
            if (quantity == TRIXI_B200_INTEGRAL_ENTROPY_TIMEDERIVATIVE) {
                const double *pd = P.du + (e * NN + node) * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) dun[v] = pd[v];
            }
            val[0] = w * eq.analysis_integrand(quantity, un, dun);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) val[v] += __shfl_xor_sync(0xffffffffu, val[v], off);
        w += __shfl_xor_sync(0xffffffffu, w, off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) s_red[warp][v] = val[v];
        s_red[warp][NV] = w;
    }
    __syncthreads();
    if (threadIdx.x <= NV) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) sum += s_red[q][threadIdx.x];
        atomicAdd(sums + threadIdx.x, sum);
    }
}

// =====================================================================================================
// Curved meshes (StructuredMesh; src/solvers/dgsem_structured/).  Same launch structure; the geometry
// enters through the contravariant vectors Ja^i[dim] (dg.jl:27-30) and the nodal inverse Jacobian.
// =====================================================================================================
template <int ND, int NN>
TB_DEV void load_ja(const KParams &P, int index, long long node, long long e, double (&ja)[ND]) {
    const double *p = P.contravariant_vectors + ((e * NN + node) * ND + index) * ND;
#pragma unroll
    for (int d = 0; d < ND; ++d) ja[d] = p[d];
}

// prolong2interfaces! + calc_interface_flux! (dgsem_structured/dg_3d.jl:619-753): the normal is the
// contravariant vector of the right element's first node layer times sign(inverse_jacobian)
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_interface_flux_curved(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / NF;
    const int fn = (int)(gid % NF);
    if (I >= P.ninterfaces) return;
    const EQ eq(P.eq);
    const long long left = P.if_neighbors[2 * I] - 1, right = P.if_neighbors[2 * I + 1] - 1;
    const int o = (int)P.if_orient[I] - 1;
    const int nl = face_to_volume_node<ND, N>(o, N - 1, fn), nr = face_to_volume_node<ND, N>(o, 0, fn);
    double ul[NV], ur[NV], f[NV], nrm[ND];
    const double *pl = P.u + (left * NN + nl) * NV, *pr = P.u + (right * NN + nr) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        ul[v] = pl[v];
        ur[v] = pr[v];
    }
    const double ij = P.inverse_jacobian[right * NN + nr];
    const double sign_jacobian = ij > 0 ? 1.0 : (ij < 0 ? -1.0 : 0.0);
    load_ja<ND, NN>(P, o, nr, right, nrm);
#pragma unroll
    for (int d = 0; d < ND; ++d) nrm[d] *= sign_jacobian;
    eq.numflux_normal(P.surface_flux, ul, ur, nrm, f);
    double *sl = P.sfv + ((left * (2 * ND) + (2 * o + 1)) * NF + fn) * NV;
    double *sr = P.sfv + ((right * (2 * ND) + (2 * o)) * NF + fn) * NV;
    if constexpr (EQ::kHasNoncons) {
        if (EQ::has_noncons(P.surface_flux)) {  // dgsem_structured/dg_3d.jl:755-828
            double gl[NV], gr[NV];
            eq.noncons_normal(ul, ur, nrm, gl);
            eq.noncons_normal(ur, ul, nrm, gr);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const double fv = sign_jacobian * f[v];
                sl[v] = fv + 0.5 * (sign_jacobian * gl[v]);
                sr[v] = fv + 0.5 * (sign_jacobian * gr[v]);
            }
            return;
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const double fv = sign_jacobian * f[v];
        sl[v] = fv;
        sr[v] = fv;
    }
}

// calc_boundary_flux! (dgsem_structured/dg_3d.jl:755-935, dg.jl:124-165)
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_boundary_flux_curved(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long B = gid / NF;
    const int fn = (int)(gid % NF);
    if (B >= P.nboundaries) return;
    const EQ eq(P.eq);
    const long long element = P.bd_neighbor[B] - 1;
    const int direction = P.bd_direction[B];  // 1-based
    const int o = (direction - 1) / 2;
    const int vn = face_to_volume_node<ND, N>(o, direction % 2 == 1 ? 0 : N - 1, fn);
    double ui[NV], f[NV], x[ND], nrm[ND];
    const double *pu = P.u + (element * NN + vn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) ui[v] = pu[v];
#pragma unroll
    for (int d = 0; d < ND; ++d) x[d] = P.node_coordinates[(element * NN + vn) * ND + d];
    const double ij = P.inverse_jacobian[element * NN + vn];
    const double sign_jacobian = ij > 0 ? 1.0 : (ij < 0 ? -1.0 : 0.0);
    load_ja<ND, NN>(P, o, vn, element, nrm);
#pragma unroll
    for (int d = 0; d < ND; ++d) nrm[d] *= sign_jacobian;
    boundary_flux_normal(eq, P.bc[direction - 1], P.bc_ic[direction - 1], P.surface_flux, ui, nrm, direction, x, P.t, f);
    double *s = P.sfv + ((element * (2 * ND) + (direction - 1)) * NF + fn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = sign_jacobian * f[v];
}

// element kernel for curved meshes: weak_form_kernel! / flux_differencing_kernel!
// (dgsem_structured/dg_3d.jl:36-175), calc_surface_integral! (dg_3d.jl:1337-1394, shared with TreeMesh),
// apply_jacobian! with the nodal inverse Jacobian (dgsem_structured/dg_3d.jl:937-956), sources, 2N stage
template <class EQ, int N, int VOLINT, bool WITH_SURFACE>
__global__ void __launch_bounds__(ElemCfg<EQ, N>::THREADS) k_element_curved(const KParams P) {
    using C = ElemCfg<EQ, N>;
    constexpr int ND = C::ND, NV = C::NV, NN = C::NN, EPB = C::EPB, US = C::US, NF = ipow(N, ND - 1);
    extern __shared__ double smem[];
    double *s_u = smem;                 // [EPB*NN][US]
    double *s_D = s_u + EPB * NN * US;  // [N*N]
    double *s_f = s_D + N * N;          // weak form: contravariant fluxes [ND][EPB*NN][US]

    const EQ eq(P.eq);
    const int tid = threadIdx.x;
    const long long e0 = (long long)blockIdx.x * EPB;
    const int nel = (int)min((long long)EPB, P.nelements - e0);
    const double *Dsrc = VOLINT == TRIXI_B200_VOLINT_WEAK_FORM ? P.dhat : P.dsplit;
    for (int q = tid; q < N * N; q += C::THREADS) s_D[q] = Dsrc[q];
    {
        const double *src = P.u + e0 * NN * NV;
        const int total = nel * NN * NV;
        for (int q = tid; q < total; q += C::THREADS) {
            const int node = q / NV, v = q - node * NV;
            s_u[node * US + v] = src[q];
        }
    }
    __syncthreads();

    const int le = tid / NN, node = tid - le * NN;
    const bool active = le < nel;
    const long long e = e0 + le;
    int idx[3];
    idx[0] = node % N;
    idx[1] = (node / N) % N;
    idx[2] = ND == 3 ? node / (N * N) : 0;
    const int stride[3] = {1, N, N * N};
    const double *ue = s_u + le * NN * US;
    double un[NV], acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        un[v] = active ? ue[node * US + v] : 0.0;
        acc[v] = 0.0;
    }

    if constexpr (VOLINT == TRIXI_B200_VOLINT_WEAK_FORM) {
        if (active) {
            double fl[ND][NV];
#pragma unroll
            for (int d = 0; d < ND; ++d) eq.flux(un, d, fl[d]);
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                double ja[ND];
                load_ja<ND, NN>(P, a, node, e, ja);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    double sum = ja[0] * fl[0][v];
#pragma unroll
                    for (int d = 1; d < ND; ++d) sum += ja[d] * fl[d][v];
                    s_f[(a * EPB * NN + le * NN + node) * US + v] = sum;
                }
            }
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const int base = node - idx[d] * stride[d];
#pragma unroll
                for (int l = 0; l < N; ++l) {
                    const double w = s_D[idx[d] + N * l];
                    const double *f = s_f + (d * EPB * NN + le * NN + base + l * stride[d]) * US;
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(w, f[v], acc[v]);
                }
            }
        }
    } else {
        if (active) {
            // VolumeIntegralShockCapturingHG (calc_volume_integral.jl:231-272) as in k_element
            double w_dg = 1.0, w_fv = 0.0;
            constexpr bool kPureFv = VOLINT == TRIXI_B200_VOLINT_PURE_LGL_FV;  // fv_kernel! alone, alpha = true
            constexpr bool kHasFv = kPureFv || VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG;
            if constexpr (VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG) {
                const double alpha = P.alpha[e];
                if (!(fabs(alpha) <= 1.8189894035458565e-12)) {
                    w_dg = 1 - alpha;
                    w_fv = alpha;
                }
            }
            if constexpr (kPureFv) w_fv = 1.0;
#pragma unroll
            for (int d = 0; d < (kPureFv ? 0 : ND); ++d) {
                const int base = node - idx[d] * stride[d];
                double ja_node[ND];
                load_ja<ND, NN>(P, d, node, e, ja_node);
#pragma unroll 1
                for (int l = 0; l < N; ++l) {
                    if (l == idx[d]) continue;
                    const int node2 = base + l * stride[d];
                    double up[NV], f[NV], ja2[ND], ja_avg[ND];
                    const double *pu = ue + node2 * US;
#pragma unroll
                    for (int v = 0; v < NV; ++v) up[v] = pu[v];
                    load_ja<ND, NN>(P, d, node2, e, ja2);
                    // 0.5 * (Ja_lower + Ja_upper): same operand order on both ends
#pragma unroll
                    for (int q = 0; q < ND; ++q) ja_avg[q] = l > idx[d] ? 0.5 * (ja_node[q] + ja2[q]) : 0.5 * (ja2[q] + ja_node[q]);
                    if (l > idx[d])
                        eq.numflux_normal(P.volume_flux, un, up, ja_avg, f);
                    else
                        eq.numflux_normal(P.volume_flux, up, un, ja_avg, f);
                    double w = s_D[idx[d] + N * l];
                    if constexpr (VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG) w = w_dg * w;
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(w, f[v], acc[v]);
                }
            }
            if constexpr (EQ::kHasNoncons && !kPureFv) {
                // nonconservative volume terms on curved meshes (dgsem_structured/dg_3d.jl:177-283):
                // 0.5 sum_d sum_l Dsplit[idx_d, l] g(u, u_l, 0.5 (Ja^d + Ja^d_l))
                if (EQ::has_noncons(P.volume_flux)) {
                    double ic[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) ic[v] = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const int base = node - idx[d] * stride[d];
                        double ja_node[ND];
                        load_ja<ND, NN>(P, d, node, e, ja_node);
#pragma unroll 1
                        for (int l = 0; l < N; ++l) {
                            const int node2 = base + l * stride[d];
                            double up[NV], g[NV], ja2[ND], ja_avg[ND];
                            const double *pu = ue + node2 * US;
#pragma unroll
                            for (int v = 0; v < NV; ++v) up[v] = pu[v];
                            load_ja<ND, NN>(P, d, node2, e, ja2);
#pragma unroll
                            for (int q = 0; q < ND; ++q) ja_avg[q] = 0.5 * (ja_node[q] + ja2[q]);
                            eq.noncons_normal(un, up, ja_avg, g);
                            const double w = s_D[idx[d] + N * l];
#pragma unroll
                            for (int v = 0; v < NV; ++v) ic[v] = fma(w, g[v], ic[v]);
                        }
                    }
                    const double half = VOLINT == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG ? w_dg * 0.5 : 0.5;
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(half, ic[v], acc[v]);
                }
            }
            if constexpr (kHasFv) {
                // fv_kernel! (dg_3d.jl:268-306) with calcflux_fv! for curved meshes (dgsem_structured/dg_3d.jl:377-436):
                // subcell fluxes along the precomputed free-stream preserving normal vectors, zero on the element boundary
                if (w_fv != 0.0) {
                    double sum[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) sum[v] = 0.0;
#pragma unroll 1
                    for (int d = 0; d < ND; ++d) {
                        // normal_vectors_d [ND, dims.., nelements] with N - 1 entries along direction d
                        int dims[3] = {N, N, ND == 3 ? N : 1};
                        dims[d] = N - 1;
                        const long long per_elem = (long long)dims[0] * dims[1] * dims[2];
                        const double *nvec = P.subcell_normals[d] + ND * per_elem * e;
                        double fl[NV], fr[NV], up[NV], nrm[ND];
#pragma unroll
                        for (int v = 0; v < NV; ++v) fl[v] = fr[v] = 0.0;
                        int pos[3] = {idx[0], idx[1], idx[2]};
                        if (idx[d] > 0) {
                            pos[d] = idx[d] - 1;
                            const long long q = pos[0] + (long long)dims[0] * (pos[1] + (long long)dims[1] * pos[2]);
#pragma unroll
                            for (int c = 0; c < ND; ++c) nrm[c] = nvec[c + ND * q];
                            const double *pu = ue + (node - stride[d]) * US;
#pragma unroll
                            for (int v = 0; v < NV; ++v) up[v] = pu[v];
                            eq.numflux_normal(P.volume_flux_fv, up, un, nrm, fl);
                            if constexpr (EQ::kHasNoncons) {
                                // calcflux_fv! with nonconservative terms (dgsem_structured/dg_3d.jl:438-530): ftilde_R =
                                // ftilde + 0.5 g(u_rr, u_ll, n), ftilde_L = ftilde + 0.5 g(u_ll, u_rr, n)
                                if (EQ::has_noncons(P.volume_flux_fv)) {
                                    double g[NV];
                                    eq.noncons_normal(un, up, nrm, g);
#pragma unroll
                                    for (int v = 0; v < NV; ++v) fl[v] = fl[v] + 0.5 * g[v];
                                }
                            }
                        }
                        if (idx[d] < N - 1) {
                            pos[d] = idx[d];
                            const long long q = pos[0] + (long long)dims[0] * (pos[1] + (long long)dims[1] * pos[2]);
#pragma unroll
                            for (int c = 0; c < ND; ++c) nrm[c] = nvec[c + ND * q];
                            const double *pu = ue + (node + stride[d]) * US;
#pragma unroll
                            for (int v = 0; v < NV; ++v) up[v] = pu[v];
                            eq.numflux_normal(P.volume_flux_fv, un, up, nrm, fr);
                            if constexpr (EQ::kHasNoncons) {
                                if (EQ::has_noncons(P.volume_flux_fv)) {
                                    double g[NV];
                                    eq.noncons_normal(un, up, nrm, g);
#pragma unroll
                                    for (int v = 0; v < NV; ++v) fr[v] = fr[v] + 0.5 * g[v];
                                }
                            }
                        }
                        const double iw = P.inv_weights_c[idx[d]];
#pragma unroll
                        for (int v = 0; v < NV; ++v) sum[v] = fma(iw, fr[v] - fl[v], sum[v]);
                    }
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = fma(w_fv, sum[v], acc[v]);
                }
            }
        }
    }
    if (!active) return;

    if constexpr (WITH_SURFACE) {
        const double *sf = P.sfv + e * (2 * ND) * NF * NV;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            int fn;
            if constexpr (ND == 2)
                fn = d == 0 ? idx[1] : idx[0];
            else
                fn = d == 0 ? idx[1] + N * idx[2] : (d == 1 ? idx[0] + N * idx[2] : idx[0] + N * idx[1]);
            if (idx[d] == 0) {
                // Structured: "-" on negative faces (dg_3d.jl:1337-1394); P4est: "+" everywhere because the
                // fluxes are taken along outward normals (dgsem_p4est/dg_3d.jl:976-1034)
                const double *sq = sf + ((2 * d) * NF + fn) * NV;
                const double w = P.p4est ? P.inv_weight0 : -P.inv_weight0;
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] = acc[v] + sq[v] * w;
            }
            if (idx[d] == N - 1) {
                const double *sq = sf + ((2 * d + 1) * NF + fn) * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] = acc[v] + sq[v] * P.inv_weight0;
            }
        }
        const double factor = -P.inverse_jacobian[e * NN + node];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] *= factor;
        if (P.source_terms != TRIXI_B200_SRC_NONE) {
            double x[ND], sv[NV];
#pragma unroll
            for (int d = 0; d < ND; ++d) x[d] = P.node_coordinates[(e * NN + node) * ND + d];
            eq.source_terms(P.source_terms, un, x, P.t, sv);
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] += sv[v];
        }
    }
    const long long off = (e * NN + node) * NV;
    if (P.mode == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) P.du[off + v] = acc[v];
    } else if (P.mode == 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double tmp = P.rk_read_tmp ? acc[v] - P.u_tmp[off + v] * P.rk_a : acc[v];
            P.u_tmp[off + v] = tmp;
            P.u_out[off + v] = un[v] + tmp * P.rk_b_dt;
        }
    } else {
        // 3S* / SSP stage
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double xn;
            const double unew = rk_stage_3s_ssp(P, acc[v], P.rk_read_tmp ? P.u_tmp[off + v] : 0.0, un[v],
                                                P.mode == 2 ? P.u_tmp2[off + v] : 0.0, xn);
            if (P.rk_write_tmp) P.u_tmp[off + v] = xn;
            P.u_out[off + v] = unew;
        }
    }
}

// max_dt for curved meshes (stepsize_dg3d.jl:79-123): per node inv_jacobian * |Ja^i . lambda|, per-direction
// maxima over the element's nodes, their sum, global max
template <class EQ, int N>
__global__ void __launch_bounds__(ElemCfg<EQ, N>::THREADS) k_max_dt_curved(const KParams P) {
    using C = ElemCfg<EQ, N>;
    constexpr int ND = C::ND, NV = C::NV, NN = C::NN, EPB = C::EPB;
    __shared__ unsigned long long s_lam[ND][EPB];
    const EQ eq(P.eq);
    const int tid = threadIdx.x;
    const long long e0 = (long long)blockIdx.x * EPB;
    const int le = tid / NN, node = tid - le * NN;
    const long long e = e0 + le;
    if (tid < ND * EPB) (&s_lam[0][0])[tid] = 0ull;
    __syncthreads();
    if (e < P.nelements) {
        double un[NV], lam[ND];
        const double *pu = P.u + (e * NN + node) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) un[v] = pu[v];
        eq.max_abs_speeds(un, lam);
        const double inv_jacobian = fabs(P.inverse_jacobian[e * NN + node]);
        double node_sum = 0.0;
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            double ja[ND];
            load_ja<ND, NN>(P, a, node, e, ja);
            double sum = ja[0] * lam[0];
#pragma unroll
            for (int d = 1; d < ND; ++d) sum += ja[d] * lam[d];
            if constexpr (EQ::kConstantSpeed)
                node_sum += fabs(sum);
            else
                atomicMax(&s_lam[a][le], cfl_encode(inv_jacobian * fabs(sum)));
        }
        // constant_speed::True (max_scaled_speed_per_element, stepsize_dg3d.jl:176-210, stepsize_dg2d.jl:207-235): the
        // maximum over the nodes of inv_jacobian * (sum of the transformed speeds), not the sum of per-direction maxima
        if constexpr (EQ::kConstantSpeed) atomicMax(&s_lam[0][le], cfl_encode(inv_jacobian * node_sum));
    }
    __syncthreads();
    if (tid < EPB && e0 + tid < P.nelements) {
        double sum = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) sum += __longlong_as_double((long long)s_lam[d][tid]);
        atomicMax(P.cfl_key + (blockIdx.x & (kCflSlots - 1)), cfl_encode(sum));
    }
}

// =====================================================================================================
// P4estMesh (src/solvers/dgsem_p4est/): unstructured conforming hexahedra/quadrilaterals.  Volume terms,
// Jacobian and max_dt are the curved kernels above; faces are addressed through symbolic node_indices.
// =====================================================================================================
TB_DEV int p4_index(int sym, int n, int i, int j) {
    // index_to_start_step_3d (dg_3d.jl:61-80) in closed form; :begin 0 :end 1 :i_forward 2 :i_backward 3
    // :j_forward 4 :j_backward 5
    return sym == 0 ? 0 : sym == 1 ? n - 1 : sym == 2 ? i : sym == 3 ? n - 1 - i : sym == 4 ? j : n - 1 - j;
}
template <int ND, int N>
TB_DEV void p4_face(const long long *idx, int i, int j, int &volume_node, int &surface_node, int &direction0) {
    int node = 0, stride = 1, s0 = 0, s1 = 0, k = 0;
    direction0 = 0;
#pragma unroll
    for (int c = 0; c < ND; ++c) {
        const int sym = (int)idx[c];
        const int q = p4_index(sym, N, i, j);
        node += stride * q;
        stride *= N;
        if (sym == 0) direction0 = 2 * c;
        if (sym == 1) direction0 = 2 * c + 1;
        if (sym > 1) {
            if (k == 0)
                s0 = q;
            else
                s1 = q;
            ++k;
        }
    }
    volume_node = node;
    surface_node = ND == 3 ? s0 + N * s1 : s0;
}

// prolong2interfaces! + calc_interface_flux! (dgsem_p4est/dg_3d.jl:94-314): outward normal of the primary
// element (dg.jl:74-86), +flux to the primary, -flux to the secondary element
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_interface_flux_p4est(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / NF;
    const int fn = (int)(gid % NF);
    if (I >= P.ninterfaces) return;
    const EQ eq(P.eq);
    const int i = fn % N, j = fn / N;
    const long long primary = P.if_neighbors[2 * I] - 1, secondary = P.if_neighbors[2 * I + 1] - 1;
    int pn, pfn, pdir, sn, sfn, sdir;
    p4_face<ND, N>(P.if_node_indices + (2 * I + 0) * ND, i, j, pn, pfn, pdir);
    p4_face<ND, N>(P.if_node_indices + (2 * I + 1) * ND, i, j, sn, sfn, sdir);
    double ul[NV], ur[NV], f[NV], nrm[ND];
    const double *pl = P.u + (primary * NN + pn) * NV, *pr = P.u + (secondary * NN + sn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        ul[v] = pl[v];
        ur[v] = pr[v];
    }
    load_ja<ND, NN>(P, pdir / 2, pn, primary, nrm);
    if (pdir % 2 == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) nrm[d] = -nrm[d];
    }
    eq.numflux_normal(P.surface_flux, ul, ur, nrm, f);
    double *sp = P.sfv + ((primary * (2 * ND) + pdir) * NF + fn) * NV;
    double *ss = P.sfv + ((secondary * (2 * ND) + sdir) * NF + sfn) * NV;
    if constexpr (EQ::kHasNoncons) {
        if (EQ::has_noncons(P.surface_flux)) {  // dgsem_p4est/dg_3d.jl:340-375
            double gp[NV], gs[NV];
            eq.noncons_normal(ul, ur, nrm, gp);
            eq.noncons_normal(ur, ul, nrm, gs);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                sp[v] = f[v] + 0.5 * gp[v];
                ss[v] = -(f[v] + 0.5 * gs[v]);
            }
            return;
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        sp[v] = f[v];
        ss[v] = -f[v];
    }
}

// prolong2mortars! + calc_mortar_flux! + mortar_fluxes_to_elements! (dgsem_p4est/dg_2d.jl:802-1055,
// dg_3d.jl:651-974), conservative equations, fused like k_mortar_flux: one block per mortar, one thread per
// (position, face node).  The mortar is aligned at the small side (its node_indices always run forward); the
// large face is read and written through its own node_indices.  The flux is taken along the outward normal of
// the small element; the large element receives the L2 projection with the sign switched and scaled by the
// ratio of the face areas, 2^(d-1).
template <class EQ, int N>
__global__ void __launch_bounds__((1 << (EQ::NDIMS - 1)) * ipow(N, EQ::NDIMS - 1)) k_mortar_flux_p4est(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND), NP = 1 << (ND - 1);
    __shared__ double s_large[NV * NF];
    __shared__ double s_tmp[NP][NV * NF];
    __shared__ double s_f[NP][NV * NF];
    const long long m = blockIdx.x;
    const int tid = threadIdx.x;
    const int p = tid / NF, fn = tid - p * NF;
    const int a = fn % N, b = fn / N;
    const EQ eq(P.eq);
    const long long *ids = P.mortar_ids + (NP + 1) * m;
    const long long *sidx = P.mortar_node_indices + (2 * m + 0) * ND, *lidx = P.mortar_node_indices + (2 * m + 1) * ND;
    const long long large = ids[NP] - 1, small = ids[p] - 1;
    const double *fwd1 = P.mortar_fwd[p & 1], *fwd2 = P.mortar_fwd[(p >> 1) & 1];
    const double *rev1 = P.mortar_rev[p & 1];
    // the large face in the orientation of the mortar
    for (int q = tid; q < NV * NF; q += NP * NF) {
        const int f = q / NV, v = q - f * NV;
        int vn, sfn, dir;
        p4_face<ND, N>(lidx, f % N, f / N, vn, sfn, dir);
        s_large[q] = P.u[(large * NN + vn) * NV + v];
    }
    __syncthreads();
    // interpolation to position p, first face coordinate first (multiply_dimensionwise! interpolation.jl:237-264)
    double up[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double acc = 0.0;
        for (int q = 0; q < N; ++q) acc += fwd1[a + N * q] * s_large[v + NV * (q + N * b)];
        up[v] = acc;
    }
    if constexpr (ND == 3) {
#pragma unroll
        for (int v = 0; v < NV; ++v) s_tmp[p][v + NV * fn] = up[v];
        __syncthreads();
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double acc = 0.0;
            for (int q = 0; q < N; ++q) acc += fwd2[b + N * q] * s_tmp[p][v + NV * (a + N * q)];
            up[v] = acc;
        }
        __syncthreads();  // s_tmp is reused by the projection below
    }
    // flux(u_small, u_large interpolated, outward normal of the small element)
    double us[NV], f[NV], nrm[ND];
    int svn, ssfn, sdir;
    p4_face<ND, N>(sidx, a, b, svn, ssfn, sdir);
    {
        const double *pu = P.u + (small * NN + svn) * NV;
#pragma unroll
        for (int v = 0; v < NV; ++v) us[v] = pu[v];
    }
    load_ja<ND, NN>(P, sdir / 2, svn, small, nrm);
    if (sdir % 2 == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) nrm[d] = -nrm[d];
    }
    eq.numflux_normal(P.surface_flux, us, up, nrm, f);
    {
        double *dst = P.sfv + ((small * (2 * ND) + sdir) * NF + fn) * NV;
        bool done = false;
        if constexpr (EQ::kHasNoncons) {
            if (EQ::has_noncons(P.surface_flux)) {  // dg_3d.jl:860-888: 0.5 g(u_ll, u_rr) primary, 0.5 g(u_rr, u_ll) secondary
                double gp[NV], gs[NV];
                eq.noncons_normal(us, up, nrm, gp);
                eq.noncons_normal(up, us, nrm, gs);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    dst[v] = f[v] + 0.5 * gp[v];
                    s_f[p][v + NV * fn] = f[v] + 0.5 * gs[v];
                }
                done = true;
            }
        }
        if (!done) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                dst[v] = f[v];
                s_f[p][v + NV * fn] = f[v];
            }
        }
    }
    __syncthreads();
    // L2 projection onto the large face
    int lvn, lsfn, ldir;
    p4_face<ND, N>(lidx, a, b, lvn, lsfn, ldir);
    double *out = P.sfv + ((large * (2 * ND) + ldir) * NF + lsfn) * NV;
    if constexpr (ND == 2) {
        if (tid < N) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double acc = 0.0;
                for (int q = 0; q < N; ++q)
                    acc += P.mortar_rev[1][tid + N * q] * s_f[1][v + NV * q] + P.mortar_rev[0][tid + N * q] * s_f[0][v + NV * q];
                out[v] = acc * -2.0;
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double acc = 0.0;
            for (int q = 0; q < N; ++q) acc += rev1[a + N * q] * s_f[p][v + NV * (q + N * b)];
            s_tmp[p][v + NV * fn] = acc;
        }
        __syncthreads();
        if (tid < NF) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double res = 0.0;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {  // positions 1..4 in this order (dg_3d.jl:914-929)
                    const double *r2 = P.mortar_rev[(pp >> 1) & 1];
                    double acc = 0.0;
                    for (int q = 0; q < N; ++q) acc += r2[b + N * q] * s_tmp[pp][v + NV * (a + N * q)];
                    res = pp == 0 ? acc : res + acc;
                }
                out[v] = res * -4.0;
            }
        }
    }
}

// (A shared-memory staged variant like k_interface_flux_staged was measured and dropped: 0.83 ms against 0.795 ms
// per launch at 16.8 M DOF -- with the LLF flux along a normal the kernel is bound by its FP64 divisions and
// square roots and the scattered 24-byte normal loads, not by the face gathers.)

// prolong2boundaries! + calc_boundary_flux! (dgsem_p4est/dg_3d.jl:412-548)
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_boundary_flux_p4est(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long B = gid / NF;
    const int fn = (int)(gid % NF);
    if (B >= P.nboundaries) return;
    const EQ eq(P.eq);
    const int i = fn % N, j = fn / N;
    const long long element = P.bd_neighbor[B] - 1;
    const int name = P.bd_direction[B] - 1;  // boundaries sorted by name = direction
    int vn, sfn, dir;
    p4_face<ND, N>(P.bd_node_indices + B * ND, i, j, vn, sfn, dir);
    double ui[NV], f[NV], x[ND], nrm[ND];
    const double *pu = P.u + (element * NN + vn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) ui[v] = pu[v];
#pragma unroll
    for (int d = 0; d < ND; ++d) x[d] = P.node_coordinates[(element * NN + vn) * ND + d];
    load_ja<ND, NN>(P, dir / 2, vn, element, nrm);
    if (dir % 2 == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) nrm[d] = -nrm[d];
    }
    const int bc = P.bc[name];
    if (bc == TRIXI_B200_BC_DIRICHLET) {  // equations.jl:206-228: flux(u_inner, u_boundary, outward normal)
        double ub[NV];
        eq.initial_condition(P.bc_ic[name], x, P.t, ub);
        eq.numflux_normal(P.surface_flux, ui, ub, nrm, f);
        if constexpr (EQ::kHasNoncons) {
            if (EQ::has_noncons(P.surface_flux)) {  // equations.jl:232-247: flux + 0.5 g(u_inner, u_boundary, n)
                double g[NV];
                eq.noncons_normal(ui, ub, nrm, g);
#pragma unroll
                for (int v = 0; v < NV; ++v) f[v] = f[v] + 0.5 * g[v];
            }
        }
    } else if (bc == TRIXI_B200_BC_SLIP_WALL) {
        eq.slip_wall_outward(ui, nrm, f);
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) f[v] = nan("");
    }
    double *s = P.sfv + ((element * (2 * ND) + dir) * NF + fn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = f[v];
}

// prolong2mpiinterfaces! (dgsem_p4est/dg_3d_parallel.jl:119-165) fused with the send: the local face state,
// aligned at the primary element, goes straight into the neighbour rank's receive buffer
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_mpi_pack_p4est(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / NF;
    const int fn = (int)(gid % NF);
    if (I >= P.nmpi) return;
    const int i = fn % N, j = fn / N;
    const long long element = P.mpi_local[I] - 1;
    int vn, sfn, dir;
    p4_face<ND, N>(P.mpi_node_indices + I * ND, i, j, vn, sfn, dir);
    const double *pu = P.u + (element * NN + vn) * NV;
    const int slot = P.mpi_peer_slot[I];
    double *dst = P.peer_recv[slot] + (P.mpi_remote_index[I] * NF + fn) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) dst[v] = pu[v];
    // shock capturing: the unsmoothed blending factor of the local element travels with its face (as on TreeMeshes)
    if (fn == 0 && P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG)
        P.peer_recv[slot][P.mpi_peer_nmpi[slot] * NF * NV + P.mpi_remote_index[I]] = P.alpha_raw[element];
}

// calc_mpi_interface_flux! (dgsem_p4est/dg_3d_parallel.jl:167-273): each rank uses the outward normal of its
// OWN element; the secondary side evaluates -f(u_ll, u_rr, -n) (so the two ranks' fluxes agree to rounding of
// the metric terms, not bit for bit -- as in the reference)
template <class EQ, int N>
__global__ void __launch_bounds__(256) k_mpi_interface_flux_p4est(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long I = gid / NF;
    const int fn = (int)(gid % NF);
    if (I >= P.nmpi) return;
    if (P.mpi_is_piece && P.mpi_is_piece[I]) return;  // exchange-only entry of an MPI mortar
    const int i = fn % N, j = fn / N;
    const EQ eq(P.eq);
    const long long element = P.mpi_local[I] - 1;
    const int side = (int)P.mpi_side[I];
    int vn, sfn, dir;
    p4_face<ND, N>(P.mpi_node_indices + I * ND, i, j, vn, sfn, dir);
    const double *pl = P.u + (element * NN + vn) * NV;
    const double *pr = P.recv + (I * NF + fn) * NV;
    double ul[NV], ur[NV], f[NV], nrm[ND];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const double a = pl[v], b = pr[v];
        ul[v] = side == 1 ? a : b;
        ur[v] = side == 1 ? b : a;
    }
    load_ja<ND, NN>(P, dir / 2, vn, element, nrm);
    // outward normal (dg.jl:74-86), negated once more on the secondary side
    if ((dir % 2 == 0) != (side == 2)) {
#pragma unroll
        for (int d = 0; d < ND; ++d) nrm[d] = -nrm[d];
    }
    eq.numflux_normal(P.surface_flux, ul, ur, nrm, f);
    double *s = P.sfv + ((element * (2 * ND) + dir) * NF + sfn) * NV;
    if constexpr (EQ::kHasNoncons) {
        if (EQ::has_noncons(P.surface_flux)) {  // dg_3d_parallel.jl:302-337
            double g[NV];
            if (side == 1)
                eq.noncons_normal(ul, ur, nrm, g);
            else
                eq.noncons_normal(ur, ul, nrm, g);
#pragma unroll
            for (int v = 0; v < NV; ++v) s[v] = side == 1 ? f[v] + 0.5 * g[v] : -f[v] + 0.5 * (-g[v]);
            return;
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = side == 1 ? f[v] : -f[v];
}

// calc_mpi_mortar_flux! + mpi_mortar_fluxes_to_elements! (dgsem_tree/dg_2d_parallel.jl:742-860, dgsem_p4est/
// dg_3d_parallel.jl:382-560) for both mesh kinds: k_mortar_flux / k_mortar_flux_p4est with the faces of remote
// elements read from the receive buffer of the halo exchange (already in the mortar's alignment) and results stored
// for local elements only.  A rank that does not own the large element evaluates just the positions of its own small
// elements.  P4est: the small elements' outward normals come precomputed (they may be remote).
template <class EQ, int N>
__global__ void __launch_bounds__((1 << (EQ::NDIMS - 1)) * ipow(N, EQ::NDIMS - 1)) k_mpi_mortar_flux(const KParams P) {
    constexpr int ND = EQ::NDIMS, NV = EQ::NVARS, NF = ipow(N, ND - 1), NN = ipow(N, ND), NP = 1 << (ND - 1);
    __shared__ double s_large[NV * NF];
    __shared__ double s_tmp[NP][NV * NF];
    __shared__ double s_f[NP][NV * NF];
    const long long m = blockIdx.x;
    const int tid = threadIdx.x;
    const int p = tid / NF, fn = tid - p * NF;
    const int a = fn % N, b = fn / N;
    const EQ eq(P.eq);
    const bool p4 = P.p4est != 0;
    const long long *ids = P.mpi_mortar_ids + (NP + 1) * m;
    const long long *sidx = p4 ? P.mpi_mortar_node_indices + (2 * m + 0) * ND : nullptr;
    const long long *lidx = p4 ? P.mpi_mortar_node_indices + (2 * m + 1) * ND : nullptr;
    const long long large_id = ids[NP], small_id = ids[p];
    const int o = p4 ? 0 : (int)P.mpi_mortar_orient[m] - 1;
    const bool large_left = !p4 && P.mpi_mortar_large_sides[m] == 1;
    const double *fwd1 = P.mortar_fwd[p & 1], *fwd2 = P.mortar_fwd[(p >> 1) & 1];
    const double *rev1 = P.mortar_rev[p & 1];
    for (int q = tid; q < NV * NF; q += NP * NF) {
        const int f = q / NV, v = q - f * NV;
        double val;
        if (large_id > 0) {
            int vn, sfn, dir;
            if (p4)
                p4_face<ND, N>(lidx, f % N, f / N, vn, sfn, dir);
            else
                vn = face_to_volume_node<ND, N>(o, large_left ? N - 1 : 0, f);
            val = P.u[((large_id - 1) * NN + vn) * NV + v];
        } else {
            val = P.recv[((-large_id - 1) * NF + f) * NV + v];
        }
        s_large[q] = val;
    }
    __syncthreads();
    double up[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double acc = 0.0;
        for (int q = 0; q < N; ++q) acc += fwd1[a + N * q] * s_large[v + NV * (q + N * b)];
        up[v] = acc;
    }
    if constexpr (ND == 3) {
#pragma unroll
        for (int v = 0; v < NV; ++v) s_tmp[p][v + NV * fn] = up[v];
        __syncthreads();
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double acc = 0.0;
            for (int q = 0; q < N; ++q) acc += fwd2[b + N * q] * s_tmp[p][v + NV * (a + N * q)];
            up[v] = acc;
        }
        __syncthreads();
    }
    double us[NV], f[NV], fs[NV];
    int sdir = large_left ? 2 * o : 2 * o + 1;
    if (small_id != 0) {
        if (small_id > 0) {
            int svn, ssfn;
            if (p4)
                p4_face<ND, N>(sidx, a, b, svn, ssfn, sdir);
            else
                svn = face_to_volume_node<ND, N>(o, large_left ? 0 : N - 1, fn);
            const double *pu = P.u + ((small_id - 1) * NN + svn) * NV;
#pragma unroll
            for (int v = 0; v < NV; ++v) us[v] = pu[v];
        } else {
            const double *pu = P.recv + ((-small_id - 1) * NF + fn) * NV;
#pragma unroll
            for (int v = 0; v < NV; ++v) us[v] = pu[v];
        }
        if (p4) {
            double nrm[ND];
            const double *pn = P.mpi_mortar_normals + ((m * NP + p) * NF + fn) * ND;
#pragma unroll
            for (int d = 0; d < ND; ++d) nrm[d] = pn[d];
            eq.numflux_normal(P.surface_flux, us, up, nrm, f);
#pragma unroll
            for (int v = 0; v < NV; ++v) fs[v] = f[v];
            if constexpr (EQ::kHasNoncons) {
                if (EQ::has_noncons(P.surface_flux)) {
                    double gp[NV], gs[NV];
                    eq.noncons_normal(us, up, nrm, gp);
                    eq.noncons_normal(up, us, nrm, gs);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        fs[v] = f[v] + 0.5 * gs[v];
                        f[v] = f[v] + 0.5 * gp[v];
                    }
                }
            }
        } else {
            mortar_point_flux<EQ>(eq, P.surface_flux, up, us, o, large_left, f, fs);
        }
        if (small_id > 0) {
            double *dst = P.sfv + (((small_id - 1) * (2 * ND) + sdir) * NF + fn) * NV;
#pragma unroll
            for (int v = 0; v < NV; ++v) dst[v] = f[v];
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) s_f[p][v + NV * fn] = fs[v];
    }
    if (large_id < 0) return;  // (uniform over the block) the large element belongs to another rank
    __syncthreads();
    int lsfn = fn, ldir = large_left ? 2 * o + 1 : 2 * o;
    if (p4) {
        int lvn;
        p4_face<ND, N>(lidx, a, b, lvn, lsfn, ldir);
    }
    double *out = P.sfv + (((large_id - 1) * (2 * ND) + ldir) * NF + lsfn) * NV;
    const double scale = p4 ? (ND == 2 ? -2.0 : -4.0) : 1.0;
    if constexpr (ND == 2) {
        if (tid < N) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double acc = 0.0;
                for (int q = 0; q < N; ++q)
                    acc += P.mortar_rev[1][tid + N * q] * s_f[1][v + NV * q] + P.mortar_rev[0][tid + N * q] * s_f[0][v + NV * q];
                out[v] = p4 ? acc * scale : acc;
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double acc = 0.0;
            for (int q = 0; q < N; ++q) acc += rev1[a + N * q] * s_f[p][v + NV * (q + N * b)];
            s_tmp[p][v + NV * fn] = acc;
        }
        __syncthreads();
        if (tid < NF) {
            // TreeMesh: upper_left, upper_right, lower_left, lower_right (dg_3d.jl:1314-1331); P4est: positions 1..4
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double res = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pp = p4 ? k : (k + 2) & 3;
                    const double *r2 = P.mortar_rev[(pp >> 1) & 1];
                    double acc = 0.0;
                    for (int q = 0; q < N; ++q) acc += r2[b + N * q] * s_tmp[pp][v + NV * (a + N * q)];
                    res = k == 0 ? acc : res + acc;
                }
                out[v] = p4 ? res * scale : res;
            }
        }
    }
}

}  // namespace tb
