// Line-sweep element kernel for any 3D equation system at polydeg 3: flux-differencing volume integral with an
// arbitrary registered two-point flux (and the nonconservative Powell term of GLM-MHD), fused with surface
// integral, Jacobian, source terms and the 2N Runge-Kutta stage.  Same skeleton as the headline kernel
// (kernel_euler3d_fd_p3.cuh) without the hoisted-logarithm trick:
//  * one warp = one CTA = one element, tiles by cp.async.bulk (u, surface_flux_values up front; u_tmp into the
//    line tile's storage once the z pass is done), results leave by bulk stores;
//  * per direction two threads share a line of 4 nodes and evaluate three of its six node pairs each, so
//    every two-point flux is computed once (the generic one-thread-per-node kernel computes each twice:
//    flux_differencing_kernel! dg_3d.jl:166-214 has the symmetric loop this restores);
//  * the line tile holds the conservative states at the swizzled node position of the headline kernel
//    (record strides 5 and 9 are odd), so the x, y and z sweeps are bank-conflict free.
// Used for Euler 3D with flux_shima_etal / kennedy_gruber / chandrashekar / central and for GLM-MHD with
// (flux_hindenlang_gassner, flux_nonconservative_powell): dg_3d.jl:216-266 for the nonconservative part.
#pragma once
#include "ranocha_common.cuh"
#include "tile_io.cuh"

namespace tb {

// Hoisted node records for the sweeps (what the reference's flux_ranocha_turbo / flux_shima_etal_turbo specializations
// do for Euler, dg_3d_compressible_euler.jl:289-309): an equation may keep primitive variables and logarithms per node in
// the line tile instead of the conservative state, so that the two-point flux of a pair needs neither cons2prim nor a
// logarithm.  Default: no record, the line tile holds the conservative variables.
template <class EQ>
struct NodeRecord {
    static constexpr bool kHas = false;
    static constexpr int N = EQ::NVARS;
};

// GLM-MHD: (rho, v1, v2, v3, p, B1, B2, B3, psi, log rho, log rho - log p).  flux_hindenlang_gassner
// (ideal_glm_mhd_3d.jl:680-779) needs ln_mean(rho_ll, rho_rr) and inv_ln_mean(rho_ll p_rr, rho_rr p_ll): with the two
// logarithms per NODE both come out of one division each (log(rho_rr p_ll / (rho_ll p_rr)) is the difference of the
// second logarithm entries); flux_nonconservative_powell (:295-340) reads velocities and fields directly.
template <>
struct NodeRecord<Mhd3D> {
    static constexpr bool kHas = true;
    static constexpr int N = 11;
    TB_DEV_HOST static bool applies(int volume_flux) {
        return volume_flux == TRIXI_B200_FLUX_HINDENLANG_GASSNER || volume_flux == TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL;
    }
    TB_DEV static void make(const Mhd3D &eq, int /*id*/, const double *u, double *r) {
        const double rho = u[0], inv_rho = fast_rcp(rho);
        const double v1 = u[1] * inv_rho, v2 = u[2] * inv_rho, v3 = u[3] * inv_rho;
        const double p = (eq.gamma - 1) * (u[4] - 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3 + u[5] * u[5] + u[6] * u[6] +
                                                         u[7] * u[7] + u[8] * u[8]));
        const double lrho = log_pos(rho);
        r[0] = rho, r[1] = v1, r[2] = v2, r[3] = v3, r[4] = p;
        r[5] = u[5], r[6] = u[6], r[7] = u[7], r[8] = u[8];
        r[9] = lrho, r[10] = lrho - log_pos(p);
    }
    // The sweeps hand over records whose velocity and field components have been rotated cyclically so that slot 1 / 5 is
    // the sweep direction (kRotated): the flux below is the orientation-1 form with compile-time component indices -- no
    // selects on the direction, no arm of a conditional evaluated in vain.  Its momentum and field components come out in
    // the same rotated order.  (The transverse sums then run in rotated order: last-bit differences to the reference.)
    static constexpr bool kRotated = true;
    // record of a node with the triples rotated for sweep direction d: slot k holds component (d + k) mod 3
    TB_DEV static void load(const double *src, int d, double *q) {
        const int c1 = d == 2 ? 0 : d + 1, c2 = d == 0 ? 2 : d - 1;
        q[0] = src[0], q[4] = src[4], q[8] = src[8], q[9] = src[9], q[10] = src[10];
        q[1] = src[1 + d], q[2] = src[1 + c1], q[3] = src[1 + c2];
        q[5] = src[5 + d], q[6] = src[5 + c1], q[7] = src[5 + c2];
    }
    // conservative component held by result slot v in direction d
    TB_DEV static void comps(int d, int (&comp)[9]) {
        const int c1 = d == 2 ? 0 : d + 1, c2 = d == 0 ? 2 : d - 1;
        comp[0] = 0, comp[4] = 4, comp[8] = 8;
        comp[1] = 1 + d, comp[2] = 1 + c1, comp[3] = 1 + c2;
        comp[5] = 5 + d, comp[6] = 5 + c1, comp[7] = 5 + c2;
    }
    // slot that holds component v after the z sweep (d = 2)
    TB_DEV static constexpr int slot_of(int v) {
        constexpr int t[9] = {0, 2, 3, 1, 4, 6, 7, 5, 8};
        return t[v];
    }
    TB_DEV static void flux(const Mhd3D &eq, int /*id*/, const double *L, const double *R, double (&f)[9]) {
        double rho_mean, inv_rho_p_mean;
        {
            const double sum = L[0] + R[0], dif = R[0] - L[0];
            const double q = dif * rcp_1nr(sum), f2 = q * q;
            const bool series = f2 < 1.0e-4;
            const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
            rho_mean = fast_div(series ? sum : dif, series ? poly : R[9] - L[9]);
        }
        {
            const double x = L[0] * R[4], y = R[0] * L[4];
            const double sum = x + y, dif = y - x;
            const double q = dif * rcp_1nr(sum), f2 = q * q;
            const bool series = f2 < 1.0e-4;
            const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
            inv_rho_p_mean = L[4] * R[4] * fast_div(series ? poly : R[10] - L[10], series ? sum : dif);
        }
        const double c_h = eq.c_h;
        const double vn_avg = 0.5 * (L[1] + R[1]), vt1_avg = 0.5 * (L[2] + R[2]), vt2_avg = 0.5 * (L[3] + R[3]);
        const double p_avg = 0.5 * (L[4] + R[4]), psi_avg = 0.5 * (L[8] + R[8]);
        const double velocity_square_avg = 0.5 * (L[1] * R[1] + L[2] * R[2] + L[3] * R[3]);
        const double magnetic_square_avg = 0.5 * (L[5] * R[5] + L[6] * R[6] + L[7] * R[7]);
        const double f1 = rho_mean * vn_avg;
        f[0] = f1;
        f[1] = f1 * vn_avg + p_avg + magnetic_square_avg - 0.5 * (L[5] * R[5] + R[5] * L[5]);
        f[2] = f1 * vt1_avg - 0.5 * (L[5] * R[6] + R[5] * L[6]);
        f[3] = f1 * vt2_avg - 0.5 * (L[5] * R[7] + R[5] * L[7]);
        f[5] = c_h * psi_avg;
        f[6] = 0.5 * (L[1] * L[6] - L[2] * L[5] + R[1] * R[6] - R[2] * R[5]);
        f[7] = 0.5 * (L[1] * L[7] - L[3] * L[5] + R[1] * R[7] - R[3] * R[5]);
        f[8] = c_h * 0.5 * (L[5] + R[5]);
        f[4] = f1 * (velocity_square_avg + inv_rho_p_mean * eq.inv_gm1) +
               0.5 * (+L[4] * R[1] + R[4] * L[1] + (L[1] * L[6] * R[6] + R[1] * R[6] * L[6]) +
                      (L[1] * L[7] * R[7] + R[1] * R[7] * L[7]) - (L[2] * L[5] * R[6] + R[2] * R[5] * L[6]) -
                      (L[3] * L[5] * R[7] + R[3] * R[5] * L[7]) + c_h * (L[5] * R[8] + R[5] * L[8]));
    }
    // flux_nonconservative_powell(u_ll, u_rr, orientation) on rotated records
    TB_DEV static void noncons(const double *L, const double *R, double (&g)[9]) {
        const double v_dot_B_ll = L[1] * L[5] + L[2] * L[6] + L[3] * L[7];
        const double Bn_rr = R[5], vo = L[1];
        g[0] = 0.0;
        g[1] = L[5] * Bn_rr, g[2] = L[6] * Bn_rr, g[3] = L[7] * Bn_rr;
        g[4] = v_dot_B_ll * Bn_rr + vo * L[8] * R[8];
        g[5] = L[1] * Bn_rr, g[6] = L[2] * Bn_rr, g[7] = L[3] * Bn_rr;
        g[8] = vo * R[8];
    }
};

// Compressible Euler with flux_ranocha (the volume flux of the shock-capturing elixirs): (rho, v1, v2, v3, p, log rho,
// log rho - log p) as in the reference's flux_ranocha_turbo specialization (dg_3d_compressible_euler.jl:289-309), same
// rotated-frame convention as above.
template <>
struct NodeRecord<Euler<3>> {
    static constexpr bool kHas = true;
    static constexpr int N = 7;
    static constexpr bool kRotated = true;
    TB_DEV_HOST static bool is_ranocha(int id) { return id == TRIXI_B200_FLUX_RANOCHA || id == TRIXI_B200_FLUX_RANOCHA_TURBO; }
    // flux_shima_etal and flux_kennedy_gruber use the primitive part only (the reference's second turbo
    // specialization is flux_shima_etal_turbo, dg_3d_compressible_euler.jl:8-263); Kennedy-Gruber keeps the specific
    // total energy in slot 5
    TB_DEV_HOST static bool applies(int volume_flux) {
        return is_ranocha(volume_flux) || volume_flux == TRIXI_B200_FLUX_SHIMA_ETAL ||
               volume_flux == TRIXI_B200_FLUX_KENNEDY_GRUBER || volume_flux == TRIXI_B200_FLUX_CHANDRASHEKAR;
    }
    TB_DEV static void make(const Euler<3> &eq, int id, const double *u, double *r) {
        const double rho = u[0], inv_rho = fast_rcp(rho);
        double v1 = u[1] * inv_rho, v2 = u[2] * inv_rho, v3 = u[3] * inv_rho;
        v1 = fma(fma(-rho, v1, u[1]), inv_rho, v1);
        v2 = fma(fma(-rho, v2, u[2]), inv_rho, v2);
        v3 = fma(fma(-rho, v3, u[3]), inv_rho, v3);
        const double p = (eq.gamma - 1) * (u[4] - 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3));
        r[0] = rho, r[1] = v1, r[2] = v2, r[3] = v3, r[4] = p;
        if (is_ranocha(id)) {
            const double lrho = log_pos(rho);
            r[5] = lrho, r[6] = lrho - log_pos(p);
        } else if (id == TRIXI_B200_FLUX_CHANDRASHEKAR) {
            // flux_chandrashekar works with beta = rho / (2 p): slot 4 holds beta, slot 6 log(rho / p) (= log beta + log 2)
            const double lrho = log_pos(rho);
            r[4] = 0.5 * fast_div(rho, p), r[5] = lrho, r[6] = lrho - log_pos(p);
        } else {
            const double e = u[4] * inv_rho;
            r[5] = fma(fma(-rho, e, u[4]), inv_rho, e), r[6] = 0.0;
        }
    }
    TB_DEV static void load(const double *src, int d, double *q) {
        const int c1 = d == 2 ? 0 : d + 1, c2 = d == 0 ? 2 : d - 1;
        q[0] = src[0], q[4] = src[4], q[5] = src[5], q[6] = src[6];
        q[1] = src[1 + d], q[2] = src[1 + c1], q[3] = src[1 + c2];
    }
    TB_DEV static void comps(int d, int (&comp)[5]) {
        const int c1 = d == 2 ? 0 : d + 1, c2 = d == 0 ? 2 : d - 1;
        comp[0] = 0, comp[4] = 4;
        comp[1] = 1 + d, comp[2] = 1 + c1, comp[3] = 1 + c2;
    }
    TB_DEV static constexpr int slot_of(int v) {
        constexpr int t[5] = {0, 2, 3, 1, 4};
        return t[v];
    }
    // flux_ranocha (compressible_euler_3d.jl:746-793) in the rotated frame: slot 1 is the normal velocity
    TB_DEV static void flux(const Euler<3> &eq, int id, const double *L, const double *R, double (&f)[5]) {
        if (id == TRIXI_B200_FLUX_CHANDRASHEKAR) {
            // flux_chandrashekar (compressible_euler_3d.jl:639-690), rotated frame; both logarithmic means from the
            // hoisted logarithms
            double rho_mean, inv_beta_mean;
            {
                const double sum = L[0] + R[0], dif = R[0] - L[0];
                const double q = dif * rcp_1nr(sum), f2 = q * q;
                const bool series = f2 < 1.0e-4;
                const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
                rho_mean = fast_div(series ? sum : dif, series ? poly : R[5] - L[5]);
            }
            {
                const double sum = L[4] + R[4], dif = R[4] - L[4];
                const double q = dif * rcp_1nr(sum), f2 = q * q;
                const bool series = f2 < 1.0e-4;
                const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
                inv_beta_mean = fast_div(series ? poly : R[6] - L[6], series ? sum : dif);
            }
            const double rho_avg = 0.5 * (L[0] + R[0]), beta_avg = 0.5 * (L[4] + R[4]);
            const double vn_avg = 0.5 * (L[1] + R[1]), vt1_avg = 0.5 * (L[2] + R[2]), vt2_avg = 0.5 * (L[3] + R[3]);
            const double p_mean = 0.5 * fast_div(rho_avg, beta_avg);
            const double velocity_square_avg = 0.5 * (L[1] * L[1] + L[2] * L[2] + L[3] * L[3]) +
                                               0.5 * (R[1] * R[1] + R[2] * R[2] + R[3] * R[3]);
            const double f1 = rho_mean * vn_avg;
            f[0] = f1;
            f[1] = f1 * vn_avg + p_mean;
            f[2] = f1 * vt1_avg;
            f[3] = f1 * vt2_avg;
            f[4] = f1 * 0.5 * (eq.inv_gm1 * inv_beta_mean - velocity_square_avg) + f[1] * vn_avg + f[2] * vt1_avg + f[3] * vt2_avg;
            return;
        }
        if (!is_ranocha(id)) {
            // flux_shima_etal (compressible_euler_3d.jl:473-510) / flux_kennedy_gruber (:560-600), rotated frame
            const double rho_avg = 0.5 * (L[0] + R[0]), p_avg = 0.5 * (L[4] + R[4]);
            const double vn_avg = 0.5 * (L[1] + R[1]), vt1_avg = 0.5 * (L[2] + R[2]), vt2_avg = 0.5 * (L[3] + R[3]);
            const double f1 = rho_avg * vn_avg;
            f[0] = f1;
            f[1] = f1 * vn_avg + p_avg;
            f[2] = f1 * vt1_avg;
            f[3] = f1 * vt2_avg;
            if (id == TRIXI_B200_FLUX_SHIMA_ETAL) {
                const double kin_avg = 0.5 * (L[1] * R[1] + L[2] * R[2] + L[3] * R[3]);
                const double pv_avg = 0.5 * (L[4] * R[1] + R[4] * L[1]);
                f[4] = p_avg * vn_avg * eq.inv_gm1 + f1 * kin_avg + pv_avg;
            } else {
                const double e_avg = 0.5 * (L[5] + R[5]);
                f[4] = (rho_avg * e_avg + p_avg) * vn_avg;
            }
            return;
        }
        double rho_mean, inv_rho_p_mean;
        {
            const double sum = L[0] + R[0], dif = R[0] - L[0];
            const double q = dif * rcp_1nr(sum), f2 = q * q;
            const bool series = f2 < 1.0e-4;
            const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
            rho_mean = fast_div(series ? sum : dif, series ? poly : R[5] - L[5]);
        }
        {
            const double x = L[0] * R[4], y = R[0] * L[4];
            const double sum = x + y, dif = y - x;
            const double q = dif * rcp_1nr(sum), f2 = q * q;
            const bool series = f2 < 1.0e-4;
            const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
            inv_rho_p_mean = L[4] * R[4] * fast_div(series ? poly : R[6] - L[6], series ? sum : dif);
        }
        const double vn_avg = 0.5 * (L[1] + R[1]), vt1_avg = 0.5 * (L[2] + R[2]), vt2_avg = 0.5 * (L[3] + R[3]);
        const double p_avg = 0.5 * (L[4] + R[4]);
        const double velocity_square_avg = 0.5 * (L[1] * R[1] + L[2] * R[2] + L[3] * R[3]);
        const double f1 = rho_mean * vn_avg;
        f[0] = f1;
        f[1] = f1 * vn_avg + p_avg;
        f[2] = f1 * vt1_avg;
        f[3] = f1 * vt2_avg;
        f[4] = f1 * (velocity_square_avg + inv_rho_p_mean * eq.inv_gm1) + 0.5 * (L[4] * R[1] + R[4] * L[1]);
    }
    TB_DEV static void noncons(const double *, const double *, double (&)[5]) {}
};

// run-time two-point flux of the line sweeps: the compressible Euler equations take their core switch (see
// Euler::numflux_core; launch.cuh sends set-ups with flux_hllc / flux_hlle as volume fluxes to the generic kernels)
template <class EQ, int NVV>
TB_DEV void numflux_tuned(const EQ &eq, int id, const double (&ul)[NVV], const double (&ur)[NVV], int o, double (&f)[NVV]) {
    if constexpr (HasFastRanocha<EQ>::value)
        eq.numflux_core(id, ul, ur, o, f);
    else
        eq.numflux(id, ul, ur, o, f);
}

template <class EQ>
struct LineSweepCfg {
    static constexpr int NV = EQ::NVARS, THREADS = 32;
    static constexpr int CONS = 64 * NV, SFV = 96 * NV;  // doubles per element
    static constexpr int LINE = 64 * NodeRecord<EQ>::N;  // the line tile: states or hoisted node records
    // s_u (natural, TMA), s_sfv (natural, TMA), s_du (swizzled), s_line (swizzled; later the u_tmp tile), mbarrier
    static constexpr size_t SMEM = sizeof(double) * (2 * CONS + SFV + LINE) + 16;
    static constexpr int BLOCKS_PER_SM = (int)((227 * 1024 + 1024) / (SMEM + 1024));  // 1 KB per CTA is reserved
    static constexpr int MIN_BLOCKS = BLOCKS_PER_SM < 8 ? BLOCKS_PER_SM : (BLOCKS_PER_SM > 16 ? 16 : BLOCKS_PER_SM);
};

// SC: VolumeIntegralShockCapturingHG (calc_volume_integral.jl:231-272).  The blending factor is per element =
// per warp, so pure-DG elements skip the finite-volume part without divergence.  Blended elements scale the
// D_split weights by 1 - alpha and add alpha w_i^-1 (f*_{i+1/2} - f*_{i-1/2}) per direction (fv_kernel!
// dg_3d.jl:268-306) for their own two nodes: thread h = 0 needs the subcell fluxes (0,1) and (1,2) of its line,
// thread h = 1 needs (2,3) and (1,2); the first is the operand pair of its first two-point flux, (1,2) is
// evaluated by both threads with identical operands (bitwise equal, so the subcell scheme stays conservative).
// GEN: also the 3S* / SSP stage updates (KParams::mode 2, 3); see kernel_euler3d_fd_p3.cuh
template <class EQ, bool WITH_SURFACE, bool SC = false, bool REC = false, bool GEN = false>
__global__ void __launch_bounds__(LineSweepCfg<EQ>::THREADS, LineSweepCfg<EQ>::MIN_BLOCKS)
    k_element_fd3d_p3(const KParams P) {
    using C = LineSweepCfg<EQ>;
    using Rec = NodeRecord<EQ>;
    static_assert(!REC || Rec::kHas, "node records need equation support");
    constexpr int NV = C::NV, CONS = C::CONS, SFV = C::SFV;
    constexpr int NR = REC ? Rec::N : NV;  // doubles per node in the line tile
    extern __shared__ __align__(128) double smem[];
    double *s_u = smem;             // [64][NV] natural: u in, updated u out
    double *s_sfv = s_u + CONS;     // [6][16][NV] natural
    double *s_du = s_sfv + SFV;     // [64][NV] swizzled
    double *s_line = s_du + CONS;   // [64][NV] swizzled copy of u for the sweeps
    double *s_ut = s_line;          // afterwards: [64][NV] natural, u_tmp in, u_tmp (or du) out
    const uint32_t bar = smem_u32(s_line + C::LINE);

    const EQ eq(P.eq);
    const int lane = threadIdx.x;
    const long long e = P.elem_begin + blockIdx.x;
    const bool rk = P.mode != 0;
    const bool need_ut = rk && P.rk_read_tmp;

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    constexpr uint32_t bu = CONS * sizeof(double), bs = SFV * sizeof(double);
    if (lane == 0) {
        mbar_expect_tx(bar, bu + (WITH_SURFACE ? bs : 0u));
        tma_load(smem_u32(s_u), P.u + e * CONS, bu, bar);
        if (WITH_SURFACE) tma_load(smem_u32(s_sfv), P.sfv + e * SFV, bs, bar);
        const long long en = e + P.prefetch_distance;
        if (P.prefetch_distance > 0 && en < P.nelements) {
            tma_prefetch_l2(P.u + en * CONS, bu);
            if (need_ut) tma_prefetch_l2(P.u_tmp + en * CONS, bu);
            if (WITH_SURFACE) tma_prefetch_l2(P.sfv + en * SFV, bs);
        }
    }
    // line-local roles as in the headline kernel: thread h of line l16 owns nodes lm[0], lm[1], sees lm[2], lm[3]
    const int h = lane >> 4, l16 = lane & 15;
    const int a0 = l16 & 3, a1 = l16 >> 2;
    const int lm[4] = {h ? 3 : 0, h ? 2 : 1, h ? 0 : 2, h ? 1 : 3};
    double w01 = P.dsplit_c[lm[0] + 4 * lm[1]], w10 = P.dsplit_c[lm[1] + 4 * lm[0]];
    double w02 = P.dsplit_c[lm[0] + 4 * lm[2]], w20 = P.dsplit_c[lm[2] + 4 * lm[0]];
    double w13 = P.dsplit_c[lm[1] + 4 * lm[3]], w31 = P.dsplit_c[lm[3] + 4 * lm[1]];
    double fv0 = 0.0, fv1 = 0.0;  // alpha * (+-) inverse_weights of the two own nodes; 0: pure DG element
    if constexpr (SC) {
        const double alpha = P.alpha[e];
        if (!(fabs(alpha) <= 1.8189894035458565e-12)) {  // isapprox(alpha, 0, atol = max(100 eps, eps^0.75))
            const double w_dg = 1 - alpha;
            w01 *= w_dg, w10 *= w_dg, w02 *= w_dg, w20 *= w_dg, w13 *= w_dg, w31 *= w_dg;
            const double sgn = h ? -alpha : alpha;
            fv0 = sgn * P.inv_weights_c[0];  // nodes 0 and 3 (inverse_weights is symmetric)
            fv1 = sgn * P.inv_weights_c[1];  // nodes 1 and 2
        }
    }
    while (!mbar_try_wait(bar, 0)) {
    }

    // 1. swizzled copy of the element's states
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = lane + 32 * r;
        const double *c = s_u + n * NV;
        double *o = s_line + swz_pos(n) * NR;
        if constexpr (REC) {
            Rec::make(eq, P.volume_flux, c, o);
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v) o[v] = c[v];
        }
    }
    __syncwarp();

    // 2. direction sweeps
    bool noncons = false;
    if constexpr (EQ::kHasNoncons) noncons = EQ::has_noncons(P.volume_flux);
    (void)noncons;
    int pos[4];
    double own[2][NV], frn[2][NV];
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        const int stride = 1 << (2 * d);
        const int base = d == 0 ? 4 * l16 : (d == 1 ? a0 + 16 * a1 : l16);
        double q[4][NR];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            pos[m] = swz_pos(base + lm[m] * stride);
            const double *src = s_line + pos[m] * NR;
            if constexpr (REC) {
                Rec::load(src, d, q[m]);  // cyclic rotation: slot k of a triple holds component (d + k) mod 3
            } else {
#pragma unroll
                for (int v = 0; v < NR; ++v) q[m][v] = src[v];
            }
        }
        // (records: the triples of own/frn are in the rotated order of this direction; component of slot v)
        int comp[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) comp[v] = v;
        if constexpr (REC) Rec::comps(d, comp);
        double f[NV], lo[NR], hi[NR];
        auto two_point = [&](const double(&a)[NR], const double(&b)[NR], double(&out)[NV]) {
            if constexpr (REC)
                Rec::flux(eq, P.volume_flux, a, b, out);
            else
                numflux_tuned(eq, P.volume_flux, a, b, d, out);
        };
        auto noncons_term = [&](const double(&a)[NR], const double(&b)[NR], double(&out)[NV]) {
            if constexpr (REC)
                Rec::noncons(a, b, out);
            else if constexpr (EQ::kHasNoncons)
                eq.noncons(a, b, d, out);
        };
        // volume_flux(u_lower, u_upper) like the reference's loop (ii > i): thread h = 1 walks its line downwards,
        // so its operands are swapped -- by value selects, the two half-warps must not diverge around the flux
        if constexpr (REC) {
            // (record fluxes are symmetric up to rounding: no operand swap for the downward half-warp)
            two_point(q[0], q[1], f);
        } else {
#pragma unroll
            for (int v = 0; v < NR; ++v) {
                lo[v] = h ? q[1][v] : q[0][v];
                hi[v] = h ? q[0][v] : q[1][v];
            }
            two_point(lo, hi, f);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            own[0][v] = w01 * f[v];
            own[1][v] = w10 * f[v];
        }
        if constexpr (REC) {
            // (record fluxes are symmetric up to rounding: no operand swap for the downward half-warp)
            two_point(q[0], q[2], f);
        } else {
#pragma unroll
            for (int v = 0; v < NR; ++v) {
                lo[v] = h ? q[2][v] : q[0][v];
                hi[v] = h ? q[0][v] : q[2][v];
            }
            two_point(lo, hi, f);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            own[0][v] = fma(w02, f[v], own[0][v]);
            frn[0][v] = w20 * f[v];
        }
        if constexpr (REC) {
            // (record fluxes are symmetric up to rounding: no operand swap for the downward half-warp)
            two_point(q[1], q[3], f);
        } else {
#pragma unroll
            for (int v = 0; v < NR; ++v) {
                lo[v] = h ? q[3][v] : q[1][v];
                hi[v] = h ? q[1][v] : q[3][v];
            }
            two_point(lo, hi, f);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            own[1][v] = fma(w13, f[v], own[1][v]);
            frn[1][v] = w31 * f[v];
        }
        if constexpr (EQ::kHasNoncons) {
            // nonconservative volume terms (dg_3d.jl:216-266): node a gets 0.5 D_split[a, b] g(u_a, u_b) from
            // every partner b of its line (D_split has a zero diagonal)
            if (noncons) {
                double g[NV];
                noncons_term(q[0], q[1], g);
#pragma unroll
                for (int v = 0; v < NV; ++v) own[0][v] = fma(0.5 * w01, g[v], own[0][v]);
                noncons_term(q[1], q[0], g);
#pragma unroll
                for (int v = 0; v < NV; ++v) own[1][v] = fma(0.5 * w10, g[v], own[1][v]);
                noncons_term(q[0], q[2], g);
#pragma unroll
                for (int v = 0; v < NV; ++v) own[0][v] = fma(0.5 * w02, g[v], own[0][v]);
                noncons_term(q[2], q[0], g);
#pragma unroll
                for (int v = 0; v < NV; ++v) frn[0][v] = fma(0.5 * w20, g[v], frn[0][v]);
                noncons_term(q[1], q[3], g);
#pragma unroll
                for (int v = 0; v < NV; ++v) own[1][v] = fma(0.5 * w13, g[v], own[1][v]);
                noncons_term(q[3], q[1], g);
#pragma unroll
                for (int v = 0; v < NV; ++v) frn[1][v] = fma(0.5 * w31, g[v], frn[1][v]);
            }
        }
        if (d < 2) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                double *t = s_du + pos[m] * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) t[comp[v]] = d == 0 ? own[m][v] : t[comp[v]] + own[m][v];
            }
            __syncwarp();
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double *t = s_du + pos[2 + m] * NV;
#pragma unroll
            for (int v = 0; v < NV; ++v) t[comp[v]] += frn[m][v];
        }
        __syncwarp();
    }

    // VolumeIntegralShockCapturingHG, blended elements only (the branch is uniform over the warp): the subcell
    // finite-volume part alpha w_i^-1 (f*_{i+1/2} - f*_{i-1/2}) (fv_kernel! dg_3d.jl:268-306) as its own pass over the
    // three directions, after the sweeps -- inside them its registers cost every pure-DG element a quarter of its
    // speed.  Thread h = 0 of a line needs the subcell fluxes (0,1) and (1,2), thread h = 1 needs (2,3) and (1,2);
    // (1,2) is evaluated by both with identical operands (bitwise equal, so the subcell scheme stays conservative).
    // The conservative states come from the natural-order u tile, the results go to the du tile.
    if constexpr (SC) {
        if (fv0 != 0.0) {
#pragma unroll 1
            for (int d = 0; d < 3; ++d) {
                const int stride = 1 << (2 * d);
                const int base = d == 0 ? 4 * l16 : (d == 1 ? a0 + 16 * a1 : l16);
                const double *c0 = s_u + (base + lm[0] * stride) * NV, *c1 = s_u + (base + lm[1] * stride) * NV;
                const double *c2 = s_u + (base + lm[2] * stride) * NV, *c3 = s_u + (base + lm[3] * stride) * NV;
                double ca[NV], cb[NV], fa[NV], fb[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    ca[v] = h ? c1[v] : c0[v];  // the pair (0,1) [h = 0] or (2,3) [h = 1], lower node first
                    cb[v] = h ? c0[v] : c1[v];
                }
                numflux_tuned(eq, P.volume_flux_fv, ca, cb, d, fa);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    ca[v] = h ? c3[v] : c1[v];  // the pair (1,2)
                    cb[v] = h ? c1[v] : c2[v];
                }
                numflux_tuned(eq, P.volume_flux_fv, ca, cb, d, fb);
                double *t0 = s_du + swz_pos(base + lm[0] * stride) * NV, *t1 = s_du + swz_pos(base + lm[1] * stride) * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    t0[v] = fma(fv0, fa[v], t0[v]);          // node 0: +f(0,1); node 3: -f(2,3)
                    t1[v] = fma(fv1, fb[v] - fa[v], t1[v]);  // node 1: f(1,2) - f(0,1); node 2: f(2,3) - f(1,2)
                }
                __syncwarp();
            }
        }
    }

    // the line tile is dead: fetch u_tmp into its storage while the surface terms are applied
    if (need_ut) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(bar, bu);
            tma_load(smem_u32(s_ut), P.u_tmp + e * CONS, bu, bar);
        }
    }

    // 3. finish the two own nodes (i, j, k) = (a0, a1, lm[0]), (a0, a1, lm[1]) of the z line
    const int i = a0, j = a1;
    const double factor = WITH_SURFACE ? -P.inverse_jacobian[e] : 1.0;
    const bool have_src = WITH_SURFACE && P.source_terms != TRIXI_B200_SRC_NONE;
    double vals[2][NV];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int k = lm[r];
        const int n = l16 + 16 * k;
        const double *t = s_du + pos[r] * NV;
        double(&val)[NV] = vals[r];
        if constexpr (REC) {
            // the z sweep left own[] in its rotated order: slot k of a triple holds component (2 + k) mod 3
#pragma unroll
            for (int v = 0; v < NV; ++v) val[v] = t[v] + own[r][Rec::slot_of(v)];
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v) val[v] = t[v] + own[r][v];
        }
        if constexpr (WITH_SURFACE) {
            // calc_surface_integral! (dg_3d.jl:1337-1394): directions 1..6 = -x,+x,-y,+y,-z,+z
            if (i == 0 || i == 3) {
                const double *sf = s_sfv + ((i == 0 ? 0 : 1) * 16 + j + 4 * k) * NV;
                const double w = i == 0 ? -P.inv_weight0 : P.inv_weight0;
#pragma unroll
                for (int v = 0; v < NV; ++v) val[v] = fma(sf[v], w, val[v]);
            }
            if (j == 0 || j == 3) {
                const double *sf = s_sfv + ((j == 0 ? 2 : 3) * 16 + i + 4 * k) * NV;
                const double w = j == 0 ? -P.inv_weight0 : P.inv_weight0;
#pragma unroll
                for (int v = 0; v < NV; ++v) val[v] = fma(sf[v], w, val[v]);
            }
            if (k == 0 || k == 3) {
                const double *sf = s_sfv + ((k == 0 ? 4 : 5) * 16 + l16) * NV;
                const double w = k == 0 ? -P.inv_weight0 : P.inv_weight0;
#pragma unroll
                for (int v = 0; v < NV; ++v) val[v] = fma(sf[v], w, val[v]);
            }
            // apply_jacobian! (dg_3d.jl:1396-1414)
#pragma unroll
            for (int v = 0; v < NV; ++v) val[v] *= factor;
            // calc_sources! (dg_3d.jl:1417-1437)
            if (have_src) {
                double un[NV], x[3], sv[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) un[v] = s_u[n * NV + v];
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) x[dd] = P.node_coordinates[(e * 64 + n) * 3 + dd];
                eq.source_terms(P.source_terms, un, x, P.t, sv);
#pragma unroll
                for (int v = 0; v < NV; ++v) val[v] += sv[v];
            }
        }
    }
    if (!need_ut) __syncwarp();  // every lane is done with the line tile before it becomes the output tile
    if (need_ut) {
        while (!mbar_try_wait(bar, 1)) {
        }
    }
    unsigned long long cfl[3] = {0ull, 0ull, 0ull};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = l16 + 16 * lm[r];
        double(&val)[NV] = vals[r];
        double *out_t = s_ut + n * NV;
        if (!rk) {
#pragma unroll
            for (int v = 0; v < NV; ++v) out_t[v] = val[v];
        } else {
            // 2N stage (methods_2N.jl:152-158): u_tmp = du - u_tmp * a; u += u_tmp * (b * dt)
            double *out_u = s_u + n * NV;
            double un[NV];
            if (!GEN || P.mode == 1) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double tmp = need_ut ? val[v] - out_t[v] * P.rk_a : val[v];
                    out_t[v] = tmp;
                    un[v] = out_u[v] + tmp * P.rk_b_dt;
                    out_u[v] = un[v];
                }
            } else {
                // 3S* / SSP stage (KParams::mode 2, 3): u_tmp2 comes straight from global memory
                const double *u2 = P.u_tmp2 + (e * 64 + n) * NV;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    double xn;
                    un[v] = rk_stage_3s_ssp(P, val[v], need_ut ? out_t[v] : 0.0, out_u[v], P.mode == 2 ? u2[v] : 0.0, xn);
                    out_t[v] = xn;
                    out_u[v] = un[v];
                }
            }
            if (P.want_cfl) {
                double lam[3];
                eq.max_abs_speeds(un, lam);
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) cfl[dd] = max(cfl[dd], cfl_encode(lam[dd]));
            }
        }
    }
    if (rk && P.want_cfl) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) cfl[dd] = max(cfl[dd], __shfl_xor_sync(0xffffffffu, cfl[dd], off));
        }
        if (lane == 0) {
            double sum = 0.0;
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) sum += __longlong_as_double((long long)cfl[dd]);
            atomicMax(P.cfl_key + (blockIdx.x & (kCflSlots - 1)), cfl_encode(P.inverse_jacobian[e] * sum));
        }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (!rk) {
            tma_store(P.du + e * CONS, smem_u32(s_ut), bu);
        } else {
            if (!GEN || P.rk_write_tmp) tma_store(P.u_tmp + e * CONS, smem_u32(s_ut), bu);
            tma_store(P.u_out + e * CONS, smem_u32(s_u), bu);
        }
        tma_store_commit_and_wait_read();
    }
}

template <class EQ, bool SC = false>
cudaError_t preload_fd3d_p3() {
    cudaError_t e = preload_kernel(k_element_fd3d_p3<EQ, true, SC>);
    if (e != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_fd3d_p3<EQ, true, SC, false, true>)) != cudaSuccess) return e;
    if constexpr (NodeRecord<EQ>::kHas) {
        if ((e = preload_kernel(k_element_fd3d_p3<EQ, true, SC, true>)) != cudaSuccess) return e;
        if ((e = preload_kernel(k_element_fd3d_p3<EQ, true, SC, true, true>)) != cudaSuccess) return e;
        if ((e = preload_kernel(k_element_fd3d_p3<EQ, false, SC, true>)) != cudaSuccess) return e;
    }
    return preload_kernel(k_element_fd3d_p3<EQ, false, SC>);
}

template <class EQ, bool WS, bool SC = false, bool REC = false, bool GEN = false>
cudaError_t launch_fd3d_p3_variant(const KParams &P, cudaStream_t s) {
    using C = LineSweepCfg<EQ>;
    static PerDeviceFlag configured;
    auto kern = k_element_fd3d_p3<EQ, WS, SC, REC, GEN>;
    if (!configured.test_and_set()) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                               cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        if (C::SMEM > 48 * 1024) {
            err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            if (err != cudaSuccess) return err;
        }
    }
    KParams Q = P;
    if (Q.prefetch_distance < 0) Q.prefetch_distance = C::BLOCKS_PER_SM * Q.sm_count;
    kern<<<(unsigned)(P.elem_end - P.elem_begin), C::THREADS, C::SMEM, s>>>(Q);
    return cudaSuccess;
}

template <class EQ, bool SC = false>
cudaError_t launch_element_fd3d_p3(const KParams &P, bool with_surface, cudaStream_t s) {
    if constexpr (NodeRecord<EQ>::kHas) {
        // hoisted node records where the volume flux has a record form (kernel_path 2 = the plain form, for A/B runs)
        if (NodeRecord<EQ>::applies(P.volume_flux) && P.kernel_path != 2) {
            if (P.mode > 1) return launch_fd3d_p3_variant<EQ, true, SC, true, true>(P, s);  // 3S* / SSP stage
            return with_surface ? launch_fd3d_p3_variant<EQ, true, SC, true>(P, s)
                                : launch_fd3d_p3_variant<EQ, false, SC, true>(P, s);
        }
    }
    if (P.mode > 1) return launch_fd3d_p3_variant<EQ, true, SC, false, true>(P, s);  // 3S* / SSP stage
    return with_surface ? launch_fd3d_p3_variant<EQ, true, SC>(P, s) : launch_fd3d_p3_variant<EQ, false, SC>(P, s);
}

}  // namespace tb
