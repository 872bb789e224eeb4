// Pointwise physics as __device__ functions: the reference's L1 layer
// (src/equations/*.jl, src/auxiliary/math.jl) re-expressed for sm_100a.  Orientation is a plain int
// (0-based) consumed through selects only -- no dynamically indexed register arrays -- so the same
// code folds to straight-line FP64 when the caller unrolls over directions (volume kernels) and
// costs a few predicated moves when the orientation is per-face data (surface kernels).
// nvcc's default -fmad=true contracts a*b+c into DFMA, mirroring the reference's @muladd.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/trixi_b200.h"

namespace tb {

struct EqParams {
    double p[8];
};

#define TB_DEV __device__ __forceinline__
#define TB_DEV_HOST __host__ __device__ __forceinline__

template <int ND>
TB_DEV double pick(const double (&v)[ND], int o) {
    if constexpr (ND == 1) return v[0];
    if constexpr (ND == 2) return o == 0 ? v[0] : v[1];
    if constexpr (ND == 3) return o == 0 ? v[0] : (o == 1 ? v[1] : v[2]);
}

// ---- src/auxiliary/math.jl ------------------------------------------------------------------------
// ln_mean (math.jl:198-210): Ismail-Roe series for f^2 < 1e-4, else (y-x)/log(y/x)
TB_DEV double ln_mean(double x, double y) {
    const double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
    if (f2 < 1.0e-4) {
        const double p = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        return (x + y) / p;
    }
    return (y - x) / log(y / x);
}
// inv_ln_mean (math.jl:238-250)
TB_DEV double inv_ln_mean(double x, double y) {
    const double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
    if (f2 < 1.0e-4) {
        const double p = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        return p / (x + y);
    }
    return log(y / x) / (y - x);
}

// ---- fast FP64 reciprocal / division (no IEEE slow path, ~1 ulp) -------------------------------------
// 1/x: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps (4 DFMA)
TB_DEV double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
// a / b: seed, ONE Newton step (relative error e1 ~ 2^-40) and a residual correction of the quotient (Markstein):
// q' = q + r (a - b q) has relative error e1^2 before its final rounding, so the second Newton step of fast_rcp
// would buy nothing here.  Correctly rounded except in rare ties.
TB_DEV double fast_div(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
// sqrt(x) for positive normal x: MUFU.RSQ64H seed, two coupled Newton steps on (g, h) = (sqrt x, 1 / (2 sqrt x)) and a
// final residual correction (within 1 ulp; no IEEE slow path, no subnormal/negative handling: NaN stays NaN)
TB_DEV double fast_sqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
}
// 1/x to ~1e-12: seed + one Newton step; enough for f^2 of the logarithmic means, which only enters the
// Ismail-Roe series (sensitivity f^2/3 <= 3e-5) and the branch choice
TB_DEV double rcp_1nr(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
}
// ln_mean / inv_ln_mean (math.jl:198-250) with fast divisions; the log is only evaluated where the
// series does not apply
TB_DEV double ln_mean_fast(double x, double y) {
    const double sum = x + y, dif = y - x;
    const double f2 = (dif * dif) * rcp_1nr(sum * sum);
    if (f2 < 1.0e-4) return fast_div(sum, fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0));
    return fast_div(dif, log(fast_div(y, x)));
}
TB_DEV double inv_ln_mean_fast(double x, double y) {
    const double sum = x + y, dif = y - x;
    const double f2 = (dif * dif) * rcp_1nr(sum * sum);
    if (f2 < 1.0e-4) return fast_div(fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0), sum);
    return fast_div(log(fast_div(y, x)), dif);
}

// ---- linear scalar advection (linear_scalar_advection_2d.jl / _3d.jl) -------------------------------
template <int ND>
struct Advection {
    static constexpr int NDIMS = ND, NVARS = 1;
    static constexpr bool kConstantSpeed = true;  // have_constant_speed (linear_scalar_advection_2d.jl:290)
    static constexpr bool kHasNoncons = false;
    double a0, a1, a2;
    __host__ __device__ explicit Advection(const EqParams &q) : a0(q.p[0]), a1(q.p[1]), a2(q.p[2]) {}
    TB_DEV double a(int o) const { return o == 0 ? a0 : (o == 1 ? a1 : a2); }

    // flux (linear_scalar_advection_2d.jl:221-225)
    TB_DEV void flux(const double (&u)[1], int o, double (&f)[1]) const { f[0] = a(o) * u[0]; }

    TB_DEV void numflux(int id, const double (&ul)[1], const double (&ur)[1], int o, double (&f)[1]) const {
        const double an = a(o);
        switch (id) {
        case TRIXI_B200_FLUX_CENTRAL:  // numerical_fluxes.jl:17-25
            f[0] = 0.5 * (an * ul[0] + an * ur[0]);
            break;
        case TRIXI_B200_FLUX_LLF:  // max_abs_speed falls back to the naive one (numerical_fluxes.jl:219-225)
        case TRIXI_B200_FLUX_LLF_NAIVE: {  // linear_scalar_advection_2d.jl:228-231
            const double lam = fabs(an);
            f[0] = 0.5 * (an * ul[0] + an * ur[0]) + (-0.5 * lam * (ur[0] - ul[0]));
            break;
        }
        case TRIXI_B200_FLUX_GODUNOV:  // linear_scalar_advection_2d.jl:248-260
            f[0] = an >= 0 ? an * ul[0] : an * ur[0];
            break;
        default:
            f[0] = nan("");
        }
    }
    TB_DEV double a_dot(const double (&n)[ND]) const {
        double s = a0 * n[0];
        if constexpr (ND > 1) s += a1 * n[1];
        if constexpr (ND > 2) s += a2 * n[2];
        return s;
    }
    // flux(u, normal_direction) (linear_scalar_advection_2d.jl:233-238)
    TB_DEV double analysis_integrand(int, const double (&)[1], const double (&)[1]) const { return nan(""); }
    TB_DEV void flux_normal(const double (&u)[1], const double (&n)[ND], double (&f)[1]) const {
        f[0] = a_dot(n) * u[0];
    }
    TB_DEV void numflux_normal(int id, const double (&ul)[1], const double (&ur)[1], const double (&n)[ND],
                               double (&f)[1]) const {
        const double an = a_dot(n);
        switch (id) {
        case TRIXI_B200_FLUX_CENTRAL:
            f[0] = 0.5 * (an * ul[0] + an * ur[0]);
            break;
        case TRIXI_B200_FLUX_LLF:
        case TRIXI_B200_FLUX_LLF_NAIVE:  // linear_scalar_advection_2d.jl:241-246
            f[0] = 0.5 * (an * ul[0] + an * ur[0]) + (-0.5 * fabs(an) * (ur[0] - ul[0]));
            break;
        case TRIXI_B200_FLUX_GODUNOV:  // :262-275
            f[0] = an >= 0 ? an * ul[0] : an * ur[0];
            break;
        default:
            f[0] = nan("");
        }
    }
    TB_DEV void slip_wall_normal(const double (&u)[1], const double (&n)[ND], int direction, double (&f)[1]) const {
        f[0] = nan("");
    }
    TB_DEV void slip_wall_outward(const double (&u)[1], const double (&n)[ND], double (&f)[1]) const { f[0] = nan(""); }
    // max_abs_speeds(equation) (linear_scalar_advection_2d.jl:292-294)
    TB_DEV void max_abs_speeds(const double (&u)[1], double (&lam)[ND]) const {
        lam[0] = fabs(a0);
        if constexpr (ND > 1) lam[1] = fabs(a1);
        if constexpr (ND > 2) lam[2] = fabs(a2);
    }
    TB_DEV void source_terms(int id, const double (&u)[1], const double (&x)[ND], double t, double (&s)[1]) const {
        s[0] = 0.0;
    }
    TB_DEV void initial_condition(int id, const double (&x)[ND], double t, double (&u)[1]) const {
        if (id == TRIXI_B200_IC_CONVERGENCE_TEST) {  // linear_scalar_advection_2d.jl:67-80
            double s = x[0] - a0 * t;
            if constexpr (ND > 1) s += x[1] - a1 * t;
            if constexpr (ND > 2) s += x[2] - a2 * t;
            u[0] = 1 + 0.5 * sin(2 * M_PI * 0.5 * s);
        } else if (id == TRIXI_B200_IC_CONSTANT) {
            u[0] = 2.0;
        } else {
            u[0] = nan("");
        }
    }
    TB_DEV void slip_wall(const double (&u)[1], int o, int direction, double (&f)[1]) const { f[0] = nan(""); }
};

// ---- compressible Euler (compressible_euler_2d.jl, compressible_euler_3d.jl) -------------------------
template <int ND>
struct Euler {
    static constexpr int NDIMS = ND, NVARS = ND + 2;
    static constexpr bool kConstantSpeed = false;
    static constexpr bool kHasNoncons = false;
    double gamma, inv_gm1;
    __host__ __device__ explicit Euler(const EqParams &q) : gamma(q.p[0]), inv_gm1(q.p[1]) {}

    // cons2prim (compressible_euler_3d.jl:1783-1793)
    TB_DEV void cons2prim(const double (&u)[NVARS], double &rho, double (&v)[ND], double &p) const {
        rho = u[0];
        double kin = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            v[d] = u[1 + d] / u[0];
            kin += u[1 + d] * v[d];
        }
        p = (gamma - 1) * (u[ND + 1] - 0.5 * kin);
    }

    // Roe averages of min_max_speed_einfeldt (compressible_euler_3d.jl:1675-1686): velocity and sound speed
    TB_DEV void roe_average(const double (&ul)[NVARS], const double (&ur)[NVARS], double rho_ll, const double (&v_ll)[ND],
                            double p_ll, double rho_rr, const double (&v_rr)[ND], double p_rr, double (&v_roe)[ND],
                            double &c_roe) const {
        const double H_ll = (ul[ND + 1] + p_ll) / rho_ll, H_rr = (ur[ND + 1] + p_rr) / rho_rr;
        const double sqrt_rho_ll = sqrt(rho_ll), sqrt_rho_rr = sqrt(rho_rr);
        const double inv_sum_sqrt_rho = 1.0 / (sqrt_rho_ll + sqrt_rho_rr);
        double v_roe_mag = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            v_roe[d] = (sqrt_rho_ll * v_ll[d] + sqrt_rho_rr * v_rr[d]) * inv_sum_sqrt_rho;
            v_roe_mag += v_roe[d] * v_roe[d];
        }
        const double H_roe = (sqrt_rho_ll * H_ll + sqrt_rho_rr * H_rr) * inv_sum_sqrt_rho;
        c_roe = sqrt((gamma - 1) * (H_roe - 0.5 * v_roe_mag));
    }

    // indicator variables of IndicatorHennemannGassner: density_pressure, density, pressure
    // (compressible_euler_3d.jl:1937-1956)
    TB_DEV double indicator_variable(int var, const double (&u)[NVARS]) const {
        double q = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) q = q + u[1 + d] * u[1 + d];
        if (var == TRIXI_B200_INDVAR_DENSITY) return u[0];
        if (var == TRIXI_B200_INDVAR_PRESSURE) return (gamma - 1) * (u[ND + 1] - 0.5 * q / u[0]);
        return (gamma - 1) * (u[0] * u[ND + 1] - 0.5 * q);
    }

    // flux(u, orientation) (compressible_euler_3d.jl:420-447)
    TB_DEV void flux(const double (&u)[NVARS], int o, double (&f)[NVARS]) const {
        double rho, v[ND], p;
        cons2prim(u, rho, v, p);
        double mom[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) mom[d] = u[1 + d];
        const double rv = pick<ND>(mom, o);
        f[0] = rv;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = rv * v[d] + (d == o ? p : 0.0);
        f[ND + 1] = (u[ND + 1] + p) * pick<ND>(v, o);
    }

    // integrands of the AnalysisCallback's analysis_integrals: entropy = entropy_math (:1959-2009), energies
    // (:2012-2023), cons2entropy(u) . du (:1796-1817, analysis_dg3d.jl:506-517); NaN for an unknown id
    TB_DEV double analysis_integrand(int quantity, const double (&u)[NVARS], const double (&du)[NVARS]) const {
        const double rho = u[0];
        double msq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) msq += u[1 + d] * u[1 + d];
        switch (quantity) {
        case TRIXI_B200_INTEGRAL_ENERGY_TOTAL: return u[ND + 1];
        case TRIXI_B200_INTEGRAL_ENERGY_KINETIC: return 0.5 * msq / rho;
        case TRIXI_B200_INTEGRAL_ENERGY_INTERNAL: return u[ND + 1] - 0.5 * msq / rho;
        case TRIXI_B200_INTEGRAL_ENTROPY: {
            const double p = (gamma - 1) * (u[ND + 1] - 0.5 * msq / rho);
            const double s = log(p) - gamma * log(rho);
            return -s * rho * inv_gm1;
        }
        case TRIXI_B200_INTEGRAL_ENTROPY_TIMEDERIVATIVE: {
            double v[ND], v_square = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                v[d] = u[1 + d] / rho;
                v_square += v[d] * v[d];
            }
            const double p = (gamma - 1) * (u[ND + 1] - 0.5 * rho * v_square);
            const double s = log(p) - gamma * log(rho);
            const double rho_p = rho / p;
            double dot = ((gamma - s) * inv_gm1 - 0.5 * rho_p * v_square) * du[0];
#pragma unroll
            for (int d = 0; d < ND; ++d) dot += rho_p * v[d] * du[1 + d];
            return dot + (-rho_p) * du[ND + 1];
        }
        default: return nan("");
        }
    }

    // flux_ranocha (compressible_euler_3d.jl:746-793), generic form: ln_mean/inv_ln_mean per pair
    TB_DEV void flux_ranocha(const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                             double (&f)[NVARS]) const {
        double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
        cons2prim(ul, rho_ll, v_ll, p_ll);
        cons2prim(ur, rho_rr, v_rr, p_rr);
        const double rho_mean = ln_mean(rho_ll, rho_rr);
        const double inv_rho_p_mean = p_ll * p_rr * inv_ln_mean(rho_ll * p_rr, rho_rr * p_ll);
        double v_avg[ND], vsq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
            vsq += v_ll[d] * v_rr[d];
        }
        const double p_avg = 0.5 * (p_ll + p_rr);
        const double velocity_square_avg = 0.5 * vsq;
        const double f1 = rho_mean * pick<ND>(v_avg, o);
        f[0] = f1;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = f1 * v_avg[d] + (d == o ? p_avg : 0.0);
        f[ND + 1] = f1 * (velocity_square_avg + inv_rho_p_mean * inv_gm1) +
                    0.5 * (p_ll * pick<ND>(v_rr, o) + p_rr * pick<ND>(v_ll, o));
    }

    // cons2prim / flux_ranocha with the fast divisions above (tuned surface kernel); same formulas
    TB_DEV void cons2prim_fast(const double (&u)[NVARS], double &rho, double (&v)[ND], double &p) const {
        rho = u[0];
        const double inv_rho = fast_rcp(rho);
        double kin = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double q = u[1 + d] * inv_rho;
            v[d] = fma(fma(-rho, q, u[1 + d]), inv_rho, q);
            kin += u[1 + d] * v[d];
        }
        p = (gamma - 1) * (u[ND + 1] - 0.5 * kin);
    }
    // FluxLaxFriedrichs (numerical_fluxes.jl:37-45,172-178) with max_abs_speed(_naive)
    // (compressible_euler_3d.jl:1112-1177) on Newton reciprocals: two MUFU seeds per face node instead of eight IEEE
    // divisions; every quotient carries a residual correction (within 1 ulp of the generic path)
    // the same flux with both states left in (shared) memory and re-read for the dissipation term: 40 instead of 62
    // registers in the staged interface kernel, i.e. 6 instead of 4 resident blocks per SM
    TB_DEV void llf_fast_prim(const double *pu, double &rho, double (&v)[ND], double &p, double &c) const {
        rho = pu[0];
        const double inv_rho = fast_rcp(rho);
        double kin = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double m = pu[1 + d], q = m * inv_rho;
            v[d] = fma(fma(-rho, q, m), inv_rho, q);
            kin += m * v[d];
        }
        p = (gamma - 1) * (pu[ND + 1] - 0.5 * kin);
        const double gp = gamma * p;
        double c2 = gp * inv_rho;
        c2 = fma(fma(-rho, c2, gp), inv_rho, c2);
        c = fast_sqrt(c2);
    }
    TB_DEV void flux_llf_fast_mem(int id, const double *pl, const double *pr, int o, double (&f)[NVARS]) const {
        double rho_l, v_l[ND], p_l, c_l, rho_r, v_r[ND], p_r, c_r;
        llf_fast_prim(pl, rho_l, v_l, p_l, c_l);
        llf_fast_prim(pr, rho_r, v_r, p_r, c_r);
        const double vo_l = pick<ND>(v_l, o), vo_r = pick<ND>(v_r, o);
        const double lam = id == TRIXI_B200_FLUX_LLF_NAIVE ? fmax(fabs(vo_l), fabs(vo_r)) + fmax(c_l, c_r)
                                                           : fmax(fabs(vo_l) + c_l, fabs(vo_r) + c_r);
        const double hl = -0.5 * lam;
        const double rv_l = pl[1 + o], rv_r = pr[1 + o];
        f[0] = 0.5 * (rv_l + rv_r) + hl * (pr[0] - pl[0]);
#pragma unroll
        for (int d = 0; d < ND; ++d)
            f[1 + d] = 0.5 * ((rv_l * v_l[d] + (d == o ? p_l : 0.0)) + (rv_r * v_r[d] + (d == o ? p_r : 0.0))) +
                       hl * (pr[1 + d] - pl[1 + d]);
        f[ND + 1] = 0.5 * ((pl[ND + 1] + p_l) * vo_l + (pr[ND + 1] + p_r) * vo_r) + hl * (pr[ND + 1] - pl[ND + 1]);
    }
    // (one arithmetic for every caller: the staged, the plain and the MPI interface kernels must agree bit for bit,
    // N ranks reproduce one rank exactly)
    TB_DEV void flux_llf_fast(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                              double (&f)[NVARS]) const {
        flux_llf_fast_mem(id, &ul[0], &ur[0], o, f);
    }
    TB_DEV void flux_ranocha_fast(const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                                  double (&f)[NVARS]) const {
        double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
        cons2prim_fast(ul, rho_ll, v_ll, p_ll);
        cons2prim_fast(ur, rho_rr, v_rr, p_rr);
        const double rho_mean = ln_mean_fast(rho_ll, rho_rr);
        const double inv_rho_p_mean = p_ll * p_rr * inv_ln_mean_fast(rho_ll * p_rr, rho_rr * p_ll);
        double v_avg[ND], vsq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
            vsq += v_ll[d] * v_rr[d];
        }
        const double p_avg = 0.5 * (p_ll + p_rr);
        const double f1 = rho_mean * pick<ND>(v_avg, o);
        f[0] = f1;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = f1 * v_avg[d] + (d == o ? p_avg : 0.0);
        f[ND + 1] = f1 * (0.5 * vsq + inv_rho_p_mean * inv_gm1) +
                    0.5 * (p_ll * pick<ND>(v_rr, o) + p_rr * pick<ND>(v_ll, o));
    }

    // flux_shima_etal (compressible_euler_3d.jl:473-510)
    TB_DEV void flux_shima(const double (&ul)[NVARS], const double (&ur)[NVARS], int o, double (&f)[NVARS]) const {
        double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
        cons2prim(ul, rho_ll, v_ll, p_ll);
        cons2prim(ur, rho_rr, v_rr, p_rr);
        const double rho_avg = 0.5 * (rho_ll + rho_rr), p_avg = 0.5 * (p_ll + p_rr);
        double v_avg[ND], kin = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
            kin += v_ll[d] * v_rr[d];
        }
        const double kin_avg = 0.5 * kin;
        const double pv_avg = 0.5 * (p_ll * pick<ND>(v_rr, o) + p_rr * pick<ND>(v_ll, o));
        const double vn = pick<ND>(v_avg, o);
        const double f1 = rho_avg * vn;
        f[0] = f1;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = f1 * v_avg[d] + (d == o ? p_avg : 0.0);
        f[ND + 1] = p_avg * vn * inv_gm1 + f1 * kin_avg + pv_avg;
    }

    // flux_kennedy_gruber (compressible_euler_3d.jl:560-600)
    TB_DEV void flux_kennedy_gruber(const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                                    double (&f)[NVARS]) const {
        double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
        cons2prim(ul, rho_ll, v_ll, p_ll);
        cons2prim(ur, rho_rr, v_rr, p_rr);
        const double rho_avg = 0.5 * (rho_ll + rho_rr), p_avg = 0.5 * (p_ll + p_rr);
        double v_avg[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
        const double e_avg = 0.5 * (ul[ND + 1] / rho_ll + ur[ND + 1] / rho_rr);
        const double vn = pick<ND>(v_avg, o);
        const double f1 = rho_avg * vn;
        f[0] = f1;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = f1 * v_avg[d] + (d == o ? p_avg : 0.0);
        f[ND + 1] = (rho_avg * e_avg + p_avg) * vn;
    }

    // flux_chandrashekar (compressible_euler_3d.jl:639-690)
    TB_DEV void flux_chandrashekar(const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                                   double (&f)[NVARS]) const {
        double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
        cons2prim(ul, rho_ll, v_ll, p_ll);
        cons2prim(ur, rho_rr, v_rr, p_rr);
        const double beta_ll = 0.5 * rho_ll / p_ll, beta_rr = 0.5 * rho_rr / p_rr;
        double kl = 0.0, kr = 0.0, v_avg[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            kl += v_ll[d] * v_ll[d];
            kr += v_rr[d] * v_rr[d];
            v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
        }
        const double rho_avg = 0.5 * (rho_ll + rho_rr);
        const double rho_mean = ln_mean(rho_ll, rho_rr);
        const double beta_mean = ln_mean(beta_ll, beta_rr);
        const double beta_avg = 0.5 * (beta_ll + beta_rr);
        const double p_mean = 0.5 * rho_avg / beta_avg;
        const double velocity_square_avg = 0.5 * kl + 0.5 * kr;
        const double f1 = rho_mean * pick<ND>(v_avg, o);
        f[0] = f1;
        double s = f1 * 0.5 * (1 / (gamma - 1) / beta_mean - velocity_square_avg);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            f[1 + d] = f1 * v_avg[d] + (d == o ? p_mean : 0.0);
            s += f[1 + d] * v_avg[d];
        }
        f[ND + 1] = s;
    }

    // flux_chandrashekar(u_ll, u_rr, normal_direction) (compressible_euler_3d.jl:693-733, compressible_euler_2d.jl:639-670)
    TB_DEV void flux_chandrashekar_normal(const double (&ul)[NVARS], const double (&ur)[NVARS], const double (&n)[ND],
                                          double (&f)[NVARS]) const {
        double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
        cons2prim(ul, rho_ll, v_ll, p_ll);
        cons2prim(ur, rho_rr, v_rr, p_rr);
        const double beta_ll = 0.5 * rho_ll / p_ll, beta_rr = 0.5 * rho_rr / p_rr;
        double kl = 0.0, kr = 0.0, v_avg[ND], vn_ll = 0.0, vn_rr = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            kl += v_ll[d] * v_ll[d];
            kr += v_rr[d] * v_rr[d];
            v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
            vn_ll += v_ll[d] * n[d];
            vn_rr += v_rr[d] * n[d];
        }
        const double rho_avg = 0.5 * (rho_ll + rho_rr);
        const double rho_mean = ln_mean(rho_ll, rho_rr);
        const double beta_mean = ln_mean(beta_ll, beta_rr);
        const double beta_avg = 0.5 * (beta_ll + beta_rr);
        const double p_mean = 0.5 * rho_avg / beta_avg;
        const double velocity_square_avg = 0.5 * kl + 0.5 * kr;
        const double f1 = rho_mean * 0.5 * (vn_ll + vn_rr);
        f[0] = f1;
        double s = f1 * 0.5 * (1 / (gamma - 1) / beta_mean - velocity_square_avg);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            f[1 + d] = f1 * v_avg[d] + p_mean * n[d];
            s += f[1 + d] * v_avg[d];
        }
        f[ND + 1] = s;
    }

    // flux_hllc (compressible_euler_3d.jl:1423-1541 with an orientation, :1543-1665 along a normal direction;
    // compressible_euler_2d.jl:1720-1925).  NORMAL = false: orientation o, n is not looked at
    template <bool NORMAL>
    TB_DEV void flux_hllc(const double (&ul)[NVARS], const double (&ur)[NVARS], int o, const double (&n)[ND],
                          double (&f)[NVARS]) const {
        const double rho_ll = ul[0], rho_rr = ur[0];
        double v_ll[ND], v_rr[ND], vsq_ll = 0.0, vsq_rr = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            v_ll[d] = ul[1 + d] / rho_ll;
            v_rr[d] = ur[1 + d] / rho_rr;
            vsq_ll += v_ll[d] * v_ll[d];
            vsq_rr += v_rr[d] * v_rr[d];
        }
        const double e_ll = ul[ND + 1] / rho_ll, e_rr = ur[ND + 1] / rho_rr;
        double p_ll, p_rr, vel_L, vel_R, norm_ = 1.0, norm_sq = 1.0, inv_norm_sq = 1.0;
        double f_ll[NVARS], f_rr[NVARS];
        if constexpr (NORMAL) {
            double r, vv[ND];
            cons2prim(ul, r, vv, p_ll);
            cons2prim(ur, r, vv, p_rr);
            vel_L = 0.0;
            vel_R = 0.0;
            double nsq = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                vel_L += v_ll[d] * n[d];
                vel_R += v_rr[d] * n[d];
                nsq += n[d] * n[d];
            }
            norm_ = sqrt(nsq);
            norm_sq = norm_ * norm_;
            inv_norm_sq = 1.0 / norm_sq;
            flux_normal(ul, n, f_ll);
            flux_normal(ur, n, f_rr);
        } else {
            p_ll = (gamma - 1) * (ul[ND + 1] - 0.5 * rho_ll * vsq_ll);
            p_rr = (gamma - 1) * (ur[ND + 1] - 0.5 * rho_rr * vsq_rr);
            vel_L = pick<ND>(v_ll, o);
            vel_R = pick<ND>(v_rr, o);
            flux(ul, o, f_ll);
            flux(ur, o, f_rr);
        }
        double c_ll = sqrt(gamma * p_ll / rho_ll), c_rr = sqrt(gamma * p_rr / rho_rr);
        if constexpr (NORMAL) {
            c_ll = c_ll * norm_;
            c_rr = c_rr * norm_;
        }
        const double sqrt_rho_ll = sqrt(rho_ll), sqrt_rho_rr = sqrt(rho_rr), sum_sqrt_rho = sqrt_rho_ll + sqrt_rho_rr;
        double vel_roe, vel_roe_mag = 0.0;
        if constexpr (NORMAL) {
            vel_roe = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double vr = (sqrt_rho_ll * v_ll[d] + sqrt_rho_rr * v_rr[d]) / sum_sqrt_rho;
                vel_roe += vr * n[d];
                vel_roe_mag += vr * vr;
            }
        } else {
            vel_roe = (sqrt_rho_ll * vel_L + sqrt_rho_rr * vel_R) / sum_sqrt_rho;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double w = sqrt_rho_ll * v_ll[d] + sqrt_rho_rr * v_rr[d];
                vel_roe_mag += w * w;
            }
            vel_roe_mag = vel_roe_mag / (sum_sqrt_rho * sum_sqrt_rho);
        }
        const double H_ll = (ul[ND + 1] + p_ll) / rho_ll, H_rr = (ur[ND + 1] + p_rr) / rho_rr;
        const double H_roe = (sqrt_rho_ll * H_ll + sqrt_rho_rr * H_rr) / sum_sqrt_rho;
        double c_roe = sqrt((gamma - 1) * (H_roe - 0.5 * vel_roe_mag));
        if constexpr (NORMAL) c_roe = c_roe * norm_;
        const double Ssl = fmin(vel_L - c_ll, vel_roe - c_roe), Ssr = fmax(vel_R + c_rr, vel_roe + c_roe);
        const double sMu_L = Ssl - vel_L, sMu_R = Ssr - vel_R;
        if (Ssl >= 0) {
#pragma unroll
            for (int v = 0; v < NVARS; ++v) f[v] = f_ll[v];
            return;
        }
        if (Ssr <= 0) {
#pragma unroll
            for (int v = 0; v < NVARS; ++v) f[v] = f_rr[v];
            return;
        }
        const double SStar =
            NORMAL ? (rho_ll * vel_L * sMu_L - rho_rr * vel_R * sMu_R + (p_rr - p_ll) * norm_sq) / (rho_ll * sMu_L - rho_rr * sMu_R)
                   : (p_rr - p_ll + rho_ll * vel_L * sMu_L - rho_rr * vel_R * sMu_R) / (rho_ll * sMu_L - rho_rr * sMu_R);
        const bool left = Ssl <= 0 && 0 <= SStar;
        const double rho_s = left ? rho_ll : rho_rr, sMu = left ? sMu_L : sMu_R, Ss = left ? Ssl : Ssr;
        const double vel_s = left ? vel_L : vel_R, e_s = left ? e_ll : e_rr, p_s = left ? p_ll : p_rr;
        const double densStar = rho_s * sMu / (Ss - SStar);
        double UStar[NVARS];
        UStar[0] = densStar;
        if constexpr (NORMAL) {
            const double enerStar = e_s + (SStar - vel_s) * (SStar * inv_norm_sq + p_s / (rho_s * sMu));
#pragma unroll
            for (int d = 0; d < ND; ++d)
                UStar[1 + d] = densStar * ((left ? v_ll[d] : v_rr[d]) + (SStar - vel_s) * n[d] * inv_norm_sq);
            UStar[ND + 1] = densStar * enerStar;
        } else {
            const double enerStar = e_s + (SStar - vel_s) * (SStar + p_s / (rho_s * sMu));
#pragma unroll
            for (int d = 0; d < ND; ++d) UStar[1 + d] = densStar * (d == o ? SStar : (left ? v_ll[d] : v_rr[d]));
            UStar[ND + 1] = densStar * enerStar;
        }
#pragma unroll
        for (int v = 0; v < NVARS; ++v) f[v] = (left ? f_ll[v] : f_rr[v]) + Ss * (UStar[v] - (left ? ul[v] : ur[v]));
    }

    // Euler<ND>::numflux / numflux_normal hold the core switch; EulerAllFluxes<ND> below holds the full one.  nvcc sizes
    // a kernel's registers for the longest body of an inlined run-time switch whether or not a run selects it: with
    // flux_hllc, flux_hlle and flux_chandrashekar along normals in the default switch the curved interface kernels grew
    // from 70-76 to 120 registers (2-4% on the curved workloads).  `create` picks the launcher table of the second type
    // when a descriptor names one of those fluxes.
    TB_DEV void numflux(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                        double (&f)[NVARS]) const {
        numflux_impl<false>(id, ul, ur, o, f);
    }
    // The same switch without flux_hllc and flux_hlle: what the tuned line-sweep kernels inline for their run-time
    // volume / subcell fluxes.  Their register budget is tuned (with the two long bodies inlined the shock-capturing
    // kernel spilled: 0.88 -> 0.96 ms per launch); set-ups with those fluxes as volume fluxes take the generic kernels.
    TB_DEV void numflux_core(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                             double (&f)[NVARS]) const {
        numflux_impl<false>(id, ul, ur, o, f);
    }
    TB_DEV_HOST static bool in_core_switch(int id) { return id != TRIXI_B200_FLUX_HLLC && id != TRIXI_B200_FLUX_HLLE; }
    template <bool FULL>
    TB_DEV void numflux_impl(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], int o,
                             double (&f)[NVARS]) const {
        if constexpr (!FULL) {
            if (!in_core_switch(id)) {
#pragma unroll
                for (int v = 0; v < NVARS; ++v) f[v] = nan("");
                return;
            }
        }
        switch (id) {
        case TRIXI_B200_FLUX_HLLC: {
            if constexpr (FULL) {
                const double nn[ND] = {};
                flux_hllc<false>(ul, ur, o, nn, f);
            }
            break;
        }
        case TRIXI_B200_FLUX_CENTRAL: {  // numerical_fluxes.jl:17-25
            double fl[NVARS], fr[NVARS];
            flux(ul, o, fl);
            flux(ur, o, fr);
#pragma unroll
            for (int v = 0; v < NVARS; ++v) f[v] = 0.5 * (fl[v] + fr[v]);
            break;
        }
        case TRIXI_B200_FLUX_LLF:
        case TRIXI_B200_FLUX_LLF_NAIVE: {  // numerical_fluxes.jl:37-45,172-178
            double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
            cons2prim(ul, rho_ll, v_ll, p_ll);
            cons2prim(ur, rho_rr, v_rr, p_rr);
            const double c_ll = sqrt(gamma * p_ll / rho_ll), c_rr = sqrt(gamma * p_rr / rho_rr);
            const double vl = pick<ND>(v_ll, o), vr = pick<ND>(v_rr, o);
            // max_abs_speed_naive (compressible_euler_3d.jl:1112-1133) / max_abs_speed (:1156-1177)
            const double lam = id == TRIXI_B200_FLUX_LLF_NAIVE ? fmax(fabs(vl), fabs(vr)) + fmax(c_ll, c_rr)
                                                               : fmax(fabs(vl) + c_ll, fabs(vr) + c_rr);
            double fl[NVARS], fr[NVARS];
            flux(ul, o, fl);
            flux(ur, o, fr);
#pragma unroll
            for (int v = 0; v < NVARS; ++v) f[v] = 0.5 * (fl[v] + fr[v]) + (-0.5 * lam * (ur[v] - ul[v]));
            break;
        }
        case TRIXI_B200_FLUX_HLLE:
        case TRIXI_B200_FLUX_HLL_DAVIS:
        case TRIXI_B200_FLUX_HLL_NAIVE: {  // FluxHLL numerical_fluxes.jl:422-440
            double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
            cons2prim(ul, rho_ll, v_ll, p_ll);
            cons2prim(ur, rho_rr, v_rr, p_rr);
            const double c_ll = sqrt(gamma * p_ll / rho_ll), c_rr = sqrt(gamma * p_rr / rho_rr);
            const double vl = pick<ND>(v_ll, o), vr = pick<ND>(v_rr, o);
            double lmin, lmax;
            if (FULL && id == TRIXI_B200_FLUX_HLLE) {  // min_max_speed_einfeldt :1662-1707
                if constexpr (FULL) {
                    double v_roe[ND];
                    double c_roe;
                    roe_average(ul, ur, rho_ll, v_ll, p_ll, rho_rr, v_rr, p_rr, v_roe, c_roe);
                    const double beta = sqrt(0.5 * (gamma - 1) / gamma), vroe = pick<ND>(v_roe, o);
                    lmin = fmin(fmin(vroe - c_roe, vl - beta * c_ll), 0.0);
                    lmax = fmax(fmax(vroe + c_roe, vr + beta * c_rr), 0.0);
                } else {
                    lmin = lmax = 0.0;
                }
            } else if (id == TRIXI_B200_FLUX_HLL_NAIVE) {  // compressible_euler_3d.jl:1201-1218
                lmin = vl - c_ll;
                lmax = vr + c_rr;
            } else {  // min_max_speed_davis :1240-1261
                lmin = fmin(vl - c_ll, vr - c_rr);
                lmax = fmax(vl + c_ll, vr + c_rr);
            }
            if (lmin >= 0 && lmax >= 0) {
                flux(ul, o, f);
            } else if (lmax <= 0 && lmin <= 0) {
                flux(ur, o, f);
            } else {
                double fl[NVARS], fr[NVARS];
                flux(ul, o, fl);
                flux(ur, o, fr);
                const double inv = 1.0 / (lmax - lmin);
                const double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
#pragma unroll
                for (int v = 0; v < NVARS; ++v)
                    f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
            }
            break;
        }
        case TRIXI_B200_FLUX_RANOCHA:
        case TRIXI_B200_FLUX_RANOCHA_TURBO:
            flux_ranocha(ul, ur, o, f);
            break;
        case TRIXI_B200_FLUX_SHIMA_ETAL:
            flux_shima(ul, ur, o, f);
            break;
        case TRIXI_B200_FLUX_KENNEDY_GRUBER:
            flux_kennedy_gruber(ul, ur, o, f);
            break;
        case TRIXI_B200_FLUX_CHANDRASHEKAR:
            flux_chandrashekar(ul, ur, o, f);
            break;
        default:
#pragma unroll
            for (int v = 0; v < NVARS; ++v) f[v] = nan("");
        }
    }

    // ---- normal-direction versions (curved meshes) ----
    // flux(u, normal_direction) (compressible_euler_3d.jl:449-463)
    TB_DEV void flux_normal(const double (&u)[NVARS], const double (&n)[ND], double (&f)[NVARS]) const {
        double rho, v[ND], p;
        cons2prim(u, rho, v, p);
        double v_normal = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) v_normal += v[d] * n[d];
        const double rho_v_normal = rho * v_normal;
        f[0] = rho_v_normal;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = rho_v_normal * v[d] + p * n[d];
        f[ND + 1] = (u[ND + 1] + p) * v_normal;
    }
    // flux_ranocha(u_ll, u_rr, normal_direction) (compressible_euler_3d.jl:795-828)
    TB_DEV void flux_ranocha_normal(const double (&ul)[NVARS], const double (&ur)[NVARS], const double (&n)[ND],
                                    double (&f)[NVARS]) const {
        double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
        cons2prim_fast(ul, rho_ll, v_ll, p_ll);  // (fast divisions as in flux_ranocha_fast: within 1 ulp)
        cons2prim_fast(ur, rho_rr, v_rr, p_rr);
        double v_dot_n_ll = 0.0, v_dot_n_rr = 0.0, v_avg[ND], vsq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            v_dot_n_ll += v_ll[d] * n[d];
            v_dot_n_rr += v_rr[d] * n[d];
            v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
            vsq += v_ll[d] * v_rr[d];
        }
        const double rho_mean = ln_mean(rho_ll, rho_rr);
        const double inv_rho_p_mean = p_ll * p_rr * inv_ln_mean(rho_ll * p_rr, rho_rr * p_ll);
        const double p_avg = 0.5 * (p_ll + p_rr);
        const double f1 = rho_mean * 0.5 * (v_dot_n_ll + v_dot_n_rr);
        f[0] = f1;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = f1 * v_avg[d] + p_avg * n[d];
        f[ND + 1] = f1 * (0.5 * vsq + inv_rho_p_mean * inv_gm1) + 0.5 * (p_ll * v_dot_n_rr + p_rr * v_dot_n_ll);
    }
    TB_DEV void numflux_normal(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], const double (&n)[ND],
                               double (&f)[NVARS]) const {
        numflux_normal_impl<false>(id, ul, ur, n, f);
    }
    // (flux_chandrashekar along normals is long as well: with the rare fluxes only)
    TB_DEV_HOST static bool in_core_switch_normal(int id) {
        return in_core_switch(id) && id != TRIXI_B200_FLUX_CHANDRASHEKAR;
    }
    template <bool FULL>
    TB_DEV void numflux_normal_impl(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], const double (&n)[ND],
                                    double (&f)[NVARS]) const {
        if constexpr (!FULL) {
            if (!in_core_switch_normal(id)) {
#pragma unroll
                for (int v = 0; v < NVARS; ++v) f[v] = nan("");
                return;
            }
        }
        switch (id) {
        case TRIXI_B200_FLUX_CENTRAL: {
            double fl[NVARS], fr[NVARS];
            flux_normal(ul, n, fl);
            flux_normal(ur, n, fr);
#pragma unroll
            for (int v = 0; v < NVARS; ++v) f[v] = 0.5 * (fl[v] + fr[v]);
            break;
        }
        case TRIXI_B200_FLUX_HLLE: {  // FluxHLL with min_max_speed_einfeldt along a normal direction (:1723-1771)
          if constexpr (FULL) {
            double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
            cons2prim(ul, rho_ll, v_ll, p_ll);
            cons2prim(ur, rho_rr, v_rr, p_rr);
            double vl = 0.0, vr = 0.0, nsq = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                vl += v_ll[d] * n[d];
                vr += v_rr[d] * n[d];
                nsq += n[d] * n[d];
            }
            const double norm_ = sqrt(nsq);
            const double c_ll = sqrt(gamma * p_ll / rho_ll) * norm_, c_rr = sqrt(gamma * p_rr / rho_rr) * norm_;
            double v_roe[ND], c_roe;
            roe_average(ul, ur, rho_ll, v_ll, p_ll, rho_rr, v_rr, p_rr, v_roe, c_roe);
            c_roe = c_roe * norm_;
            double vroe = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) vroe += v_roe[d] * n[d];
            const double beta = sqrt(0.5 * (gamma - 1) / gamma);
            const double lmin = fmin(fmin(vroe - c_roe, vl - beta * c_ll), 0.0);
            const double lmax = fmax(fmax(vroe + c_roe, vr + beta * c_rr), 0.0);
            if (lmin >= 0 && lmax >= 0) {
                flux_normal(ul, n, f);
            } else if (lmax <= 0 && lmin <= 0) {
                flux_normal(ur, n, f);
            } else {
                double fl[NVARS], fr[NVARS];
                flux_normal(ul, n, fl);
                flux_normal(ur, n, fr);
                const double inv = 1.0 / (lmax - lmin);
                const double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
#pragma unroll
                for (int v = 0; v < NVARS; ++v)
                    f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
            }
          }
            break;
        }
        case TRIXI_B200_FLUX_LLF:
        case TRIXI_B200_FLUX_LLF_NAIVE:
        case TRIXI_B200_FLUX_HLL_DAVIS:
        case TRIXI_B200_FLUX_HLL_NAIVE: {
            // primitives, sound speeds and |n| on Newton reciprocals / square roots with residual corrections (within
            // 1 ulp of the IEEE forms; the generic path spent most of the curved interface kernels' time in eight
            // IEEE divisions and three square roots per face node: every flux(u, n) repeated cons2prim), and both
            // physical fluxes from those primitives
            double rho_ll, v_ll[ND], p_ll, c_ll, rho_rr, v_rr[ND], p_rr, c_rr;
            llf_fast_prim(&ul[0], rho_ll, v_ll, p_ll, c_ll);
            llf_fast_prim(&ur[0], rho_rr, v_rr, p_rr, c_rr);
            double vl = 0.0, vr = 0.0, nsq = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                vl += v_ll[d] * n[d];
                vr += v_rr[d] * n[d];
                nsq += n[d] * n[d];
            }
            const double norm_ = fast_sqrt(nsq);
            double fl[NVARS], fr[NVARS];
            // flux(u, normal_direction) (:449-463) from the primitives at hand
            auto flux_prim = [&](double rho, const double(&v)[ND], double pp, double vn, double rho_e, double(&g)[NVARS]) {
                const double rvn = rho * vn;
                g[0] = rvn;
#pragma unroll
                for (int d = 0; d < ND; ++d) g[1 + d] = rvn * v[d] + pp * n[d];
                g[ND + 1] = (rho_e + pp) * vn;
            };
            if (id == TRIXI_B200_FLUX_LLF || id == TRIXI_B200_FLUX_LLF_NAIVE) {
                // max_abs_speed_naive (:1135-1153) / max_abs_speed (:1180-1199)
                const double lam = id == TRIXI_B200_FLUX_LLF_NAIVE
                                       ? fmax(fabs(vl), fabs(vr)) + fmax(c_ll, c_rr) * norm_
                                       : fmax(fabs(vl) + c_ll * norm_, fabs(vr) + c_rr * norm_);
                flux_prim(rho_ll, v_ll, p_ll, vl, ul[ND + 1], fl);
                flux_prim(rho_rr, v_rr, p_rr, vr, ur[ND + 1], fr);
#pragma unroll
                for (int v = 0; v < NVARS; ++v) f[v] = 0.5 * (fl[v] + fr[v]) + (-0.5 * lam * (ur[v] - ul[v]));
            } else {
                // min_max_speed_naive (:1220-1237) / min_max_speed_davis (:1263-1285)
                const double cl = c_ll * norm_, cr = c_rr * norm_;
                double lmin, lmax;
                if (id == TRIXI_B200_FLUX_HLL_NAIVE) {
                    lmin = vl - cl;
                    lmax = vr + cr;
                } else {
                    lmin = fmin(vl - cl, vr - cr);
                    lmax = fmax(vl + cl, vr + cr);
                }
                if (lmin >= 0 && lmax >= 0) {
                    flux_prim(rho_ll, v_ll, p_ll, vl, ul[ND + 1], f);
                } else if (lmax <= 0 && lmin <= 0) {
                    flux_prim(rho_rr, v_rr, p_rr, vr, ur[ND + 1], f);
                } else {
                    flux_prim(rho_ll, v_ll, p_ll, vl, ul[ND + 1], fl);
                    flux_prim(rho_rr, v_rr, p_rr, vr, ur[ND + 1], fr);
                    const double inv = 1.0 / (lmax - lmin);
                    const double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
#pragma unroll
                    for (int v = 0; v < NVARS; ++v)
                        f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
                }
            }
            break;
        }
        case TRIXI_B200_FLUX_RANOCHA:
        case TRIXI_B200_FLUX_RANOCHA_TURBO:
            flux_ranocha_normal(ul, ur, n, f);
            break;
        case TRIXI_B200_FLUX_KENNEDY_GRUBER:
        case TRIXI_B200_FLUX_SHIMA_ETAL: {  // compressible_euler_3d.jl:602-627 / :512-547
            double rho_ll, v_ll[ND], p_ll, rho_rr, v_rr[ND], p_rr;
            cons2prim(ul, rho_ll, v_ll, p_ll);
            cons2prim(ur, rho_rr, v_rr, p_rr);
            const double rho_avg = 0.5 * (rho_ll + rho_rr), p_avg = 0.5 * (p_ll + p_rr);
            double v_avg[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
            if (id == TRIXI_B200_FLUX_KENNEDY_GRUBER) {
                const double e_avg = 0.5 * (ul[ND + 1] / rho_ll + ur[ND + 1] / rho_rr);
                double v_dot_n_avg = 0.0;
#pragma unroll
                for (int d = 0; d < ND; ++d) v_dot_n_avg += v_avg[d] * n[d];
                const double f1 = rho_avg * v_dot_n_avg;
                f[0] = f1;
#pragma unroll
                for (int d = 0; d < ND; ++d) f[1 + d] = f1 * v_avg[d] + p_avg * n[d];
                f[ND + 1] = f1 * e_avg + p_avg * v_dot_n_avg;
            } else {
                double vl = 0.0, vr = 0.0, vsq = 0.0;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    vl += v_ll[d] * n[d];
                    vr += v_rr[d] * n[d];
                    vsq += v_ll[d] * v_rr[d];
                }
                const double v_dot_n_avg = 0.5 * (vl + vr);
                const double f1 = rho_avg * v_dot_n_avg;
                f[0] = f1;
#pragma unroll
                for (int d = 0; d < ND; ++d) f[1 + d] = f1 * v_avg[d] + p_avg * n[d];
                f[ND + 1] = f1 * (0.5 * vsq) + p_avg * v_dot_n_avg * inv_gm1 + 0.5 * (p_ll * vr + p_rr * vl);
            }
            break;
        }
        case TRIXI_B200_FLUX_CHANDRASHEKAR:
            if constexpr (FULL) flux_chandrashekar_normal(ul, ur, n, f);
            break;
        case TRIXI_B200_FLUX_HLLC:
            if constexpr (FULL) flux_hllc<true>(ul, ur, 0, n, f);
            break;
        default:
#pragma unroll
            for (int v = 0; v < NVARS; ++v) f[v] = nan("");
        }
    }
    // boundary_condition_slip_wall(u_inner, normal_direction, direction, ...) for StructuredMesh
    // (compressible_euler_3d.jl:315-366,398-414)
    TB_DEV void slip_wall_normal(const double (&u)[NVARS], const double (&n)[ND], int direction,
                                 double (&f)[NVARS]) const {
        const double sgn = (direction % 2 == 1) ? -1.0 : 1.0;  // outward normal = sgn * n
        double nsq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) nsq += n[d] * n[d];
        const double norm_ = sqrt(nsq);
        double rho, v[ND], p_local;
        cons2prim(u, rho, v, p_local);
        double v_normal = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) v_normal += v[d] * (sgn * n[d] / norm_);
        double p_star;
        if (v_normal <= 0) {
            const double sound_speed = sqrt(gamma * p_local / rho);
            const double base = 1 + 0.5 * (gamma - 1) * v_normal / sound_speed;
            p_star = base >= 0 ? p_local * pow(base, 2 * gamma * inv_gm1) : 0.0;
        } else {
            const double A = 2 / ((gamma + 1) * rho);
            const double B = p_local * (gamma - 1) / (gamma + 1);
            p_star = p_local + 0.5 * v_normal / A * (v_normal + sqrt(v_normal * v_normal + 4 * A * (p_local + B)));
        }
        // p_star * (sgn n / |n|) * |n|, negated again on the - side: net p_star * n
        f[0] = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) f[1 + d] = p_star * (n[d] / norm_) * norm_;
        f[ND + 1] = 0.0;
    }

    // boundary_condition_slip_wall(u_inner, outward normal, x, t, ...) for P4estMesh (:315-366)
    TB_DEV void slip_wall_outward(const double (&u)[NVARS], const double (&n)[ND], double (&f)[NVARS]) const {
        slip_wall_normal(u, n, 2, f);  // even direction: the normal is used as is
    }

    // max_abs_speeds (compressible_euler_3d.jl:1770-1775)
    TB_DEV void max_abs_speeds(const double (&u)[NVARS], double (&lam)[ND]) const {
        double rho, v[ND], p;
        cons2prim(u, rho, v, p);
        const double c = sqrt(gamma * p / rho);
#pragma unroll
        for (int d = 0; d < ND; ++d) lam[d] = fabs(v[d]) + c;
    }

    TB_DEV void source_terms(int id, const double (&u)[NVARS], const double (&x)[ND], double t,
                             double (&s)[NVARS]) const {
        double xs = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) xs += x[d];
        if (id == TRIXI_B200_SRC_CONVERGENCE_TEST) {
            const double omega = 2 * M_PI * 0.5;
            double si, co;
            sincos(omega * (xs - t), &si, &co);
            const double rho = 2 + 0.1 * si;
            const double rho_x = omega * 0.1 * co;
            if constexpr (ND == 3) {  // compressible_euler_3d.jl:127-153
                const double tmp = (2 * rho - 1.5) * (gamma - 1);
                s[0] = 2 * rho_x;
                s[1] = s[2] = s[3] = rho_x * (2 + tmp);
                s[4] = rho_x * (4 * rho + 3 * tmp);
            } else {  // compressible_euler_2d.jl:120-145
                const double tmp = (2 * rho - 1) * (gamma - 1);
                s[0] = rho_x;
                s[1] = s[2] = rho_x * (1 + tmp);
                s[3] = 2 * rho_x * (rho + tmp);
            }
        } else if (id == TRIXI_B200_SRC_EOC_TEST_EULER) {
            double si, co;
            sincospi(xs - t, &si, &co);
            const double rhox = 0.1 * M_PI * co, rho = 2 + 0.1 * si;
            if constexpr (ND == 3) {  // compressible_euler_3d.jl:265-284
                const double C_grav = -4.0 * 1 / (3 * M_PI);
                s[0] = rhox * 2;
                s[1] = s[2] = s[3] = rhox * (2 - C_grav * rho);
                s[4] = rhox * (3 - 5 * C_grav * rho);
            } else {  // compressible_euler_2d.jl:272-292
                const double C_grav = -2.0 * 1 / M_PI;
                s[0] = rhox;
                s[1] = s[2] = rhox * (1 - C_grav * rho);
                s[3] = rhox * (1 - 3 * C_grav * rho);
            }
        } else if (id == TRIXI_B200_SRC_EOC_TEST_COUPLED_EULER_GRAVITY) {
            double si, co;
            sincospi(xs - t, &si, &co);
            const double rhox = 0.1 * M_PI * co, rho = 2 + 0.1 * si;
            if constexpr (ND == 3) {  // compressible_euler_3d.jl:228-249
                const double C_grav = -4.0 * 1 / (3 * M_PI);
                s[0] = s[1] = s[2] = s[3] = 2 * rhox;
                s[4] = 2 * rhox * (1.5 - C_grav * rho);
            } else {  // compressible_euler_2d.jl:241-261
                const double C_grav = -2.0 * 1 / M_PI;
                s[0] = s[1] = s[2] = rhox;
                s[3] = (1 - C_grav * rho) * rhox;
            }
        } else {
#pragma unroll
            for (int v = 0; v < NVARS; ++v) s[v] = 0.0;
        }
    }

    TB_DEV void initial_condition(int id, const double (&x)[ND], double t, double (&u)[NVARS]) const {
        if (id == TRIXI_B200_IC_EOC_TEST_COUPLED_EULER_GRAVITY) {
            // compressible_euler_3d.jl:196-215 / compressible_euler_2d.jl:212-230 (gamma = 2 is the caller's business)
            double xs = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) xs += x[d];
            const double ini = 2 + 0.1 * sinpi(xs - t);
            const double p = ND == 3 ? ini * ini * 1 * 2 / (3 * M_PI) : ini * ini * 1 / M_PI;
            u[0] = ini;
#pragma unroll
            for (int d = 0; d < ND; ++d) u[1 + d] = ini * 1.0;
            u[ND + 1] = p * inv_gm1 + 0.5 * (ND * (ini * 1.0) * 1.0);
        } else if (id == TRIXI_B200_IC_CONSTANT) {  // compressible_euler_3d.jl:78-86
            u[0] = 1.0;
            u[1] = 0.1;
            u[2] = -0.2;
            if constexpr (ND == 3) u[3] = 0.7;
            u[ND + 1] = 10.0;
        } else if (id == TRIXI_B200_IC_CONVERGENCE_TEST) {  // compressible_euler_3d.jl:94-111
            double xs = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) xs += x[d];
            const double ini = 2 + 0.1 * sin(2 * M_PI * 0.5 * (xs - t));
#pragma unroll
            for (int v = 0; v <= ND; ++v) u[v] = ini;
            u[ND + 1] = ini * ini;
        } else if (id == TRIXI_B200_IC_WEAK_BLAST_WAVE) {
            // compressible_euler_3d.jl:163-184 / compressible_euler_2d.jl:241-263 (only used by the error norms)
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) r2 += x[d] * x[d];
            const double r = sqrt(r2);
            const bool outside = r > 0.5;
            const double phi = atan2(x[1], x[0]);
            double vel[3] = {0.0, 0.0, 0.0};
            if constexpr (ND == 3) {
                const double theta = r == 0 ? 0.0 : acos(x[2] / r);
                vel[0] = 0.1882 * cos(phi) * sin(theta);
                vel[1] = 0.1882 * sin(phi) * sin(theta);
                vel[2] = 0.1882 * cos(theta);
            } else {
                vel[0] = 0.1882 * cos(phi);
                vel[1] = 0.1882 * sin(phi);
            }
            const double rho = outside ? 1.0 : 1.1691, p = outside ? 1.0 : 1.245;
            double ke = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double vd = outside ? 0.0 : vel[d];
                u[1 + d] = rho * vd;
                ke += rho * vd * vd;
            }
            u[0] = rho;
            u[ND + 1] = p * inv_gm1 + 0.5 * ke;
        } else {
#pragma unroll
            for (int v = 0; v < NVARS; ++v) u[v] = nan("");
        }
    }

    // boundary_condition_slip_wall for TreeMesh (compressible_euler_3d.jl:315-414): unit normal along
    // `o`, outward sign from `direction` (1-based), pressure from the 1D Riemann problem
    TB_DEV void slip_wall(const double (&u)[NVARS], int o, int direction, double (&f)[NVARS]) const {
        const double sgn = (direction % 2 == 1) ? -1.0 : 1.0;  // outward normal = sgn * e_o
        double rho, v[ND], p_local;
        cons2prim(u, rho, v, p_local);
        const double v_normal = sgn * pick<ND>(v, o);
        double p_star;
        if (v_normal <= 0) {
            const double sound_speed = sqrt(gamma * p_local / rho);
            const double base = 1 + 0.5 * (gamma - 1) * v_normal / sound_speed;
            p_star = base >= 0 ? p_local * pow(base, 2 * gamma * inv_gm1) : 0.0;
        } else {
            const double A = 2 / ((gamma + 1) * rho);
            const double B = p_local * (gamma - 1) / (gamma + 1);
            p_star = p_local + 0.5 * v_normal / A * (v_normal + sqrt(v_normal * v_normal + 4 * A * (p_local + B)));
        }
        // flux = p_star * n_outward, then negated again on the - side (:398-414): net +p_star e_o
#pragma unroll
        for (int v2 = 0; v2 < NVARS; ++v2) f[v2] = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d)
            if (d == o) f[1 + d] = p_star;
    }
};

// The compressible Euler equations with every registered flux in the run-time switches (see Euler::numflux): the
// equation type of the second launcher table, used when a descriptor names flux_hllc, flux_hlle or, on curved meshes,
// flux_chandrashekar.  Always on the generic kernels (the tuned ones are tied to Euler<3>).
template <int ND>
struct EulerAllFluxes : Euler<ND> {
    using Base = Euler<ND>;
    static constexpr int NVARS = Base::NVARS;
    __host__ __device__ explicit EulerAllFluxes(const EqParams &q) : Base(q) {}
    TB_DEV void numflux(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], int o, double (&f)[NVARS]) const {
        this->template numflux_impl<true>(id, ul, ur, o, f);
    }
    TB_DEV void numflux_normal(int id, const double (&ul)[NVARS], const double (&ur)[NVARS], const double (&n)[ND],
                               double (&f)[NVARS]) const {
        this->template numflux_normal_impl<true>(id, ul, ur, n, f);
    }
};

// ---- ideal GLM-MHD 3D (ideal_glm_mhd_3d.jl) ------------------------------------------------------------
struct Mhd3D {
    static constexpr int NDIMS = 3, NVARS = 9;
    static constexpr bool kConstantSpeed = false;
    static constexpr bool kHasNoncons = true;  // have_nonconservative_terms (ideal_glm_mhd_3d.jl:81)
    double gamma, inv_gm1, c_h;
    __host__ __device__ explicit Mhd3D(const EqParams &q) : gamma(q.p[0]), inv_gm1(q.p[1]), c_h(q.p[2]) {}

    TB_DEV static bool has_noncons(int flux_id) {
        return flux_id == TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL || flux_id == TRIXI_B200_FLUX_LLF_MHD_POWELL ||
               flux_id == TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL || flux_id == TRIXI_B200_FLUX_HLLE_MHD_POWELL ||
               flux_id == TRIXI_B200_FLUX_CENTRAL_MHD_POWELL;
    }
    TB_DEV static double sel3(double a, double b, double c, int o) { return o == 0 ? a : (o == 1 ? b : c); }

    // indicator variables of IndicatorHennemannGassner: density_pressure, density, pressure (:1318-1347)
    TB_DEV double indicator_variable(int var, const double (&u)[9]) const {
        const double mom2 = u[1] * u[1] + u[2] * u[2] + u[3] * u[3];
        const double mag = u[5] * u[5] + u[6] * u[6] + u[7] * u[7], psi2 = u[8] * u[8];
        if (var == TRIXI_B200_INDVAR_DENSITY) return u[0];
        if (var == TRIXI_B200_INDVAR_PRESSURE) return (gamma - 1) * (u[4] - 0.5 * (mom2 / u[0] + mag + psi2));
        return (gamma - 1) * (u[0] * u[4] - 0.5 * (mom2 + u[0] * (mag + psi2)));
    }

    // flux(u, orientation) (:187-234)
    TB_DEV void flux(const double (&u)[9], int o, double (&f)[9]) const {
        const double psi = u[8], inv_rho = fast_rcp(u[0]);
        const double v[3] = {u[1] * inv_rho, u[2] * inv_rho, u[3] * inv_rho};
        const double B[3] = {u[5], u[6], u[7]};
        const double kin_en = 0.5 * (u[1] * v[0] + u[2] * v[1] + u[3] * v[2]);
        const double mag_en = 0.5 * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
        const double pogm1 = u[4] - kin_en - mag_en - 0.5 * psi * psi;
        const double p = (gamma - 1) * pogm1;
        const double rv = sel3(u[1], u[2], u[3], o), vo = sel3(v[0], v[1], v[2], o), Bo = sel3(B[0], B[1], B[2], o);
        f[0] = rv;
#pragma unroll
        for (int d = 0; d < 3; ++d) f[1 + d] = d == o ? rv * vo + p + mag_en - Bo * Bo : rv * v[d] - Bo * B[d];
        f[4] = (kin_en + gamma * pogm1 + 2 * mag_en) * vo - Bo * (v[0] * B[0] + v[1] * B[1] + v[2] * B[2]) +
               c_h * psi * Bo;
#pragma unroll
        for (int d = 0; d < 3; ++d) f[5 + d] = d == o ? c_h * psi : vo * B[d] - v[d] * Bo;
        f[8] = c_h * Bo;
    }

    // flux_nonconservative_powell(u_ll, u_rr, orientation) (:295-340)
    TB_DEV void noncons(const double (&ul)[9], const double (&ur)[9], int o, double (&f)[9]) const {
        const double inv_rho_ll = fast_rcp(ul[0]);
        const double v_ll[3] = {ul[1] * inv_rho_ll, ul[2] * inv_rho_ll, ul[3] * inv_rho_ll};
        const double v_dot_B_ll = v_ll[0] * ul[5] + v_ll[1] * ul[6] + v_ll[2] * ul[7];
        const double Bn_rr = sel3(ur[5], ur[6], ur[7], o), vo = sel3(v_ll[0], v_ll[1], v_ll[2], o);
        f[0] = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) f[1 + d] = ul[5 + d] * Bn_rr;
        f[4] = v_dot_B_ll * Bn_rr + vo * ul[8] * ur[8];
#pragma unroll
        for (int d = 0; d < 3; ++d) f[5 + d] = v_ll[d] * Bn_rr;
        f[8] = vo * ur[8];
    }

    // cons2prim (:1231-1243)
    TB_DEV void cons2prim(const double (&u)[9], double (&q)[9]) const {
        const double rho = u[0], inv_rho = fast_rcp(rho);
        const double v1 = u[1] * inv_rho, v2 = u[2] * inv_rho, v3 = u[3] * inv_rho;
        q[0] = rho;
        q[1] = v1;
        q[2] = v2;
        q[3] = v3;
        q[4] = (gamma - 1) * (u[4] - 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3 + u[5] * u[5] + u[6] * u[6] +
                                            u[7] * u[7] + u[8] * u[8]));
        q[5] = u[5];
        q[6] = u[6];
        q[7] = u[7];
        q[8] = u[8];
    }

    // flux_hindenlang_gassner(u_ll, u_rr, orientation) (:680-779)
    TB_DEV void flux_hindenlang_gassner(const double (&ul)[9], const double (&ur)[9], int o, double (&f)[9]) const {
        double L[9], R[9];
        cons2prim(ul, L);
        cons2prim(ur, R);
        const double rho_mean = ln_mean_fast(L[0], R[0]);
        const double inv_rho_p_mean = L[4] * R[4] * inv_ln_mean_fast(L[0] * R[4], R[0] * L[4]);
        double v_avg[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) v_avg[d] = 0.5 * (L[1 + d] + R[1 + d]);
        const double p_avg = 0.5 * (L[4] + R[4]), psi_avg = 0.5 * (L[8] + R[8]);
        const double velocity_square_avg = 0.5 * (L[1] * R[1] + L[2] * R[2] + L[3] * R[3]);
        const double magnetic_square_avg = 0.5 * (L[5] * R[5] + L[6] * R[6] + L[7] * R[7]);
        const double vo_l = sel3(L[1], L[2], L[3], o), vo_r = sel3(R[1], R[2], R[3], o);
        const double Bo_l = sel3(L[5], L[6], L[7], o), Bo_r = sel3(R[5], R[6], R[7], o);
        const double f1 = rho_mean * sel3(v_avg[0], v_avg[1], v_avg[2], o);
        f[0] = f1;
#pragma unroll
        for (int d = 0; d < 3; ++d)
            f[1 + d] = d == o ? f1 * v_avg[d] + p_avg + magnetic_square_avg - 0.5 * (Bo_l * Bo_r + Bo_r * Bo_l)
                              : f1 * v_avg[d] - 0.5 * (Bo_l * R[5 + d] + Bo_r * L[5 + d]);
#pragma unroll
        for (int d = 0; d < 3; ++d)
            f[5 + d] = d == o ? c_h * psi_avg
                              : 0.5 * (vo_l * L[5 + d] - L[1 + d] * Bo_l + vo_r * R[5 + d] - R[1 + d] * Bo_r);
        f[8] = c_h * 0.5 * (Bo_l + Bo_r);
        // transverse directions in ascending order, as the reference writes them out
        const int t1 = o == 0 ? 1 : 0, t2 = o == 2 ? 1 : 2;
        const double vt1_l = sel3(L[1], L[2], L[3], t1), vt1_r = sel3(R[1], R[2], R[3], t1);
        const double vt2_l = sel3(L[1], L[2], L[3], t2), vt2_r = sel3(R[1], R[2], R[3], t2);
        const double Bt1_l = sel3(L[5], L[6], L[7], t1), Bt1_r = sel3(R[5], R[6], R[7], t1);
        const double Bt2_l = sel3(L[5], L[6], L[7], t2), Bt2_r = sel3(R[5], R[6], R[7], t2);
        f[4] = f1 * (velocity_square_avg + inv_rho_p_mean * inv_gm1) +
               0.5 * (+L[4] * vo_r + R[4] * vo_l + (vo_l * Bt1_l * Bt1_r + vo_r * Bt1_r * Bt1_l) +
                      (vo_l * Bt2_l * Bt2_r + vo_r * Bt2_r * Bt2_l) - (vt1_l * Bo_l * Bt1_r + vt1_r * Bo_r * Bt1_l) -
                      (vt2_l * Bo_l * Bt2_r + vt2_r * Bo_r * Bt2_l) + c_h * (Bo_l * R[8] + Bo_r * L[8]));
    }

    // calc_fast_wavespeed(cons, orientation) (:1350-1376)
    TB_DEV double fast_wavespeed(const double (&u)[9], int o) const {
        // divisions by rho and sqrt(rho) folded into one Newton reciprocal: b_i^2 = B_i^2 / rho
        const double psi = u[8], inv_rho = fast_rcp(u[0]);
        const double v1 = u[1] * inv_rho, v2 = u[2] * inv_rho, v3 = u[3] * inv_rho;
        const double kin_en = 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3);
        const double mag_en = 0.5 * (u[5] * u[5] + u[6] * u[6] + u[7] * u[7]);
        const double p = (gamma - 1) * (u[4] - kin_en - mag_en - 0.5 * psi * psi);
        const double a_square = gamma * p * inv_rho;
        const double b_square = 2 * mag_en * inv_rho;
        const double Bo = sel3(u[5], u[6], u[7], o), sum = a_square + b_square;
        return sqrt(0.5 * sum + 0.5 * sqrt(sum * sum - 4 * a_square * (Bo * Bo * inv_rho)));
    }

    TB_DEV double analysis_integrand(int, const double (&)[9], const double (&)[9]) const { return nan(""); }
    // calc_fast_wavespeed_roe(u_ll, u_rr, orientation) (:1415-1491): Roe averages of Cargo & Gallice
    TB_DEV void fast_wavespeed_roe(const double (&ul)[9], const double (&ur)[9], int o, double &vel_out, double &c_f) const {
        const double inv_rho_ll = 1.0 / ul[0], inv_rho_rr = 1.0 / ur[0];
        const double v_ll[3] = {ul[1] * inv_rho_ll, ul[2] * inv_rho_ll, ul[3] * inv_rho_ll};
        const double v_rr[3] = {ur[1] * inv_rho_rr, ur[2] * inv_rho_rr, ur[3] * inv_rho_rr};
        const double kin_en_ll = 0.5 * (ul[1] * v_ll[0] + ul[2] * v_ll[1] + ul[3] * v_ll[2]);
        const double mag_norm_ll = ul[5] * ul[5] + ul[6] * ul[6] + ul[7] * ul[7];
        const double p_ll = (gamma - 1) * (ul[4] - kin_en_ll - 0.5 * mag_norm_ll - 0.5 * ul[8] * ul[8]);
        const double kin_en_rr = 0.5 * (ur[1] * v_rr[0] + ur[2] * v_rr[1] + ur[3] * v_rr[2]);
        const double mag_norm_rr = ur[5] * ur[5] + ur[6] * ur[6] + ur[7] * ur[7];
        const double p_rr = (gamma - 1) * (ur[4] - kin_en_rr - 0.5 * mag_norm_rr - 0.5 * ur[8] * ur[8]);
        const double p_total_ll = p_ll + 0.5 * mag_norm_ll, p_total_rr = p_rr + 0.5 * mag_norm_rr;
        const double sqrt_rho_ll = sqrt(ul[0]), sqrt_rho_rr = sqrt(ur[0]);
        const double inv_sqrt_rho_add = 1.0 / (sqrt_rho_ll + sqrt_rho_rr), inv_sqrt_rho_prod = 1.0 / (sqrt_rho_ll * sqrt_rho_rr);
        const double rho_ll_roe = sqrt_rho_ll * inv_sqrt_rho_add, rho_rr_roe = sqrt_rho_rr * inv_sqrt_rho_add;
        double v_roe[3], B_roe[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            v_roe[d] = v_ll[d] * rho_ll_roe + v_rr[d] * rho_rr_roe;
            B_roe[d] = ul[5 + d] * rho_ll_roe + ur[5 + d] * rho_rr_roe;
        }
        const double H_ll = (ul[4] + p_total_ll) * inv_rho_ll, H_rr = (ur[4] + p_total_rr) * inv_rho_rr;
        const double H_roe = H_ll * rho_ll_roe + H_rr * rho_rr_roe;
        const double dB0 = ul[5] - ur[5], dB1 = ul[6] - ur[6], dB2 = ul[7] - ur[7];
        const double X = 0.5 * (dB0 * dB0 + dB1 * dB1 + dB2 * dB2) * (inv_sqrt_rho_add * inv_sqrt_rho_add);
        const double b_square_roe = (B_roe[0] * B_roe[0] + B_roe[1] * B_roe[1] + B_roe[2] * B_roe[2]) * inv_sqrt_rho_prod;
        const double a_square_roe =
            (2 - gamma) * X + (gamma - 1) * (H_roe - 0.5 * (v_roe[0] * v_roe[0] + v_roe[1] * v_roe[1] + v_roe[2] * v_roe[2]) -
                                             b_square_roe);
        const double Bo = sel3(B_roe[0], B_roe[1], B_roe[2], o);
        const double c_a_roe = Bo * Bo * inv_sqrt_rho_prod, sum = a_square_roe + b_square_roe;
        const double a_star_roe = sqrt(sum * sum - 4 * a_square_roe * c_a_roe);
        c_f = sqrt(0.5 * (a_square_roe + b_square_roe + a_star_roe));
        vel_out = sel3(v_roe[0], v_roe[1], v_roe[2], o);
    }

    // conservative part of the surface/volume flux (tuples are encoded as one id)
    TB_DEV void numflux(int id, const double (&ul)[9], const double (&ur)[9], int o, double (&f)[9]) const {
        switch (id) {
        case TRIXI_B200_FLUX_HLLE_MHD_POWELL: {  // FluxHLL (numerical_fluxes.jl:422-449), min_max_speed_einfeldt (:1094-1130)
            const double cf_ll = fast_wavespeed(ul, o), cf_rr = fast_wavespeed(ur, o);
            double vel_roe, cf_roe;
            fast_wavespeed_roe(ul, ur, o, vel_roe, cf_roe);
            const double lmin = fmin(sel3(ul[1], ul[2], ul[3], o) / ul[0] - cf_ll, vel_roe - cf_roe);
            const double lmax = fmax(sel3(ur[1], ur[2], ur[3], o) / ur[0] + cf_rr, vel_roe + cf_roe);
            if (lmin >= 0 && lmax >= 0) {
                flux(ul, o, f);
            } else if (lmax <= 0 && lmin <= 0) {
                flux(ur, o, f);
            } else {
                double fl[9], fr[9];
                flux(ul, o, fl);
                flux(ur, o, fr);
                const double inv = 1.0 / (lmax - lmin);
                const double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
#pragma unroll
                for (int v = 0; v < 9; ++v) f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
            }
            break;
        }
        case TRIXI_B200_FLUX_CENTRAL:
        case TRIXI_B200_FLUX_CENTRAL_MHD_POWELL: {
            double fl[9], fr[9];
            flux(ul, o, fl);
            flux(ur, o, fr);
#pragma unroll
            for (int v = 0; v < 9; ++v) f[v] = 0.5 * (fl[v] + fr[v]);
            break;
        }
        case TRIXI_B200_FLUX_LLF:
        case TRIXI_B200_FLUX_LLF_NAIVE:
        case TRIXI_B200_FLUX_LLF_MHD_POWELL:
        case TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL: {  // max_abs_speed(_naive) (:857-928)
            const bool naive = id == TRIXI_B200_FLUX_LLF_NAIVE || id == TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL;
            const double v_ll = sel3(ul[1], ul[2], ul[3], o) * fast_rcp(ul[0]);
            const double v_rr = sel3(ur[1], ur[2], ur[3], o) * fast_rcp(ur[0]);
            const double cf_ll = fast_wavespeed(ul, o), cf_rr = fast_wavespeed(ur, o);
            const double lam = naive ? fmax(fabs(v_ll), fabs(v_rr)) + fmax(cf_ll, cf_rr)
                                     : fmax(fabs(v_ll) + cf_ll, fabs(v_rr) + cf_rr);
            double fl[9], fr[9];
            flux(ul, o, fl);
            flux(ur, o, fr);
#pragma unroll
            for (int v = 0; v < 9; ++v) f[v] = 0.5 * (fl[v] + fr[v]) + (-0.5 * lam * (ur[v] - ul[v]));
            break;
        }
        case TRIXI_B200_FLUX_HINDENLANG_GASSNER:
        case TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL:
            flux_hindenlang_gassner(ul, ur, o, f);
            break;
        default:
#pragma unroll
            for (int v = 0; v < 9; ++v) f[v] = nan("");
        }
    }
    // max_abs_speeds (:1218-1228)
    TB_DEV void max_abs_speeds(const double (&u)[9], double (&lam)[3]) const {
#pragma unroll
        for (int d = 0; d < 3; ++d) lam[d] = fabs(u[1 + d] * fast_rcp(u[0])) + fast_wavespeed(u, d);
    }
    TB_DEV void source_terms(int id, const double (&u)[9], const double (&x)[3], double t, double (&s)[9]) const {
#pragma unroll
        for (int v = 0; v < 9; ++v) s[v] = 0.0;
    }
    TB_DEV void initial_condition(int id, const double (&x)[3], double t, double (&u)[9]) const {
        if (id == TRIXI_B200_IC_CONVERGENCE_TEST) {  // Alfven wave (:124-151) + prim2cons (:1273-1284)
            const double p = 1, omega = 2 * 3.141592653589793, r = 2, e = 0.2;
            const double nx = 1 / sqrt(r * r + 1), ny = r / sqrt(r * r + 1), sqr = 1;
            const double Va = omega / (ny * sqr);
            const double phi_alv = omega / ny * (nx * (x[0] - 0.5 * r) + ny * (x[1] - 0.5 * r)) - Va * t;
            double sn, cs;
            sincos(phi_alv, &sn, &cs);
            const double rho = 1;
            const double v1 = -e * ny * cs / rho, v2 = e * nx * cs / rho, v3 = e * sn / rho;
            const double B1 = nx - rho * v1 * sqr, B2 = ny - rho * v2 * sqr, B3 = -rho * v3 * sqr, psi = 0;
            u[0] = rho, u[1] = rho * v1, u[2] = rho * v2, u[3] = rho * v3;
            u[4] = p * inv_gm1 + 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3) + 0.5 * (B1 * B1 + B2 * B2 + B3 * B3) + 0.5 * psi * psi;
            u[5] = B1, u[6] = B2, u[7] = B3, u[8] = psi;
            return;
        }
        const double c[9] = {1.0, 0.1, -0.2, -0.5, 50.0, 3.0, -1.2, 0.5, 0.0};  // initial_condition_constant (:101-113)
#pragma unroll
        for (int v = 0; v < 9; ++v) u[v] = id == TRIXI_B200_IC_CONSTANT ? c[v] : nan("");
    }
    // ---- curved meshes: fluxes along a (non-normalised) normal vector -----------------------------------------
    // flux(u, normal_direction) (:236-275)
    TB_DEV void flux_normal(const double (&u)[9], const double (&n)[3], double (&f)[9]) const {
        const double rho = u[0], psi = u[8], inv_rho = 1.0 / rho;
        const double v1 = u[1] * inv_rho, v2 = u[2] * inv_rho, v3 = u[3] * inv_rho;
        const double B1 = u[5], B2 = u[6], B3 = u[7];
        const double kin_en = 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3);
        const double mag_en = 0.5 * (B1 * B1 + B2 * B2 + B3 * B3);
        const double pogm1 = u[4] - kin_en - mag_en - 0.5 * psi * psi;
        const double p = (gamma - 1) * pogm1;
        const double v_normal = v1 * n[0] + v2 * n[1] + v3 * n[2];
        const double B_normal = B1 * n[0] + B2 * n[1] + B3 * n[2];
        const double rho_v_normal = rho * v_normal;
        f[0] = rho_v_normal;
        f[1] = rho_v_normal * v1 - B1 * B_normal + (p + mag_en) * n[0];
        f[2] = rho_v_normal * v2 - B2 * B_normal + (p + mag_en) * n[1];
        f[3] = rho_v_normal * v3 - B3 * B_normal + (p + mag_en) * n[2];
        f[4] = (kin_en + gamma * pogm1 + 2 * mag_en) * v_normal - B_normal * (v1 * B1 + v2 * B2 + v3 * B3) + c_h * psi * B_normal;
        f[5] = c_h * psi * n[0] + (v2 * B1 - v1 * B2) * n[1] + (v3 * B1 - v1 * B3) * n[2];
        f[6] = (v1 * B2 - v2 * B1) * n[0] + c_h * psi * n[1] + (v3 * B2 - v2 * B3) * n[2];
        f[7] = (v1 * B3 - v3 * B1) * n[0] + (v2 * B3 - v3 * B2) * n[1] + c_h * psi * n[2];
        f[8] = c_h * B_normal;
    }
    // flux_nonconservative_powell(u_ll, u_rr, normal_direction) (:342-374)
    TB_DEV void noncons_normal(const double (&ul)[9], const double (&ur)[9], const double (&n)[3], double (&f)[9]) const {
        const double inv_rho_ll = 1.0 / ul[0];
        const double v1 = ul[1] * inv_rho_ll, v2 = ul[2] * inv_rho_ll, v3 = ul[3] * inv_rho_ll;
        const double v_dot_B_ll = v1 * ul[5] + v2 * ul[6] + v3 * ul[7];
        const double v_dot_n_ll = v1 * n[0] + v2 * n[1] + v3 * n[2];
        const double B_dot_n_rr = ur[5] * n[0] + ur[6] * n[1] + ur[7] * n[2];
        f[0] = 0.0;
        f[1] = ul[5] * B_dot_n_rr, f[2] = ul[6] * B_dot_n_rr, f[3] = ul[7] * B_dot_n_rr;
        f[4] = v_dot_B_ll * B_dot_n_rr + v_dot_n_ll * ul[8] * ur[8];
        f[5] = v1 * B_dot_n_rr, f[6] = v2 * B_dot_n_rr, f[7] = v3 * B_dot_n_rr;
        f[8] = v_dot_n_ll * ur[8];
    }
    // flux_hindenlang_gassner(u_ll, u_rr, normal_direction) (:781-855)
    TB_DEV void flux_hindenlang_gassner_normal(const double (&ul)[9], const double (&ur)[9], const double (&n)[3],
                                               double (&f)[9]) const {
        double L[9], R[9];
        cons2prim(ul, L);
        cons2prim(ur, R);
        const double v_dot_n_ll = L[1] * n[0] + L[2] * n[1] + L[3] * n[2], v_dot_n_rr = R[1] * n[0] + R[2] * n[1] + R[3] * n[2];
        const double B_dot_n_ll = L[5] * n[0] + L[6] * n[1] + L[7] * n[2], B_dot_n_rr = R[5] * n[0] + R[6] * n[1] + R[7] * n[2];
        const double rho_mean = ln_mean_fast(L[0], R[0]);
        const double inv_rho_p_mean = L[4] * R[4] * inv_ln_mean_fast(L[0] * R[4], R[0] * L[4]);
        const double p_avg = 0.5 * (L[4] + R[4]), psi_avg = 0.5 * (L[8] + R[8]);
        const double velocity_square_avg = 0.5 * (L[1] * R[1] + L[2] * R[2] + L[3] * R[3]);
        const double magnetic_square_avg = 0.5 * (L[5] * R[5] + L[6] * R[6] + L[7] * R[7]);
        const double f1 = rho_mean * 0.5 * (v_dot_n_ll + v_dot_n_rr);
        f[0] = f1;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            f[1 + d] = f1 * (0.5 * (L[1 + d] + R[1 + d])) + (p_avg + magnetic_square_avg) * n[d] -
                       0.5 * (B_dot_n_ll * R[5 + d] + B_dot_n_rr * L[5 + d]);
            f[5 + d] = c_h * psi_avg * n[d] + 0.5 * (v_dot_n_ll * L[5 + d] - L[1 + d] * B_dot_n_ll + v_dot_n_rr * R[5 + d] -
                                                     R[1 + d] * B_dot_n_rr);
        }
        f[8] = c_h * 0.5 * (B_dot_n_ll + B_dot_n_rr);
        f[4] = f1 * (velocity_square_avg + inv_rho_p_mean * inv_gm1) +
               0.5 * (+L[4] * v_dot_n_rr + R[4] * v_dot_n_ll + (v_dot_n_ll * L[5] * R[5] + v_dot_n_rr * R[5] * L[5]) +
                      (v_dot_n_ll * L[6] * R[6] + v_dot_n_rr * R[6] * L[6]) + (v_dot_n_ll * L[7] * R[7] + v_dot_n_rr * R[7] * L[7]) -
                      (L[1] * B_dot_n_ll * R[5] + R[1] * B_dot_n_rr * L[5]) - (L[2] * B_dot_n_ll * R[6] + R[2] * B_dot_n_rr * L[6]) -
                      (L[3] * B_dot_n_ll * R[7] + R[3] * B_dot_n_rr * L[7]) + c_h * (B_dot_n_ll * R[8] + B_dot_n_rr * L[8]));
    }
    // calc_fast_wavespeed(cons, normal_direction) (:1378-1404)
    TB_DEV double fast_wavespeed_normal(const double (&u)[9], const double (&n)[3]) const {
        const double psi = u[8], inv_rho = 1.0 / u[0];
        const double v1 = u[1] * inv_rho, v2 = u[2] * inv_rho, v3 = u[3] * inv_rho;
        const double kin_en = 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3);
        const double mag_en = 0.5 * (u[5] * u[5] + u[6] * u[6] + u[7] * u[7]);
        const double p = (gamma - 1) * (u[4] - kin_en - mag_en - 0.5 * psi * psi);
        const double a_square = gamma * p * inv_rho;
        const double inv_sqrt_rho = 1.0 / sqrt(u[0]);
        const double b1 = u[5] * inv_sqrt_rho, b2 = u[6] * inv_sqrt_rho, b3 = u[7] * inv_sqrt_rho;
        const double b_square = b1 * b1 + b2 * b2 + b3 * b3;
        const double norm_squared = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        const double bn = b1 * n[0] + b2 * n[1] + b3 * n[2];
        const double b_dot_n_squared = bn * bn / norm_squared, sum = a_square + b_square;
        return sqrt((0.5 * sum + 0.5 * sqrt(sum * sum - 4 * a_square * b_dot_n_squared)) * norm_squared);
    }
    // calc_fast_wavespeed_roe(u_ll, u_rr, normal_direction) (:1493-1565)
    TB_DEV void fast_wavespeed_roe_normal(const double (&ul)[9], const double (&ur)[9], const double (&n)[3], double &vel_out,
                                          double &c_f) const {
        const double inv_rho_ll = 1.0 / ul[0], inv_rho_rr = 1.0 / ur[0];
        const double v_ll[3] = {ul[1] * inv_rho_ll, ul[2] * inv_rho_ll, ul[3] * inv_rho_ll};
        const double v_rr[3] = {ur[1] * inv_rho_rr, ur[2] * inv_rho_rr, ur[3] * inv_rho_rr};
        const double kin_en_ll = 0.5 * (ul[1] * v_ll[0] + ul[2] * v_ll[1] + ul[3] * v_ll[2]);
        const double mag_norm_ll = ul[5] * ul[5] + ul[6] * ul[6] + ul[7] * ul[7];
        const double p_ll = (gamma - 1) * (ul[4] - kin_en_ll - 0.5 * mag_norm_ll - 0.5 * ul[8] * ul[8]);
        const double kin_en_rr = 0.5 * (ur[1] * v_rr[0] + ur[2] * v_rr[1] + ur[3] * v_rr[2]);
        const double mag_norm_rr = ur[5] * ur[5] + ur[6] * ur[6] + ur[7] * ur[7];
        const double p_rr = (gamma - 1) * (ur[4] - kin_en_rr - 0.5 * mag_norm_rr - 0.5 * ur[8] * ur[8]);
        const double p_total_ll = p_ll + 0.5 * mag_norm_ll, p_total_rr = p_rr + 0.5 * mag_norm_rr;
        const double sqrt_rho_ll = sqrt(ul[0]), sqrt_rho_rr = sqrt(ur[0]);
        const double inv_sqrt_rho_add = 1.0 / (sqrt_rho_ll + sqrt_rho_rr), inv_sqrt_rho_prod = 1.0 / (sqrt_rho_ll * sqrt_rho_rr);
        const double rho_ll_roe = sqrt_rho_ll * inv_sqrt_rho_add, rho_rr_roe = sqrt_rho_rr * inv_sqrt_rho_add;
        double v_roe[3], B_roe[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            v_roe[d] = v_ll[d] * rho_ll_roe + v_rr[d] * rho_rr_roe;
            B_roe[d] = ul[5 + d] * rho_ll_roe + ur[5 + d] * rho_rr_roe;
        }
        const double H_ll = (ul[4] + p_total_ll) * inv_rho_ll, H_rr = (ur[4] + p_total_rr) * inv_rho_rr;
        const double H_roe = H_ll * rho_ll_roe + H_rr * rho_rr_roe;
        const double dB0 = ul[5] - ur[5], dB1 = ul[6] - ur[6], dB2 = ul[7] - ur[7];
        const double X = 0.5 * (dB0 * dB0 + dB1 * dB1 + dB2 * dB2) * (inv_sqrt_rho_add * inv_sqrt_rho_add);
        const double b_square_roe = (B_roe[0] * B_roe[0] + B_roe[1] * B_roe[1] + B_roe[2] * B_roe[2]) * inv_sqrt_rho_prod;
        const double a_square_roe =
            (2 - gamma) * X + (gamma - 1) * (H_roe - 0.5 * (v_roe[0] * v_roe[0] + v_roe[1] * v_roe[1] + v_roe[2] * v_roe[2]) -
                                             b_square_roe);
        const double norm_squared = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        const double Bn = B_roe[0] * n[0] + B_roe[1] * n[1] + B_roe[2] * n[2];
        const double c_a_roe = Bn * Bn / norm_squared * inv_sqrt_rho_prod, sum = a_square_roe + b_square_roe;
        const double a_star_roe = sqrt(sum * sum - 4 * a_square_roe * c_a_roe);
        c_f = sqrt(0.5 * (a_square_roe + b_square_roe + a_star_roe) * norm_squared);
        vel_out = v_roe[0] * n[0] + v_roe[1] * n[1] + v_roe[2] * n[2];
    }
    TB_DEV void numflux_normal(int id, const double (&ul)[9], const double (&ur)[9], const double (&n)[3],
                               double (&f)[9]) const {
        switch (id) {
        case TRIXI_B200_FLUX_CENTRAL:
        case TRIXI_B200_FLUX_CENTRAL_MHD_POWELL: {
            double fl[9], fr[9];
            flux_normal(ul, n, fl);
            flux_normal(ur, n, fr);
#pragma unroll
            for (int v = 0; v < 9; ++v) f[v] = 0.5 * (fl[v] + fr[v]);
            break;
        }
        case TRIXI_B200_FLUX_LLF:
        case TRIXI_B200_FLUX_LLF_NAIVE:
        case TRIXI_B200_FLUX_LLF_MHD_POWELL:
        case TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL: {  // max_abs_speed(_naive) (:879-956)
            const bool naive = id == TRIXI_B200_FLUX_LLF_NAIVE || id == TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL;
            const double v_ll = (ul[1] * n[0] + ul[2] * n[1] + ul[3] * n[2]) / ul[0];
            const double v_rr = (ur[1] * n[0] + ur[2] * n[1] + ur[3] * n[2]) / ur[0];
            const double cf_ll = fast_wavespeed_normal(ul, n), cf_rr = fast_wavespeed_normal(ur, n);
            const double lam = naive ? fmax(fabs(v_ll), fabs(v_rr)) + fmax(cf_ll, cf_rr)
                                     : fmax(fabs(v_ll) + cf_ll, fabs(v_rr) + cf_rr);
            double fl[9], fr[9];
            flux_normal(ul, n, fl);
            flux_normal(ur, n, fr);
#pragma unroll
            for (int v = 0; v < 9; ++v) f[v] = 0.5 * (fl[v] + fr[v]) + (-0.5 * lam * (ur[v] - ul[v]));
            break;
        }
        case TRIXI_B200_FLUX_HINDENLANG_GASSNER:
        case TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL:
            flux_hindenlang_gassner_normal(ul, ur, n, f);
            break;
        case TRIXI_B200_FLUX_HLLE_MHD_POWELL: {  // min_max_speed_einfeldt (:1132-1166)
            const double vn_ll = (ul[1] * n[0] + ul[2] * n[1] + ul[3] * n[2]) / ul[0];
            const double vn_rr = (ur[1] * n[0] + ur[2] * n[1] + ur[3] * n[2]) / ur[0];
            const double cf_ll = fast_wavespeed_normal(ul, n), cf_rr = fast_wavespeed_normal(ur, n);
            double v_roe, cf_roe;
            fast_wavespeed_roe_normal(ul, ur, n, v_roe, cf_roe);
            const double lmin = fmin(vn_ll - cf_ll, v_roe - cf_roe), lmax = fmax(vn_rr + cf_rr, v_roe + cf_roe);
            if (lmin >= 0 && lmax >= 0) {
                flux_normal(ul, n, f);
            } else if (lmax <= 0 && lmin <= 0) {
                flux_normal(ur, n, f);
            } else {
                double fl[9], fr[9];
                flux_normal(ul, n, fl);
                flux_normal(ur, n, fr);
                const double inv = 1.0 / (lmax - lmin);
                const double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
#pragma unroll
                for (int v = 0; v < 9; ++v) f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
            }
            break;
        }
        default: nanfill(f);
        }
    }
    // not part of this build: boundary walls
    TB_DEV void slip_wall(const double (&u)[9], int o, int direction, double (&f)[9]) const { nanfill(f); }
    TB_DEV void slip_wall_normal(const double (&u)[9], const double (&n)[3], int direction, double (&f)[9]) const {
        nanfill(f);
    }
    TB_DEV void slip_wall_outward(const double (&u)[9], const double (&n)[3], double (&f)[9]) const { nanfill(f); }
    TB_DEV static void nanfill(double (&f)[9]) {
#pragma unroll
        for (int v = 0; v < 9; ++v) f[v] = nan("");
    }
};

// bc(u_inner, orientation, direction, x, t, surface_flux, equations) for TreeMesh (dg_3d.jl:757)
template <class EQ>
TB_DEV void boundary_flux(const EQ &eq, int bc, int ic, int surface_flux, const double (&u_inner)[EQ::NVARS],
                          int o, int direction, const double (&x)[EQ::NDIMS], double t,
                          double (&f)[EQ::NVARS]) {
    if (bc == TRIXI_B200_BC_DIRICHLET) {  // equations.jl:164-183
        double ub[EQ::NVARS];
        eq.initial_condition(ic, x, t, ub);
        if (direction % 2 == 0)
            eq.numflux(surface_flux, u_inner, ub, o, f);
        else
            eq.numflux(surface_flux, ub, u_inner, o, f);
    } else if (bc == TRIXI_B200_BC_SLIP_WALL) {
        eq.slip_wall(u_inner, o, direction, f);
    } else {
#pragma unroll
        for (int v = 0; v < EQ::NVARS; ++v) f[v] = nan("");
    }
}

// bc(u_inner, normal, direction, x, t, surface_flux, equations) for curved meshes (dgsem_structured/dg.jl:124-165)
template <class EQ>
TB_DEV void boundary_flux_normal(const EQ &eq, int bc, int ic, int surface_flux, const double (&u_inner)[EQ::NVARS],
                                 const double (&n)[EQ::NDIMS], int direction, const double (&x)[EQ::NDIMS], double t,
                                 double (&f)[EQ::NVARS]) {
    if (bc == TRIXI_B200_BC_DIRICHLET) {
        double ub[EQ::NVARS];
        eq.initial_condition(ic, x, t, ub);
        if (direction % 2 == 0)
            eq.numflux_normal(surface_flux, u_inner, ub, n, f);
        else
            eq.numflux_normal(surface_flux, ub, u_inner, n, f);
    } else if (bc == TRIXI_B200_BC_SLIP_WALL) {
        eq.slip_wall_normal(u_inner, n, direction, f);
    } else {
#pragma unroll
        for (int v = 0; v < EQ::NVARS; ++v) f[v] = nan("");
    }
}

}  // namespace tb
