// Pieces shared by the tuned flux_ranocha element kernels (TreeMesh p = 3, curved p = 3, curved p = 5): the hoisted
// node record, the logarithm for positive normal arguments, the two logarithmic means and the two-point flux in its
// rotated (Cartesian) and normal-direction (curved) forms.  Everything here has internal linkage: the kernels live in
// more than one translation unit.
#pragma once
#include <cstdint>

#include "tile_io.cuh"

namespace tb {

// Node record in the prim tile: rho, v1, v2, v3, 2 p, log(rho), log(rho) - log(p)
constexpr int kNP = 7;

// log(x) for positive, finite, normal x (fdlibm's __ieee754_log kernel: x = 2^k (1 + f), sqrt(1/2) <= 1 + f <
// sqrt(2), log(1 + f) = f - hfsq + s (hfsq + R(s^2)), s = f / (2 + f); < 1 ulp).  Everything else (zero,
// negative, subnormal, inf, NaN) takes libm's log out of line, so the special values propagate like the
// reference's.  The coefficients sit in the constant bank and are consumed as DFMA operands.
static __constant__ double kLogC[9] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                1.479819860511658591e-01, 6.93147180369123816490e-01, 1.90821492927058770002e-10};
static __device__ __noinline__ double log_special(double x) { return log(x); }
TB_DEV double log_pos(double x) {
    const int hi = __double2hiint(x);
    if (__builtin_expect((unsigned)(hi - 0x00100000) >= 0x7fe00000u, 0)) return log_special(x);
    const int hx = hi + (0x3ff00000 - 0x3fe6a09e);
    const int k = (hx >> 20) - 0x3ff;
    const double m = __hiloint2double((hx & 0x000fffff) + 0x3fe6a09e, __double2loint(x));
    const double f = m - 1.0;
    const double hfsq = (0.5 * f) * f;
    const double s = f * fast_rcp(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, kLogC[5], kLogC[3]), kLogC[1]);
    const double t2 = z * fma(w, fma(w, fma(w, kLogC[6], kLogC[4]), kLogC[2]), kLogC[0]);
    const double R = t2 + t1;
    const double dk = (double)k;
    return fma(dk, kLogC[7], (fma(s, hfsq + R, dk * kLogC[8]) - hfsq) + f);
}

// ln_mean(rho_ll, rho_rr) and 2 p_ll p_rr inv_ln_mean(rho_ll p_rr, rho_rr p_ll) (math.jl:198-250) from node records
// (rho, ., ., ., 2 p, log rho, log rho - log p): f^2 = ((x - y) / (x + y))^2 decides between the Ismail-Roe series
// and the hoisted logarithms, and the chosen operands go through ONE division each.
TB_DEV void ranocha_means(const double *L, const double *R, double &rho_mean, double &inv_rho_p_mean2) {
    const double rho_ll = L[0], p2_ll = L[4], rho_rr = R[0], p2_rr = R[4];
    {
        const double sum = rho_ll + rho_rr, dif = rho_rr - rho_ll;
        const double f = dif * rcp_1nr(sum), f2 = f * f;
        const bool series = f2 < 1.0e-4;
        const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        rho_mean = fast_div(series ? sum : dif, series ? poly : R[5] - L[5]);
    }
    {
        // x = 2 rho_ll p_rr, y = 2 rho_rr p_ll: p2_ll p2_rr inv_ln_mean(x, y) = 2 p_ll p_rr inv_ln_mean(x / 2, y / 2)
        const double x = rho_ll * p2_rr, y = rho_rr * p2_ll;
        const double sum = x + y, dif = y - x;
        const double f = dif * rcp_1nr(sum), f2 = f * f;
        const bool series = f2 < 1.0e-4;
        const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        // log(y / x) = (log rho_rr - log p_rr) - (log rho_ll - log p_ll)
        const double m = fast_div(series ? poly : R[6] - L[6], series ? sum : dif);
        inv_rho_p_mean2 = p2_ll * p2_rr * m;
    }
}

// 4 * flux_ranocha(u_ll, u_rr, orientation) (compressible_euler_3d.jl:746-793) on hoisted node records whose
// velocity components have been rotated so that slot 1 is the normal one: (rho, vn, vt1, vt2, 2 p, log rho,
// log rho - log p).  The output is rotated the same way and scaled by powers of two that the caller's D_split
// weights undo: g = (2 f_rho, 4 f_n, 4 f_t1, 4 f_t2, 4 f_E) -- the halves of the arithmetic means never get
// multiplied out, and the record carries 2 p so that the doubled pressure terms need no doubling either (exact:
// scaling by 2 commutes with rounding).  igm1 = 1 / (gamma - 1).
TB_DEV void ranocha_pair_rot(const double (&L)[kNP], const double (&R)[kNP], double igm1, double (&g)[5]) {
    const double p2_ll = L[4], p2_rr = R[4];
    double rho_mean, inv_rho_p_mean2;
    ranocha_means(L, R, rho_mean, inv_rho_p_mean2);
    const double sn = L[1] + R[1], st1 = L[2] + R[2], st2 = L[3] + R[3];  // 2 v_avg
    const double vs = L[1] * R[1] + L[2] * R[2] + L[3] * R[3];            // 2 velocity_square_avg
    const double f1 = rho_mean * sn;                                       // 2 f_rho
    g[0] = f1;
    g[1] = fma(f1, sn, p2_ll + p2_rr);                                      // 4 (f_rho v_avg + p_avg)
    g[2] = f1 * st1;
    g[3] = f1 * st2;
    g[4] = fma(f1, fma(inv_rho_p_mean2, igm1, vs), p2_ll * R[1] + p2_rr * L[1]);
}

constexpr int kNPC = 16, kNPC_STRIDE = 17;  // rho, v1, v2, v3, 2 p, log rho, log rho - log p, Ja^1, Ja^2, Ja^3

// 8 * flux_ranocha(u_ll, u_rr, 0.5 (ja_ll + ja_rr)) except for the density flux, which comes as 4 f_rho:
// n = ja_ll + ja_rr, f1 = rho_mean (v_ll . n + v_rr . n) = 4 f_rho, g_m = f1 (v_ll + v_rr) + (2 p_ll + 2 p_rr) n,
// g_E = f1 (2 velocity_square_avg + 2 inv_rho_p_mean / (gamma - 1)) + (2 p_ll v_rr . n + 2 p_rr v_ll . n)
TB_DEV void ranocha_pair_normal(const double *L, const double *R, int d, double igm1, double (&g)[5]) {
    double rho_mean, inv_rho_p_mean2;
    ranocha_means(L, R, rho_mean, inv_rho_p_mean2);
    const double *jl = L + 7 + 3 * d, *jr = R + 7 + 3 * d;
    const double n1 = jl[0] + jr[0], n2 = jl[1] + jr[1], n3 = jl[2] + jr[2];
    const double vn_ll = L[1] * n1 + L[2] * n2 + L[3] * n3, vn_rr = R[1] * n1 + R[2] * n2 + R[3] * n3;
    const double p2s = L[4] + R[4];
    const double vs = L[1] * R[1] + L[2] * R[2] + L[3] * R[3];
    const double f1 = rho_mean * (vn_ll + vn_rr);
    g[0] = f1;
    g[1] = fma(f1, L[1] + R[1], p2s * n1);
    g[2] = fma(f1, L[2] + R[2], p2s * n2);
    g[3] = fma(f1, L[3] + R[3], p2s * n3);
    g[4] = fma(f1, fma(inv_rho_p_mean2, igm1, vs), L[4] * vn_rr + R[4] * vn_ll);
}

// u_resident(P): must match between the launcher (shared memory size) and the kernel
TB_DEV_HOST bool tuned_u_resident(const KParams &P, bool with_surface) {
    const bool have_src = with_surface && P.source_terms != TRIXI_B200_SRC_NONE;
    const bool rk = P.mode != 0;
    // (the 3S* and SSP stage updates, modes 2 and 3, are not u += increment: they need u in the epilogue)
    return have_src || (rk && (P.mode != 1 || P.want_cfl || !P.rk_reduce_update || P.u_out != P.u));
}

}  // namespace tb
