/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C + OpenMP) of the reference's DGSEM hot path, stage by stage, used only
 *   - by tests/ as the parity oracle for the CUDA kernels,
 *   - by __graft_entry__.smoke() as the checker,
 *   - by bench.py's cpu_baseline / `--impl reference` leg (kind = "port").
 * Nothing under trixi.jl_b200/ may import, link or call this file.
 *
 * The reference (Trixi.jl, pure Julia) cannot run in the build container or on the GPU box (no
 * julia).  Parity is pinned instead against the reference's own golden L2/Linf vectors
 * (test/test_tree_3d_euler.jl, test/test_tree_2d_advection.jl, ...): tests/test_oracle_golden.py runs
 * this oracle through the full elixir configurations and reproduces them.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 * Arrays use the reference's layouts (column-major, variable fastest, 1-based int64 indices).
 * Compiled with -ffp-contract=fast to mirror `@muladd` (dg_3d.jl:5).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/trixi_b200.h"

#define MAXV 9

typedef struct {
    int nd, nv, id;
    double gamma, inv_gm1;
    double a[3];
    double c_h;
} eqn_t;

static eqn_t make_eqn(const trixi_b200_desc *d) {
    eqn_t e;
    memset(&e, 0, sizeof(e));
    e.nd = d->ndims;
    e.nv = d->nvars;
    e.id = d->equation;
    if (e.id == TRIXI_B200_EQ_ADVECTION_2D || e.id == TRIXI_B200_EQ_ADVECTION_3D) {
        e.a[0] = d->eq_params[0];
        e.a[1] = d->eq_params[1];
        e.a[2] = d->eq_params[2];
    } else {
        e.gamma = d->eq_params[0];
        e.inv_gm1 = d->eq_params[1];
        e.c_h = d->eq_params[2];
    }
    return e;
}

static inline int ipow(int b, int e) {
    int r = 1;
    for (int i = 0; i < e; ++i) r *= b;
    return r;
}

/* ---- src/auxiliary/math.jl --------------------------------------------------------------------- */
/* ln_mean math.jl:198-210 */
static inline double ln_mean(double x, double y) {
    const double epsilon_f2 = 1.0e-4;
    double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
    if (f2 < epsilon_f2) {
        /* @evalpoly(f2, 2, 2/3, 2/5, 2/7): Horner with muladd */
        double p = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        return (x + y) / p;
    } else {
        return (y - x) / log(y / x);
    }
}
/* inv_ln_mean math.jl:238-250 */
static inline double inv_ln_mean(double x, double y) {
    const double epsilon_f2 = 1.0e-4;
    double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
    if (f2 < epsilon_f2) {
        double p = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        return p / (x + y);
    } else {
        return log(y / x) / (y - x);
    }
}

/* ---- compressible Euler 2D/3D (compressible_euler_3d.jl, compressible_euler_2d.jl) -------------- */
/* cons2prim compressible_euler_3d.jl:1783-1793 / compressible_euler_2d.jl:2038-2047;
 * prim = (rho, v[0..nd-1], p) */
static inline void euler_cons2prim(const eqn_t *eq, const double *u, double *rho, double *v, double *p) {
    int nd = eq->nd;
    *rho = u[0];
    double kin = 0.0;
    for (int d = 0; d < nd; ++d) {
        v[d] = u[1 + d] / u[0];
        kin += u[1 + d] * v[d];
    }
    *p = (eq->gamma - 1) * (u[nd + 1] - 0.5 * kin);
}

/* flux(u, orientation, eq) compressible_euler_3d.jl:420-447 */
static inline void euler_flux(const eqn_t *eq, const double *u, int o, double *f) {
    int nd = eq->nd;
    double rho, v[3], p;
    euler_cons2prim(eq, u, &rho, v, &p);
    double rv = u[1 + o];
    f[0] = rv;
    for (int d = 0; d < nd; ++d) f[1 + d] = rv * v[d];
    f[1 + o] += p;
    f[nd + 1] = (u[nd + 1] + p) * v[o];
}

/* flux_ranocha compressible_euler_3d.jl:746-793 */
static inline void euler_flux_ranocha(const eqn_t *eq, const double *ul, const double *ur, int o, double *f) {
    int nd = eq->nd;
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double rho_mean = ln_mean(rho_ll, rho_rr);
    double inv_rho_p_mean = p_ll * p_rr * inv_ln_mean(rho_ll * p_rr, rho_rr * p_ll);
    double v_avg[3], vsq = 0.0;
    for (int d = 0; d < nd; ++d) {
        v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
        vsq += v_ll[d] * v_rr[d];
    }
    double p_avg = 0.5 * (p_ll + p_rr);
    double velocity_square_avg = 0.5 * vsq;
    double f1 = rho_mean * v_avg[o];
    f[0] = f1;
    for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d];
    f[1 + o] += p_avg;
    f[nd + 1] = f1 * (velocity_square_avg + inv_rho_p_mean * eq->inv_gm1) +
                0.5 * (p_ll * v_rr[o] + p_rr * v_ll[o]);
}

/* flux_shima_etal compressible_euler_3d.jl:473-510 */
static inline void euler_flux_shima(const eqn_t *eq, const double *ul, const double *ur, int o, double *f) {
    int nd = eq->nd;
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double rho_avg = 0.5 * (rho_ll + rho_rr), p_avg = 0.5 * (p_ll + p_rr);
    double v_avg[3], kin = 0.0;
    for (int d = 0; d < nd; ++d) {
        v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
        kin += v_ll[d] * v_rr[d];
    }
    double kin_avg = 0.5 * kin;
    double pv_avg = 0.5 * (p_ll * v_rr[o] + p_rr * v_ll[o]);
    double f1 = rho_avg * v_avg[o];
    f[0] = f1;
    for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d];
    f[1 + o] += p_avg;
    f[nd + 1] = p_avg * v_avg[o] * eq->inv_gm1 + f1 * kin_avg + pv_avg;
}

/* flux_kennedy_gruber compressible_euler_3d.jl:560-600 */
static inline void euler_flux_kennedy_gruber(const eqn_t *eq, const double *ul, const double *ur, int o,
                                             double *f) {
    int nd = eq->nd;
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double rho_avg = 0.5 * (rho_ll + rho_rr), p_avg = 0.5 * (p_ll + p_rr);
    double v_avg[3];
    for (int d = 0; d < nd; ++d) v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
    double e_avg = 0.5 * (ul[nd + 1] / rho_ll + ur[nd + 1] / rho_rr);
    double f1 = rho_avg * v_avg[o];
    f[0] = f1;
    for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d];
    f[1 + o] += p_avg;
    f[nd + 1] = (rho_avg * e_avg + p_avg) * v_avg[o];
}

/* flux_chandrashekar compressible_euler_3d.jl:639-690 */
static inline void euler_flux_chandrashekar(const eqn_t *eq, const double *ul, const double *ur, int o,
                                            double *f) {
    int nd = eq->nd;
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double beta_ll = 0.5 * rho_ll / p_ll, beta_rr = 0.5 * rho_rr / p_rr;
    double kl = 0.0, kr = 0.0, v_avg[3];
    for (int d = 0; d < nd; ++d) {
        kl += v_ll[d] * v_ll[d];
        kr += v_rr[d] * v_rr[d];
        v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
    }
    double specific_kin_ll = 0.5 * kl, specific_kin_rr = 0.5 * kr;
    double rho_avg = 0.5 * (rho_ll + rho_rr);
    double rho_mean = ln_mean(rho_ll, rho_rr);
    double beta_mean = ln_mean(beta_ll, beta_rr);
    double beta_avg = 0.5 * (beta_ll + beta_rr);
    double p_mean = 0.5 * rho_avg / beta_avg;
    double velocity_square_avg = specific_kin_ll + specific_kin_rr;
    double f1 = rho_mean * v_avg[o];
    f[0] = f1;
    for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d];
    f[1 + o] += p_mean;
    double s = f1 * 0.5 * (1 / (eq->gamma - 1) / beta_mean - velocity_square_avg);
    for (int d = 0; d < nd; ++d) s += f[1 + d] * v_avg[d];
    f[nd + 1] = s;
}

/* flux_chandrashekar(u_ll, u_rr, normal_direction) compressible_euler_3d.jl:693-733, compressible_euler_2d.jl:639-670 */
static inline void euler_flux_chandrashekar_normal(const eqn_t *eq, const double *ul, const double *ur, const double *n,
                                                   double *f) {
    int nd = eq->nd;
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double v_dot_n_ll = 0.0, v_dot_n_rr = 0.0;
    for (int d = 0; d < nd; ++d) {
        v_dot_n_ll += v_ll[d] * n[d];
        v_dot_n_rr += v_rr[d] * n[d];
    }
    double beta_ll = 0.5 * rho_ll / p_ll, beta_rr = 0.5 * rho_rr / p_rr;
    double kl = 0.0, kr = 0.0, v_avg[3];
    for (int d = 0; d < nd; ++d) {
        kl += v_ll[d] * v_ll[d];
        kr += v_rr[d] * v_rr[d];
        v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
    }
    double specific_kin_ll = 0.5 * kl, specific_kin_rr = 0.5 * kr;
    double rho_avg = 0.5 * (rho_ll + rho_rr);
    double rho_mean = ln_mean(rho_ll, rho_rr);
    double beta_mean = ln_mean(beta_ll, beta_rr);
    double beta_avg = 0.5 * (beta_ll + beta_rr);
    double p_mean = 0.5 * rho_avg / beta_avg;
    double velocity_square_avg = specific_kin_ll + specific_kin_rr;
    double f1 = rho_mean * 0.5 * (v_dot_n_ll + v_dot_n_rr);
    f[0] = f1;
    for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d] + p_mean * n[d];
    double s = f1 * 0.5 * (1 / (eq->gamma - 1) / beta_mean - velocity_square_avg);
    for (int d = 0; d < nd; ++d) s += f[1 + d] * v_avg[d];
    f[nd + 1] = s;
}

/* max_abs_speed_naive compressible_euler_3d.jl:1112-1133; max_abs_speed :1156-1177 */
static inline double euler_max_abs_speed(const eqn_t *eq, const double *ul, const double *ur, int o, int naive) {
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double c_ll = sqrt(eq->gamma * p_ll / rho_ll);
    double c_rr = sqrt(eq->gamma * p_rr / rho_rr);
    if (naive) return fmax(fabs(v_ll[o]), fabs(v_rr[o])) + fmax(c_ll, c_rr);
    return fmax(fabs(v_ll[o]) + c_ll, fabs(v_rr[o]) + c_rr);
}

/* min_max_speed_davis compressible_euler_3d.jl:1240-1261; min_max_speed_naive :1202-1220 */
static inline void euler_min_max_speed(const eqn_t *eq, const double *ul, const double *ur, int o, int naive,
                                       double *lmin, double *lmax) {
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double c_ll = sqrt(eq->gamma * p_ll / rho_ll);
    double c_rr = sqrt(eq->gamma * p_rr / rho_rr);
    if (naive) {
        *lmin = v_ll[o] - c_ll;
        *lmax = v_rr[o] + c_rr;
    } else {
        *lmin = fmin(v_ll[o] - c_ll, v_rr[o] - c_rr);
        *lmax = fmax(v_ll[o] + c_ll, v_rr[o] + c_rr);
    }
}

static inline double vec_norm(int nd, const double *n);
/* min_max_speed_einfeldt (compressible_euler_3d.jl:1662-1707 with an orientation, :1723-1771 along a normal direction;
 * compressible_euler_2d.jl:1925-1966, :1982-2025): Roe averages with the positivity-preserving bound of Einfeldt et al.
 * n == NULL: orientation o (0-based) */
static inline void euler_min_max_speed_einfeldt(const eqn_t *eq, const double *ul, const double *ur, int o, const double *n,
                                                double *lmin, double *lmax) {
    int nd = eq->nd;
    double rho_ll, v_ll[3] = {0, 0, 0}, p_ll, rho_rr, v_rr[3] = {0, 0, 0}, p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double vn_ll, vn_rr, norm_ = 1.0;
    if (n) {
        vn_ll = 0.0;
        vn_rr = 0.0;
        for (int d = 0; d < nd; ++d) {
            vn_ll += v_ll[d] * n[d];
            vn_rr += v_rr[d] * n[d];
        }
        norm_ = vec_norm(nd, n);
    } else {
        vn_ll = v_ll[o];
        vn_rr = v_rr[o];
    }
    double H_ll = (ul[nd + 1] + p_ll) / rho_ll, H_rr = (ur[nd + 1] + p_rr) / rho_rr;
    double c_ll = sqrt(eq->gamma * p_ll / rho_ll), c_rr = sqrt(eq->gamma * p_rr / rho_rr);
    if (n) {
        c_ll = c_ll * norm_;
        c_rr = c_rr * norm_;
    }
    double sqrt_rho_ll = sqrt(rho_ll), sqrt_rho_rr = sqrt(rho_rr);
    double inv_sum_sqrt_rho = 1.0 / (sqrt_rho_ll + sqrt_rho_rr);
    double v_roe[3] = {0, 0, 0}, v_roe_mag = 0.0, vn_roe = 0.0;
    for (int d = 0; d < nd; ++d) {
        v_roe[d] = (sqrt_rho_ll * v_ll[d] + sqrt_rho_rr * v_rr[d]) * inv_sum_sqrt_rho;
        v_roe_mag += v_roe[d] * v_roe[d];
    }
    if (n)
        for (int d = 0; d < nd; ++d) vn_roe += v_roe[d] * n[d];
    else
        vn_roe = v_roe[o];
    double H_roe = (sqrt_rho_ll * H_ll + sqrt_rho_rr * H_rr) * inv_sum_sqrt_rho;
    double c_roe = sqrt((eq->gamma - 1) * (H_roe - 0.5 * v_roe_mag));
    if (n) c_roe = c_roe * norm_;
    double beta = sqrt(0.5 * (eq->gamma - 1) / eq->gamma);
    *lmin = fmin(fmin(vn_roe - c_roe, vn_ll - beta * c_ll), 0.0);
    *lmax = fmax(fmax(vn_roe + c_roe, vn_rr + beta * c_rr), 0.0);
}

static inline void euler_flux(const eqn_t *eq, const double *u, int o, double *f);
static inline void euler_flux_normal(const eqn_t *eq, const double *u, const double *n, double *f);
/* flux_hllc (compressible_euler_3d.jl:1423-1541 with an orientation, :1543-1665 along a normal direction;
 * compressible_euler_2d.jl:1720-1925).  n == NULL: orientation o (0-based) */
static inline void euler_flux_hllc(const eqn_t *eq, const double *ul, const double *ur, int o, const double *n,
                                   double *f) {
    int nd = eq->nd, nv = eq->nv;
    double rho_ll = ul[0], rho_rr = ur[0], v_ll[3] = {0, 0, 0}, v_rr[3] = {0, 0, 0};
    double vsq_ll = 0.0, vsq_rr = 0.0;
    for (int d = 0; d < nd; ++d) {
        v_ll[d] = ul[1 + d] / rho_ll;
        v_rr[d] = ur[1 + d] / rho_rr;
        vsq_ll += v_ll[d] * v_ll[d];
        vsq_rr += v_rr[d] * v_rr[d];
    }
    double e_ll = ul[nd + 1] / rho_ll, e_rr = ur[nd + 1] / rho_rr;
    double p_ll, p_rr;
    if (n) { /* cons2prim */
        double rl, vl3[3], rr_, vr3[3];
        euler_cons2prim(eq, ul, &rl, vl3, &p_ll);
        euler_cons2prim(eq, ur, &rr_, vr3, &p_rr);
    } else {
        p_ll = (eq->gamma - 1) * (ul[nd + 1] - 0.5 * rho_ll * vsq_ll);
        p_rr = (eq->gamma - 1) * (ur[nd + 1] - 0.5 * rho_rr * vsq_rr);
    }
    double norm_ = 1.0, norm_sq = 1.0, inv_norm_sq = 1.0, vel_L, vel_R;
    if (n) {
        vel_L = 0.0;
        vel_R = 0.0;
        for (int d = 0; d < nd; ++d) {
            vel_L += v_ll[d] * n[d];
            vel_R += v_rr[d] * n[d];
        }
        norm_ = vec_norm(nd, n);
        norm_sq = norm_ * norm_;
        inv_norm_sq = 1.0 / norm_sq;
    } else {
        vel_L = v_ll[o];
        vel_R = v_rr[o];
    }
    double c_ll = sqrt(eq->gamma * p_ll / rho_ll), c_rr = sqrt(eq->gamma * p_rr / rho_rr);
    if (n) {
        c_ll = c_ll * norm_;
        c_rr = c_rr * norm_;
    }
    double f_ll[MAXV], f_rr[MAXV];
    if (n) {
        euler_flux_normal(eq, ul, n, f_ll);
        euler_flux_normal(eq, ur, n, f_rr);
    } else {
        euler_flux(eq, ul, o, f_ll);
        euler_flux(eq, ur, o, f_rr);
    }
    double sqrt_rho_ll = sqrt(rho_ll), sqrt_rho_rr = sqrt(rho_rr), sum_sqrt_rho = sqrt_rho_ll + sqrt_rho_rr;
    double vel_roe, vel_roe_mag = 0.0;
    if (n) {
        double v_roe[3];
        vel_roe = 0.0;
        for (int d = 0; d < nd; ++d) {
            v_roe[d] = (sqrt_rho_ll * v_ll[d] + sqrt_rho_rr * v_rr[d]) / sum_sqrt_rho;
            vel_roe += v_roe[d] * n[d];
            vel_roe_mag += v_roe[d] * v_roe[d];
        }
    } else {
        vel_roe = (sqrt_rho_ll * vel_L + sqrt_rho_rr * vel_R) / sum_sqrt_rho;
        for (int d = 0; d < nd; ++d) {
            double w = sqrt_rho_ll * v_ll[d] + sqrt_rho_rr * v_rr[d];
            vel_roe_mag += w * w;
        }
        vel_roe_mag = vel_roe_mag / (sum_sqrt_rho * sum_sqrt_rho);
    }
    double H_ll = (ul[nd + 1] + p_ll) / rho_ll, H_rr = (ur[nd + 1] + p_rr) / rho_rr;
    double H_roe = (sqrt_rho_ll * H_ll + sqrt_rho_rr * H_rr) / sum_sqrt_rho;
    double c_roe = sqrt((eq->gamma - 1) * (H_roe - 0.5 * vel_roe_mag));
    if (n) c_roe = c_roe * norm_;
    double Ssl = fmin(vel_L - c_ll, vel_roe - c_roe), Ssr = fmax(vel_R + c_rr, vel_roe + c_roe);
    double sMu_L = Ssl - vel_L, sMu_R = Ssr - vel_R;
    if (Ssl >= 0) {
        for (int v = 0; v < nv; ++v) f[v] = f_ll[v];
        return;
    }
    if (Ssr <= 0) {
        for (int v = 0; v < nv; ++v) f[v] = f_rr[v];
        return;
    }
    double SStar = n ? (rho_ll * vel_L * sMu_L - rho_rr * vel_R * sMu_R + (p_rr - p_ll) * norm_sq) /
                           (rho_ll * sMu_L - rho_rr * sMu_R)
                     : (p_rr - p_ll + rho_ll * vel_L * sMu_L - rho_rr * vel_R * sMu_R) / (rho_ll * sMu_L - rho_rr * sMu_R);
    int left = Ssl <= 0 && 0 <= SStar;
    const double *us = left ? ul : ur, *fs = left ? f_ll : f_rr, *vs = left ? v_ll : v_rr;
    double rho_s = left ? rho_ll : rho_rr, sMu = left ? sMu_L : sMu_R, Ss = left ? Ssl : Ssr;
    double vel_s = left ? vel_L : vel_R, e_s = left ? e_ll : e_rr, p_s = left ? p_ll : p_rr;
    double densStar = rho_s * sMu / (Ss - SStar);
    double UStar[MAXV];
    UStar[0] = densStar;
    if (n) {
        double enerStar = e_s + (SStar - vel_s) * (SStar * inv_norm_sq + p_s / (rho_s * sMu));
        for (int d = 0; d < nd; ++d) UStar[1 + d] = densStar * (vs[d] + (SStar - vel_s) * n[d] * inv_norm_sq);
        UStar[nd + 1] = densStar * enerStar;
    } else {
        double enerStar = e_s + (SStar - vel_s) * (SStar + p_s / (rho_s * sMu));
        for (int d = 0; d < nd; ++d) UStar[1 + d] = densStar * (d == o ? SStar : vs[d]);
        UStar[nd + 1] = densStar * enerStar;
    }
    for (int v = 0; v < nv; ++v) f[v] = fs[v] + Ss * (UStar[v] - us[v]);
}

/* ---- ideal GLM-MHD 3D (ideal_glm_mhd_3d.jl) ------------------------------------------------------------ */
/* cons2prim :1231-1243: (rho, v1, v2, v3, p, B1, B2, B3, psi) */
static inline void mhd_cons2prim(const eqn_t *eq, const double *u, double *prim) {
    double rho = u[0];
    double v1 = u[1] / rho, v2 = u[2] / rho, v3 = u[3] / rho;
    double p = (eq->gamma - 1) *
               (u[4] - 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3 + u[5] * u[5] + u[6] * u[6] + u[7] * u[7] + u[8] * u[8]));
    prim[0] = rho;
    prim[1] = v1;
    prim[2] = v2;
    prim[3] = v3;
    prim[4] = p;
    prim[5] = u[5];
    prim[6] = u[6];
    prim[7] = u[7];
    prim[8] = u[8];
}

/* flux(u, orientation) :187-234 */
static inline void mhd_flux(const eqn_t *eq, const double *u, int o, double *f) {
    double rho = u[0], psi = u[8];
    double v[3] = {u[1] / rho, u[2] / rho, u[3] / rho};
    const double *B = u + 5;
    double kin_en = 0.5 * (u[1] * v[0] + u[2] * v[1] + u[3] * v[2]);
    double mag_en = 0.5 * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
    double p_over_gamma_minus_one = (u[4] - kin_en - mag_en - 0.5 * psi * psi);
    double p = (eq->gamma - 1) * p_over_gamma_minus_one;
    double rv = u[1 + o];
    f[0] = rv;
    for (int d = 0; d < 3; ++d) f[1 + d] = rv * v[d] - B[o] * B[d];
    f[1 + o] = rv * v[o] + p + mag_en - B[o] * B[o];
    f[4] = (kin_en + eq->gamma * p_over_gamma_minus_one + 2 * mag_en) * v[o] -
           B[o] * (v[0] * B[0] + v[1] * B[1] + v[2] * B[2]) + eq->c_h * psi * B[o];
    for (int d = 0; d < 3; ++d) f[5 + d] = v[o] * B[d] - v[d] * B[o];
    f[5 + o] = eq->c_h * psi;
    f[8] = eq->c_h * B[o];
}

/* flux_nonconservative_powell(u_ll, u_rr, orientation) :295-340 */
static inline void mhd_noncons_powell(const eqn_t *eq, const double *ul, const double *ur, int o, double *f) {
    (void)eq;
    double v_ll[3] = {ul[1] / ul[0], ul[2] / ul[0], ul[3] / ul[0]};
    const double *B_ll = ul + 5;
    double psi_ll = ul[8], psi_rr = ur[8];
    double v_dot_B_ll = v_ll[0] * B_ll[0] + v_ll[1] * B_ll[1] + v_ll[2] * B_ll[2];
    double Bn_rr = ur[5 + o];
    f[0] = 0.0;
    for (int d = 0; d < 3; ++d) f[1 + d] = B_ll[d] * Bn_rr;
    f[4] = v_dot_B_ll * Bn_rr + v_ll[o] * psi_ll * psi_rr;
    for (int d = 0; d < 3; ++d) f[5 + d] = v_ll[d] * Bn_rr;
    f[8] = v_ll[o] * psi_rr;
}

/* flux_hindenlang_gassner(u_ll, u_rr, orientation) :680-779 */
static inline void mhd_flux_hindenlang_gassner(const eqn_t *eq, const double *ul, const double *ur, int o,
                                               double *f) {
    double L[9], R[9];
    mhd_cons2prim(eq, ul, L);
    mhd_cons2prim(eq, ur, R);
    double rho_ll = L[0], p_ll = L[4], psi_ll = L[8], rho_rr = R[0], p_rr = R[4], psi_rr = R[8];
    const double *v_ll = L + 1, *v_rr = R + 1, *B_ll = L + 5, *B_rr = R + 5;
    double rho_mean = ln_mean(rho_ll, rho_rr);
    double inv_rho_p_mean = p_ll * p_rr * inv_ln_mean(rho_ll * p_rr, rho_rr * p_ll);
    double v_avg[3];
    for (int d = 0; d < 3; ++d) v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
    double p_avg = 0.5 * (p_ll + p_rr), psi_avg = 0.5 * (psi_ll + psi_rr);
    double velocity_square_avg = 0.5 * (v_ll[0] * v_rr[0] + v_ll[1] * v_rr[1] + v_ll[2] * v_rr[2]);
    double magnetic_square_avg = 0.5 * (B_ll[0] * B_rr[0] + B_ll[1] * B_rr[1] + B_ll[2] * B_rr[2]);
    double f1 = rho_mean * v_avg[o];
    f[0] = f1;
    for (int d = 0; d < 3; ++d) f[1 + d] = f1 * v_avg[d] - 0.5 * (B_ll[o] * B_rr[d] + B_rr[o] * B_ll[d]);
    f[1 + o] = f1 * v_avg[o] + p_avg + magnetic_square_avg - 0.5 * (B_ll[o] * B_rr[o] + B_rr[o] * B_ll[o]);
    for (int d = 0; d < 3; ++d)
        f[5 + d] = 0.5 * (v_ll[o] * B_ll[d] - v_ll[d] * B_ll[o] + v_rr[o] * B_rr[d] - v_rr[d] * B_rr[o]);
    f[5 + o] = eq->c_h * psi_avg;
    f[8] = eq->c_h * 0.5 * (B_ll[o] + B_rr[o]);
    /* energy flux: the reference lists the two transverse directions in ascending order */
    int t1 = o == 0 ? 1 : 0, t2 = o == 2 ? 1 : 2;
    f[4] = f1 * (velocity_square_avg + inv_rho_p_mean * eq->inv_gm1) +
           0.5 * (+p_ll * v_rr[o] + p_rr * v_ll[o] + (v_ll[o] * B_ll[t1] * B_rr[t1] + v_rr[o] * B_rr[t1] * B_ll[t1]) +
                  (v_ll[o] * B_ll[t2] * B_rr[t2] + v_rr[o] * B_rr[t2] * B_ll[t2]) -
                  (v_ll[t1] * B_ll[o] * B_rr[t1] + v_rr[t1] * B_rr[o] * B_ll[t1]) -
                  (v_ll[t2] * B_ll[o] * B_rr[t2] + v_rr[t2] * B_rr[o] * B_ll[t2]) +
                  eq->c_h * (B_ll[o] * psi_rr + B_rr[o] * psi_ll));
}

/* calc_fast_wavespeed(cons, orientation) :1350-1376 */
static inline double mhd_fast_wavespeed(const eqn_t *eq, const double *u, int o) {
    double rho = u[0], psi = u[8];
    double v1 = u[1] / rho, v2 = u[2] / rho, v3 = u[3] / rho;
    double kin_en = 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3);
    double mag_en = 0.5 * (u[5] * u[5] + u[6] * u[6] + u[7] * u[7]);
    double p = (eq->gamma - 1) * (u[4] - kin_en - mag_en - 0.5 * psi * psi);
    double a_square = eq->gamma * p / rho;
    double sqrt_rho = sqrt(rho);
    double b[3] = {u[5] / sqrt_rho, u[6] / sqrt_rho, u[7] / sqrt_rho};
    double b_square = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    double s = a_square + b_square;
    return sqrt(0.5 * s + 0.5 * sqrt(s * s - 4 * a_square * b[o] * b[o]));
}

/* calc_fast_wavespeed_roe(u_ll, u_rr, orientation) :1415-1491 (Cargo & Gallice Roe averages) */
static inline void mhd_fast_wavespeed_roe(const eqn_t *eq, const double *ul, const double *ur, int o, double *vel_out,
                                          double *c_f) {
    double rho_ll = ul[0], rho_rr = ur[0];
    double v_ll[3] = {ul[1] / rho_ll, ul[2] / rho_ll, ul[3] / rho_ll};
    double v_rr[3] = {ur[1] / rho_rr, ur[2] / rho_rr, ur[3] / rho_rr};
    const double *B_ll = ul + 5, *B_rr = ur + 5;
    double kin_en_ll = 0.5 * (ul[1] * v_ll[0] + ul[2] * v_ll[1] + ul[3] * v_ll[2]);
    double mag_norm_ll = B_ll[0] * B_ll[0] + B_ll[1] * B_ll[1] + B_ll[2] * B_ll[2];
    double p_ll = (eq->gamma - 1) * (ul[4] - kin_en_ll - 0.5 * mag_norm_ll - 0.5 * ul[8] * ul[8]);
    double kin_en_rr = 0.5 * (ur[1] * v_rr[0] + ur[2] * v_rr[1] + ur[3] * v_rr[2]);
    double mag_norm_rr = B_rr[0] * B_rr[0] + B_rr[1] * B_rr[1] + B_rr[2] * B_rr[2];
    double p_rr = (eq->gamma - 1) * (ur[4] - kin_en_rr - 0.5 * mag_norm_rr - 0.5 * ur[8] * ur[8]);
    double p_total_ll = p_ll + 0.5 * mag_norm_ll, p_total_rr = p_rr + 0.5 * mag_norm_rr;
    double sqrt_rho_ll = sqrt(rho_ll), sqrt_rho_rr = sqrt(rho_rr);
    double inv_sqrt_rho_add = 1 / (sqrt_rho_ll + sqrt_rho_rr), inv_sqrt_rho_prod = 1 / (sqrt_rho_ll * sqrt_rho_rr);
    double rho_ll_roe = sqrt_rho_ll * inv_sqrt_rho_add, rho_rr_roe = sqrt_rho_rr * inv_sqrt_rho_add;
    double v_roe[3], B_roe[3];
    for (int d = 0; d < 3; ++d) {
        v_roe[d] = v_ll[d] * rho_ll_roe + v_rr[d] * rho_rr_roe;
        B_roe[d] = B_ll[d] * rho_ll_roe + B_rr[d] * rho_rr_roe;
    }
    double H_ll = (ul[4] + p_total_ll) / rho_ll, H_rr = (ur[4] + p_total_rr) / rho_rr;
    double H_roe = H_ll * rho_ll_roe + H_rr * rho_rr_roe;
    double dB0 = B_ll[0] - B_rr[0], dB1 = B_ll[1] - B_rr[1], dB2 = B_ll[2] - B_rr[2];
    double X = 0.5 * (dB0 * dB0 + dB1 * dB1 + dB2 * dB2) * (inv_sqrt_rho_add * inv_sqrt_rho_add);
    double b_square_roe = (B_roe[0] * B_roe[0] + B_roe[1] * B_roe[1] + B_roe[2] * B_roe[2]) * inv_sqrt_rho_prod;
    double a_square_roe = ((2 - eq->gamma) * X +
                           (eq->gamma - 1) * (H_roe - 0.5 * (v_roe[0] * v_roe[0] + v_roe[1] * v_roe[1] + v_roe[2] * v_roe[2]) -
                                              b_square_roe));
    double c_a_roe = B_roe[o] * B_roe[o] * inv_sqrt_rho_prod;
    double s = a_square_roe + b_square_roe;
    double a_star_roe = sqrt(s * s - 4 * a_square_roe * c_a_roe);
    *c_f = sqrt(0.5 * (a_square_roe + b_square_roe + a_star_roe));
    *vel_out = v_roe[o];
}

/* flux(u, normal_direction) :236-275 */
static inline void mhd_flux_normal(const eqn_t *eq, const double *u, const double *n, double *f) {
    double rho = u[0], psi = u[8];
    double v1 = u[1] / rho, v2 = u[2] / rho, v3 = u[3] / rho;
    double B1 = u[5], B2 = u[6], B3 = u[7];
    double kin_en = 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3);
    double mag_en = 0.5 * (B1 * B1 + B2 * B2 + B3 * B3);
    double p_over_gamma_minus_one = (u[4] - kin_en - mag_en - 0.5 * psi * psi);
    double p = (eq->gamma - 1) * p_over_gamma_minus_one;
    double v_normal = v1 * n[0] + v2 * n[1] + v3 * n[2];
    double B_normal = B1 * n[0] + B2 * n[1] + B3 * n[2];
    double rho_v_normal = rho * v_normal;
    f[0] = rho_v_normal;
    f[1] = rho_v_normal * v1 - B1 * B_normal + (p + mag_en) * n[0];
    f[2] = rho_v_normal * v2 - B2 * B_normal + (p + mag_en) * n[1];
    f[3] = rho_v_normal * v3 - B3 * B_normal + (p + mag_en) * n[2];
    f[4] = ((kin_en + eq->gamma * p_over_gamma_minus_one + 2 * mag_en) * v_normal - B_normal * (v1 * B1 + v2 * B2 + v3 * B3) +
            eq->c_h * psi * B_normal);
    f[5] = (eq->c_h * psi * n[0] + (v2 * B1 - v1 * B2) * n[1] + (v3 * B1 - v1 * B3) * n[2]);
    f[6] = ((v1 * B2 - v2 * B1) * n[0] + eq->c_h * psi * n[1] + (v3 * B2 - v2 * B3) * n[2]);
    f[7] = ((v1 * B3 - v3 * B1) * n[0] + (v2 * B3 - v3 * B2) * n[1] + eq->c_h * psi * n[2]);
    f[8] = eq->c_h * B_normal;
}

/* flux_nonconservative_powell(u_ll, u_rr, normal_direction) :342-374 */
static inline void mhd_noncons_powell_normal(const eqn_t *eq, const double *ul, const double *ur, const double *n,
                                             double *f) {
    (void)eq;
    double v1_ll = ul[1] / ul[0], v2_ll = ul[2] / ul[0], v3_ll = ul[3] / ul[0];
    double B1_ll = ul[5], B2_ll = ul[6], B3_ll = ul[7], psi_ll = ul[8], psi_rr = ur[8];
    double v_dot_B_ll = v1_ll * B1_ll + v2_ll * B2_ll + v3_ll * B3_ll;
    double v_dot_n_ll = v1_ll * n[0] + v2_ll * n[1] + v3_ll * n[2];
    double B_dot_n_rr = ur[5] * n[0] + ur[6] * n[1] + ur[7] * n[2];
    f[0] = 0.0;
    f[1] = B1_ll * B_dot_n_rr;
    f[2] = B2_ll * B_dot_n_rr;
    f[3] = B3_ll * B_dot_n_rr;
    f[4] = v_dot_B_ll * B_dot_n_rr + v_dot_n_ll * psi_ll * psi_rr;
    f[5] = v1_ll * B_dot_n_rr;
    f[6] = v2_ll * B_dot_n_rr;
    f[7] = v3_ll * B_dot_n_rr;
    f[8] = v_dot_n_ll * psi_rr;
}

/* flux_hindenlang_gassner(u_ll, u_rr, normal_direction) :781-855 */
static inline void mhd_flux_hindenlang_gassner_normal(const eqn_t *eq, const double *ul, const double *ur, const double *n,
                                                      double *f) {
    double L[9], R[9];
    mhd_cons2prim(eq, ul, L);
    mhd_cons2prim(eq, ur, R);
    double rho_ll = L[0], v1_ll = L[1], v2_ll = L[2], v3_ll = L[3], p_ll = L[4], B1_ll = L[5], B2_ll = L[6], B3_ll = L[7],
           psi_ll = L[8];
    double rho_rr = R[0], v1_rr = R[1], v2_rr = R[2], v3_rr = R[3], p_rr = R[4], B1_rr = R[5], B2_rr = R[6], B3_rr = R[7],
           psi_rr = R[8];
    double v_dot_n_ll = v1_ll * n[0] + v2_ll * n[1] + v3_ll * n[2];
    double v_dot_n_rr = v1_rr * n[0] + v2_rr * n[1] + v3_rr * n[2];
    double B_dot_n_ll = B1_ll * n[0] + B2_ll * n[1] + B3_ll * n[2];
    double B_dot_n_rr = B1_rr * n[0] + B2_rr * n[1] + B3_rr * n[2];
    double rho_mean = ln_mean(rho_ll, rho_rr);
    double inv_rho_p_mean = p_ll * p_rr * inv_ln_mean(rho_ll * p_rr, rho_rr * p_ll);
    double v1_avg = 0.5 * (v1_ll + v1_rr), v2_avg = 0.5 * (v2_ll + v2_rr), v3_avg = 0.5 * (v3_ll + v3_rr);
    double p_avg = 0.5 * (p_ll + p_rr), psi_avg = 0.5 * (psi_ll + psi_rr);
    double velocity_square_avg = 0.5 * (v1_ll * v1_rr + v2_ll * v2_rr + v3_ll * v3_rr);
    double magnetic_square_avg = 0.5 * (B1_ll * B1_rr + B2_ll * B2_rr + B3_ll * B3_rr);
    double f1 = rho_mean * 0.5 * (v_dot_n_ll + v_dot_n_rr);
    f[0] = f1;
    f[1] = (f1 * v1_avg + (p_avg + magnetic_square_avg) * n[0] - 0.5 * (B_dot_n_ll * B1_rr + B_dot_n_rr * B1_ll));
    f[2] = (f1 * v2_avg + (p_avg + magnetic_square_avg) * n[1] - 0.5 * (B_dot_n_ll * B2_rr + B_dot_n_rr * B2_ll));
    f[3] = (f1 * v3_avg + (p_avg + magnetic_square_avg) * n[2] - 0.5 * (B_dot_n_ll * B3_rr + B_dot_n_rr * B3_ll));
    f[5] = (eq->c_h * psi_avg * n[0] +
            0.5 * (v_dot_n_ll * B1_ll - v1_ll * B_dot_n_ll + v_dot_n_rr * B1_rr - v1_rr * B_dot_n_rr));
    f[6] = (eq->c_h * psi_avg * n[1] +
            0.5 * (v_dot_n_ll * B2_ll - v2_ll * B_dot_n_ll + v_dot_n_rr * B2_rr - v2_rr * B_dot_n_rr));
    f[7] = (eq->c_h * psi_avg * n[2] +
            0.5 * (v_dot_n_ll * B3_ll - v3_ll * B_dot_n_ll + v_dot_n_rr * B3_rr - v3_rr * B_dot_n_rr));
    f[8] = eq->c_h * 0.5 * (B_dot_n_ll + B_dot_n_rr);
    f[4] = (f1 * (velocity_square_avg + inv_rho_p_mean * eq->inv_gm1) +
            0.5 * (+p_ll * v_dot_n_rr + p_rr * v_dot_n_ll + (v_dot_n_ll * B1_ll * B1_rr + v_dot_n_rr * B1_rr * B1_ll) +
                   (v_dot_n_ll * B2_ll * B2_rr + v_dot_n_rr * B2_rr * B2_ll) +
                   (v_dot_n_ll * B3_ll * B3_rr + v_dot_n_rr * B3_rr * B3_ll) -
                   (v1_ll * B_dot_n_ll * B1_rr + v1_rr * B_dot_n_rr * B1_ll) -
                   (v2_ll * B_dot_n_ll * B2_rr + v2_rr * B_dot_n_rr * B2_ll) -
                   (v3_ll * B_dot_n_ll * B3_rr + v3_rr * B_dot_n_rr * B3_ll) +
                   eq->c_h * (B_dot_n_ll * psi_rr + B_dot_n_rr * psi_ll)));
}

/* calc_fast_wavespeed(cons, normal_direction) :1378-1404 */
static inline double mhd_fast_wavespeed_normal(const eqn_t *eq, const double *u, const double *n) {
    double rho = u[0], psi = u[8];
    double v1 = u[1] / rho, v2 = u[2] / rho, v3 = u[3] / rho;
    double kin_en = 0.5 * (u[1] * v1 + u[2] * v2 + u[3] * v3);
    double mag_en = 0.5 * (u[5] * u[5] + u[6] * u[6] + u[7] * u[7]);
    double p = (eq->gamma - 1) * (u[4] - kin_en - mag_en - 0.5 * psi * psi);
    double a_square = eq->gamma * p / rho;
    double sqrt_rho = sqrt(rho);
    double b1 = u[5] / sqrt_rho, b2 = u[6] / sqrt_rho, b3 = u[7] / sqrt_rho;
    double b_square = b1 * b1 + b2 * b2 + b3 * b3;
    double norm_squared = (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    double bn = (b1 * n[0] + b2 * n[1] + b3 * n[2]);
    double b_dot_n_squared = bn * bn / norm_squared;
    double s = a_square + b_square;
    return sqrt((0.5 * s + 0.5 * sqrt(s * s - 4 * a_square * b_dot_n_squared)) * norm_squared);
}

/* calc_fast_wavespeed_roe(u_ll, u_rr, normal_direction) :1493-1565 */
static inline void mhd_fast_wavespeed_roe_normal(const eqn_t *eq, const double *ul, const double *ur, const double *n,
                                                 double *vel_out, double *c_f) {
    double rho_ll = ul[0], rho_rr = ur[0];
    double v_ll[3] = {ul[1] / rho_ll, ul[2] / rho_ll, ul[3] / rho_ll};
    double v_rr[3] = {ur[1] / rho_rr, ur[2] / rho_rr, ur[3] / rho_rr};
    const double *B_ll = ul + 5, *B_rr = ur + 5;
    double kin_en_ll = 0.5 * (ul[1] * v_ll[0] + ul[2] * v_ll[1] + ul[3] * v_ll[2]);
    double mag_norm_ll = B_ll[0] * B_ll[0] + B_ll[1] * B_ll[1] + B_ll[2] * B_ll[2];
    double p_ll = (eq->gamma - 1) * (ul[4] - kin_en_ll - 0.5 * mag_norm_ll - 0.5 * ul[8] * ul[8]);
    double kin_en_rr = 0.5 * (ur[1] * v_rr[0] + ur[2] * v_rr[1] + ur[3] * v_rr[2]);
    double mag_norm_rr = B_rr[0] * B_rr[0] + B_rr[1] * B_rr[1] + B_rr[2] * B_rr[2];
    double p_rr = (eq->gamma - 1) * (ur[4] - kin_en_rr - 0.5 * mag_norm_rr - 0.5 * ur[8] * ur[8]);
    double p_total_ll = p_ll + 0.5 * mag_norm_ll, p_total_rr = p_rr + 0.5 * mag_norm_rr;
    double sqrt_rho_ll = sqrt(rho_ll), sqrt_rho_rr = sqrt(rho_rr);
    double inv_sqrt_rho_add = 1 / (sqrt_rho_ll + sqrt_rho_rr), inv_sqrt_rho_prod = 1 / (sqrt_rho_ll * sqrt_rho_rr);
    double rho_ll_roe = sqrt_rho_ll * inv_sqrt_rho_add, rho_rr_roe = sqrt_rho_rr * inv_sqrt_rho_add;
    double v_roe[3], B_roe[3];
    for (int d = 0; d < 3; ++d) {
        v_roe[d] = v_ll[d] * rho_ll_roe + v_rr[d] * rho_rr_roe;
        B_roe[d] = B_ll[d] * rho_ll_roe + B_rr[d] * rho_rr_roe;
    }
    double H_ll = (ul[4] + p_total_ll) / rho_ll, H_rr = (ur[4] + p_total_rr) / rho_rr;
    double H_roe = H_ll * rho_ll_roe + H_rr * rho_rr_roe;
    double dB0 = B_ll[0] - B_rr[0], dB1 = B_ll[1] - B_rr[1], dB2 = B_ll[2] - B_rr[2];
    double X = 0.5 * (dB0 * dB0 + dB1 * dB1 + dB2 * dB2) * (inv_sqrt_rho_add * inv_sqrt_rho_add);
    double b_square_roe = (B_roe[0] * B_roe[0] + B_roe[1] * B_roe[1] + B_roe[2] * B_roe[2]) * inv_sqrt_rho_prod;
    double a_square_roe = ((2 - eq->gamma) * X +
                           (eq->gamma - 1) * (H_roe - 0.5 * (v_roe[0] * v_roe[0] + v_roe[1] * v_roe[1] + v_roe[2] * v_roe[2]) -
                                              b_square_roe));
    double norm_squared = (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    double Bn = (B_roe[0] * n[0] + B_roe[1] * n[1] + B_roe[2] * n[2]);
    double B_roe_dot_n_squared = Bn * Bn / norm_squared;
    double c_a_roe = B_roe_dot_n_squared * inv_sqrt_rho_prod;
    double s = a_square_roe + b_square_roe;
    double a_star_roe = sqrt(s * s - 4 * a_square_roe * c_a_roe);
    *c_f = sqrt(0.5 * (a_square_roe + b_square_roe + a_star_roe) * norm_squared);
    *vel_out = (v_roe[0] * n[0] + v_roe[1] * n[1] + v_roe[2] * n[2]);
}

/* conservative part of the MHD surface / volume fluxes along a normal vector (tuples are encoded as one id) */
static void mhd_numflux_normal(const eqn_t *eq, int flux_id, const double *ul, const double *ur, const double *n, double *f) {
    switch (flux_id) {
    case TRIXI_B200_FLUX_CENTRAL:
    case TRIXI_B200_FLUX_CENTRAL_MHD_POWELL: {
        double fl[9], fr[9];
        mhd_flux_normal(eq, ul, n, fl);
        mhd_flux_normal(eq, ur, n, fr);
        for (int v = 0; v < 9; ++v) f[v] = 0.5 * (fl[v] + fr[v]);
        return;
    }
    case TRIXI_B200_FLUX_LLF:
    case TRIXI_B200_FLUX_LLF_NAIVE:
    case TRIXI_B200_FLUX_LLF_MHD_POWELL:
    case TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL: { /* max_abs_speed(_naive) :879-956 */
        int naive = flux_id == TRIXI_B200_FLUX_LLF_NAIVE || flux_id == TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL;
        double v_ll = (ul[1] / ul[0] * n[0] + ul[2] / ul[0] * n[1] + ul[3] / ul[0] * n[2]);
        double v_rr = (ur[1] / ur[0] * n[0] + ur[2] / ur[0] * n[1] + ur[3] / ur[0] * n[2]);
        double cf_ll = mhd_fast_wavespeed_normal(eq, ul, n), cf_rr = mhd_fast_wavespeed_normal(eq, ur, n);
        double lam = naive ? fmax(fabs(v_ll), fabs(v_rr)) + fmax(cf_ll, cf_rr) : fmax(fabs(v_ll) + cf_ll, fabs(v_rr) + cf_rr);
        double fl[9], fr[9];
        mhd_flux_normal(eq, ul, n, fl);
        mhd_flux_normal(eq, ur, n, fr);
        for (int v = 0; v < 9; ++v) f[v] = 0.5 * (fl[v] + fr[v]) + (-0.5 * lam * (ur[v] - ul[v]));
        return;
    }
    case TRIXI_B200_FLUX_HINDENLANG_GASSNER:
    case TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL:
        mhd_flux_hindenlang_gassner_normal(eq, ul, ur, n, f);
        return;
    case TRIXI_B200_FLUX_HLLE_MHD_POWELL: { /* min_max_speed_einfeldt :1132-1166 */
        double v_normal_ll = (ul[1] / ul[0] * n[0] + ul[2] / ul[0] * n[1] + ul[3] / ul[0] * n[2]);
        double v_normal_rr = (ur[1] / ur[0] * n[0] + ur[2] / ur[0] * n[1] + ur[3] / ur[0] * n[2]);
        double c_f_ll = mhd_fast_wavespeed_normal(eq, ul, n), c_f_rr = mhd_fast_wavespeed_normal(eq, ur, n), v_roe, c_f_roe;
        mhd_fast_wavespeed_roe_normal(eq, ul, ur, n, &v_roe, &c_f_roe);
        double lmin = fmin(v_normal_ll - c_f_ll, v_roe - c_f_roe), lmax = fmax(v_normal_rr + c_f_rr, v_roe + c_f_roe);
        if (lmin >= 0 && lmax >= 0) {
            mhd_flux_normal(eq, ul, n, f);
        } else if (lmax <= 0 && lmin <= 0) {
            mhd_flux_normal(eq, ur, n, f);
        } else {
            double fl[9], fr[9];
            mhd_flux_normal(eq, ul, n, fl);
            mhd_flux_normal(eq, ur, n, fr);
            double inv = 1 / (lmax - lmin);
            double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
            for (int v = 0; v < 9; ++v) f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
        }
        return;
    }
    default:
        for (int v = 0; v < 9; ++v) f[v] = NAN;
    }
}

/* ---- linear scalar advection (linear_scalar_advection_2d.jl:221-246) ----------------------------- */
static inline void adv_flux(const eqn_t *eq, const double *u, int o, double *f) { f[0] = eq->a[o] * u[0]; }

/* ---- generic pointwise dispatch --------------------------------------------------------------------- */
static inline int is_euler(const eqn_t *eq) {
    return eq->id == TRIXI_B200_EQ_EULER_2D || eq->id == TRIXI_B200_EQ_EULER_3D;
}

static inline int is_mhd(const eqn_t *eq) { return eq->id == TRIXI_B200_EQ_MHD_3D; }

static inline void phys_flux(const eqn_t *eq, const double *u, int o, double *f) {
    if (is_euler(eq))
        euler_flux(eq, u, o, f);
    else if (is_mhd(eq))
        mhd_flux(eq, u, o, f);
    else
        adv_flux(eq, u, o, f);
}

static inline double max_abs_speed_disp(const eqn_t *eq, const double *ul, const double *ur, int o, int naive) {
    if (is_euler(eq)) return euler_max_abs_speed(eq, ul, ur, o, naive);
    if (is_mhd(eq)) { /* ideal_glm_mhd_3d.jl:857-928 */
        double v_ll = ul[1 + o] / ul[0], v_rr = ur[1 + o] / ur[0];
        double cf_ll = mhd_fast_wavespeed(eq, ul, o), cf_rr = mhd_fast_wavespeed(eq, ur, o);
        return naive ? fmax(fabs(v_ll), fabs(v_rr)) + fmax(cf_ll, cf_rr) : fmax(fabs(v_ll) + cf_ll, fabs(v_rr) + cf_rr);
    }
    /* advection: max_abs_speed falls back to max_abs_speed_naive = |a| (numerical_fluxes.jl:219-225) */
    return fabs(eq->a[o]);
}

/* two-point numerical flux with an integer orientation (o is 0-based here) */
static void numflux(const eqn_t *eq, int flux_id, const double *ul, const double *ur, int o, double *f) {
    int nv = eq->nv;
    switch (flux_id) {
    case TRIXI_B200_FLUX_CENTRAL_MHD_POWELL:
    case TRIXI_B200_FLUX_CENTRAL: { /* numerical_fluxes.jl:17-25 */
        double fl[MAXV], fr[MAXV];
        phys_flux(eq, ul, o, fl);
        phys_flux(eq, ur, o, fr);
        for (int v = 0; v < nv; ++v) f[v] = 0.5 * (fl[v] + fr[v]);
        return;
    }
    case TRIXI_B200_FLUX_LLF:
    case TRIXI_B200_FLUX_LLF_NAIVE: { /* FluxPlusDissipation :37-45 + DissipationLocalLaxFriedrichs :172-178 */
        double fl[MAXV], fr[MAXV];
        phys_flux(eq, ul, o, fl);
        phys_flux(eq, ur, o, fr);
        double lam = max_abs_speed_disp(eq, ul, ur, o, flux_id == TRIXI_B200_FLUX_LLF_NAIVE);
        for (int v = 0; v < nv; ++v) f[v] = 0.5 * (fl[v] + fr[v]) + (-0.5 * lam * (ur[v] - ul[v]));
        return;
    }
    case TRIXI_B200_FLUX_HLLC:
        euler_flux_hllc(eq, ul, ur, o, NULL, f);
        return;
    case TRIXI_B200_FLUX_HLLE:
    case TRIXI_B200_FLUX_HLL_DAVIS:
    case TRIXI_B200_FLUX_HLL_NAIVE: { /* FluxHLL numerical_fluxes.jl:422-440 */
        double lmin, lmax;
        if (flux_id == TRIXI_B200_FLUX_HLLE)
            euler_min_max_speed_einfeldt(eq, ul, ur, o, NULL, &lmin, &lmax);
        else
            euler_min_max_speed(eq, ul, ur, o, flux_id == TRIXI_B200_FLUX_HLL_NAIVE, &lmin, &lmax);
        if (lmin >= 0 && lmax >= 0) {
            phys_flux(eq, ul, o, f);
        } else if (lmax <= 0 && lmin <= 0) {
            phys_flux(eq, ur, o, f);
        } else {
            double fl[MAXV], fr[MAXV];
            phys_flux(eq, ul, o, fl);
            phys_flux(eq, ur, o, fr);
            double inv = 1.0 / (lmax - lmin);
            double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
            for (int v = 0; v < nv; ++v)
                f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
        }
        return;
    }
    case TRIXI_B200_FLUX_RANOCHA:
    case TRIXI_B200_FLUX_RANOCHA_TURBO: /* turbo computes the same flux with hoisted logs; oracle = generic */
        euler_flux_ranocha(eq, ul, ur, o, f);
        return;
    case TRIXI_B200_FLUX_SHIMA_ETAL:
        euler_flux_shima(eq, ul, ur, o, f);
        return;
    case TRIXI_B200_FLUX_KENNEDY_GRUBER:
        euler_flux_kennedy_gruber(eq, ul, ur, o, f);
        return;
    case TRIXI_B200_FLUX_CHANDRASHEKAR:
        euler_flux_chandrashekar(eq, ul, ur, o, f);
        return;
    case TRIXI_B200_FLUX_GODUNOV: /* linear_scalar_advection_2d.jl:248-260 */
        f[0] = eq->a[o] >= 0 ? eq->a[o] * ul[0] : eq->a[o] * ur[0];
        return;
    case TRIXI_B200_FLUX_HINDENLANG_GASSNER:
    case TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL: /* conservative part of the tuple */
        mhd_flux_hindenlang_gassner(eq, ul, ur, o, f);
        return;
    case TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL: /* (FluxLaxFriedrichs(max_abs_speed_naive), powell) */
        numflux(eq, TRIXI_B200_FLUX_LLF_NAIVE, ul, ur, o, f);
        return;
    case TRIXI_B200_FLUX_HLLE_MHD_POWELL: { /* FluxHLL numerical_fluxes.jl:422-449 with min_max_speed_einfeldt
                                              ideal_glm_mhd_3d.jl:1094-1130 */
        double c_f_ll = mhd_fast_wavespeed(eq, ul, o), c_f_rr = mhd_fast_wavespeed(eq, ur, o), vel_roe, c_f_roe;
        mhd_fast_wavespeed_roe(eq, ul, ur, o, &vel_roe, &c_f_roe);
        double lmin = fmin(ul[1 + o] / ul[0] - c_f_ll, vel_roe - c_f_roe);
        double lmax = fmax(ur[1 + o] / ur[0] + c_f_rr, vel_roe + c_f_roe);
        if (lmin >= 0 && lmax >= 0) {
            mhd_flux(eq, ul, o, f);
        } else if (lmax <= 0 && lmin <= 0) {
            mhd_flux(eq, ur, o, f);
        } else {
            double fl[MAXV], fr[MAXV];
            mhd_flux(eq, ul, o, fl);
            mhd_flux(eq, ur, o, fr);
            double inv = 1 / (lmax - lmin);
            double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
            for (int v = 0; v < nv; ++v) f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
        }
        return;
    }
    case TRIXI_B200_FLUX_LLF_MHD_POWELL: /* (flux_lax_friedrichs, powell) */
        numflux(eq, TRIXI_B200_FLUX_LLF, ul, ur, o, f);
        return;
    default:
        for (int v = 0; v < nv; ++v) f[v] = NAN;
    }
}

static inline int flux_has_noncons(int flux_id) {
    return flux_id == TRIXI_B200_FLUX_HINDENLANG_GASSNER_POWELL || flux_id == TRIXI_B200_FLUX_LLF_MHD_POWELL ||
           flux_id == TRIXI_B200_FLUX_LLF_NAIVE_MHD_POWELL || flux_id == TRIXI_B200_FLUX_HLLE_MHD_POWELL ||
           flux_id == TRIXI_B200_FLUX_CENTRAL_MHD_POWELL;
}

/* ---- normal-direction versions (curved meshes) -------------------------------------------------- */
/* flux(u, normal_direction, eq) compressible_euler_3d.jl:449-463 */
static inline void euler_flux_normal(const eqn_t *eq, const double *u, const double *n, double *f) {
    int nd = eq->nd;
    double rho, v[3], p;
    euler_cons2prim(eq, u, &rho, v, &p);
    double v_normal = 0.0;
    for (int d = 0; d < nd; ++d) v_normal += v[d] * n[d];
    double rho_v_normal = rho * v_normal;
    f[0] = rho_v_normal;
    for (int d = 0; d < nd; ++d) f[1 + d] = rho_v_normal * v[d] + p * n[d];
    f[nd + 1] = (u[nd + 1] + p) * v_normal;
}

/* flux_ranocha(u_ll, u_rr, normal_direction, eq) compressible_euler_3d.jl:795-828 */
static inline void euler_flux_ranocha_normal(const eqn_t *eq, const double *ul, const double *ur, const double *n,
                                             double *f) {
    int nd = eq->nd;
    double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
    euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
    euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
    double v_dot_n_ll = 0.0, v_dot_n_rr = 0.0;
    for (int d = 0; d < nd; ++d) {
        v_dot_n_ll += v_ll[d] * n[d];
        v_dot_n_rr += v_rr[d] * n[d];
    }
    double rho_mean = ln_mean(rho_ll, rho_rr);
    double inv_rho_p_mean = p_ll * p_rr * inv_ln_mean(rho_ll * p_rr, rho_rr * p_ll);
    double v_avg[3], vsq = 0.0;
    for (int d = 0; d < nd; ++d) {
        v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
        vsq += v_ll[d] * v_rr[d];
    }
    double p_avg = 0.5 * (p_ll + p_rr);
    double velocity_square_avg = 0.5 * vsq;
    double f1 = rho_mean * 0.5 * (v_dot_n_ll + v_dot_n_rr);
    f[0] = f1;
    for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d] + p_avg * n[d];
    f[nd + 1] = f1 * (velocity_square_avg + inv_rho_p_mean * eq->inv_gm1) +
                0.5 * (p_ll * v_dot_n_rr + p_rr * v_dot_n_ll);
}

static inline double vec_norm(int nd, const double *n) {
    double s = 0.0;
    for (int d = 0; d < nd; ++d) s += n[d] * n[d];
    return sqrt(s);
}

static inline void phys_flux_normal(const eqn_t *eq, const double *u, const double *n, double *f) {
    if (is_euler(eq)) {
        euler_flux_normal(eq, u, n, f);
    } else if (is_mhd(eq)) {
        mhd_flux_normal(eq, u, n, f);
    } else { /* linear_scalar_advection_2d.jl:233-238 */
        double a = 0.0;
        for (int d = 0; d < eq->nd; ++d) a += eq->a[d] * n[d];
        f[0] = a * u[0];
    }
}

/* two-point numerical flux with a (non-normalised) normal vector */
static void numflux_normal(const eqn_t *eq, int flux_id, const double *ul, const double *ur, const double *n,
                           double *f) {
    int nv = eq->nv, nd = eq->nd;
    if (is_mhd(eq)) {
        mhd_numflux_normal(eq, flux_id, ul, ur, n, f);
        return;
    }
    switch (flux_id) {
    case TRIXI_B200_FLUX_CENTRAL: {
        double fl[MAXV], fr[MAXV];
        phys_flux_normal(eq, ul, n, fl);
        phys_flux_normal(eq, ur, n, fr);
        for (int v = 0; v < nv; ++v) f[v] = 0.5 * (fl[v] + fr[v]);
        return;
    }
    case TRIXI_B200_FLUX_LLF:
    case TRIXI_B200_FLUX_LLF_NAIVE: {
        double fl[MAXV], fr[MAXV], lam;
        phys_flux_normal(eq, ul, n, fl);
        phys_flux_normal(eq, ur, n, fr);
        if (is_euler(eq)) {
            /* max_abs_speed_naive :1135-1153, max_abs_speed :1180-1199 */
            double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
            euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
            euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
            double vl = 0.0, vr = 0.0;
            for (int d = 0; d < nd; ++d) {
                vl += v_ll[d] * n[d];
                vr += v_rr[d] * n[d];
            }
            double c_ll = sqrt(eq->gamma * p_ll / rho_ll), c_rr = sqrt(eq->gamma * p_rr / rho_rr);
            double norm_ = vec_norm(nd, n);
            lam = flux_id == TRIXI_B200_FLUX_LLF_NAIVE ? fmax(fabs(vl), fabs(vr)) + fmax(c_ll, c_rr) * norm_
                                                       : fmax(fabs(vl) + c_ll * norm_, fabs(vr) + c_rr * norm_);
        } else { /* linear_scalar_advection_2d.jl:241-246 */
            double a = 0.0;
            for (int d = 0; d < nd; ++d) a += eq->a[d] * n[d];
            lam = fabs(a);
        }
        for (int v = 0; v < nv; ++v) f[v] = 0.5 * (fl[v] + fr[v]) + (-0.5 * lam * (ur[v] - ul[v]));
        return;
    }
    case TRIXI_B200_FLUX_HLLE:
    case TRIXI_B200_FLUX_HLL_DAVIS:
    case TRIXI_B200_FLUX_HLL_NAIVE: { /* compressible_euler_3d.jl:1220-1237, 1263-1285, 1723-1771 */
        double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
        euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
        euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
        double vl = 0.0, vr = 0.0;
        for (int d = 0; d < nd; ++d) {
            vl += v_ll[d] * n[d];
            vr += v_rr[d] * n[d];
        }
        double norm_ = vec_norm(nd, n);
        double c_ll = sqrt(eq->gamma * p_ll / rho_ll) * norm_, c_rr = sqrt(eq->gamma * p_rr / rho_rr) * norm_;
        double lmin, lmax;
        if (flux_id == TRIXI_B200_FLUX_HLLE) {
            euler_min_max_speed_einfeldt(eq, ul, ur, 0, n, &lmin, &lmax);
        } else if (flux_id == TRIXI_B200_FLUX_HLL_NAIVE) {
            lmin = vl - c_ll;
            lmax = vr + c_rr;
        } else {
            lmin = fmin(vl - c_ll, vr - c_rr);
            lmax = fmax(vl + c_ll, vr + c_rr);
        }
        if (lmin >= 0 && lmax >= 0) {
            phys_flux_normal(eq, ul, n, f);
        } else if (lmax <= 0 && lmin <= 0) {
            phys_flux_normal(eq, ur, n, f);
        } else {
            double fl[MAXV], fr[MAXV];
            phys_flux_normal(eq, ul, n, fl);
            phys_flux_normal(eq, ur, n, fr);
            double inv = 1.0 / (lmax - lmin);
            double factor_ll = lmax * inv, factor_rr = lmin * inv, factor_diss = lmin * lmax * inv;
            for (int v = 0; v < nv; ++v)
                f[v] = factor_ll * fl[v] - factor_rr * fr[v] + factor_diss * (ur[v] - ul[v]);
        }
        return;
    }
    case TRIXI_B200_FLUX_RANOCHA:
    case TRIXI_B200_FLUX_RANOCHA_TURBO:
        euler_flux_ranocha_normal(eq, ul, ur, n, f);
        return;
    case TRIXI_B200_FLUX_KENNEDY_GRUBER:
    case TRIXI_B200_FLUX_SHIMA_ETAL: { /* compressible_euler_3d.jl:602-627 / :512-547 */
        double rho_ll, v_ll[3], p_ll, rho_rr, v_rr[3], p_rr;
        euler_cons2prim(eq, ul, &rho_ll, v_ll, &p_ll);
        euler_cons2prim(eq, ur, &rho_rr, v_rr, &p_rr);
        double rho_avg = 0.5 * (rho_ll + rho_rr), p_avg = 0.5 * (p_ll + p_rr), v_avg[3];
        for (int d = 0; d < nd; ++d) v_avg[d] = 0.5 * (v_ll[d] + v_rr[d]);
        if (flux_id == TRIXI_B200_FLUX_KENNEDY_GRUBER) {
            double e_avg = 0.5 * (ul[nd + 1] / rho_ll + ur[nd + 1] / rho_rr);
            double v_dot_n_avg = 0.0;
            for (int d = 0; d < nd; ++d) v_dot_n_avg += v_avg[d] * n[d];
            double f1 = rho_avg * v_dot_n_avg;
            f[0] = f1;
            for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d] + p_avg * n[d];
            f[nd + 1] = f1 * e_avg + p_avg * v_dot_n_avg;
        } else {
            double vl = 0.0, vr = 0.0, vsq = 0.0;
            for (int d = 0; d < nd; ++d) {
                vl += v_ll[d] * n[d];
                vr += v_rr[d] * n[d];
                vsq += v_ll[d] * v_rr[d];
            }
            double v_dot_n_avg = 0.5 * (vl + vr), velocity_square_avg = 0.5 * vsq;
            double f1 = rho_avg * v_dot_n_avg;
            f[0] = f1;
            for (int d = 0; d < nd; ++d) f[1 + d] = f1 * v_avg[d] + p_avg * n[d];
            f[nd + 1] = f1 * velocity_square_avg + p_avg * v_dot_n_avg * eq->inv_gm1 + 0.5 * (p_ll * vr + p_rr * vl);
        }
        return;
    }
    case TRIXI_B200_FLUX_GODUNOV: { /* linear_scalar_advection_2d.jl:262-275 */
        double a = 0.0;
        for (int d = 0; d < nd; ++d) a += eq->a[d] * n[d];
        f[0] = a >= 0 ? a * ul[0] : a * ur[0];
        return;
    }
    case TRIXI_B200_FLUX_CHANDRASHEKAR:
        euler_flux_chandrashekar_normal(eq, ul, ur, n, f);
        return;
    case TRIXI_B200_FLUX_HLLC:
        euler_flux_hllc(eq, ul, ur, 0, n, f);
        return;
    default:
        for (int v = 0; v < nv; ++v) f[v] = NAN;
    }
}

/* ---- initial conditions used as Dirichlet data ----------------------------------------------------- */
static void ic_eval(const eqn_t *eq, int ic, const double *x, double t, double *u) {
    int nd = eq->nd;
    if (is_euler(eq)) {
        switch (ic) {
        case TRIXI_B200_IC_CONSTANT: /* compressible_euler_3d.jl:78-86, _2d.jl:71-78 */
            u[0] = 1.0;
            u[1] = 0.1;
            u[2] = -0.2;
            if (nd == 3) u[3] = 0.7;
            u[nd + 1] = 10.0;
            return;
        case TRIXI_B200_IC_CONVERGENCE_TEST: { /* compressible_euler_3d.jl:94-111 */
            double s = 0.0;
            for (int d = 0; d < nd; ++d) s += x[d];
            double omega = 2 * M_PI * 0.5;
            double ini = 2 + 0.1 * sin(omega * (s - t));
            for (int v = 0; v <= nd; ++v) u[v] = ini;
            u[nd + 1] = ini * ini;
            return;
        }
        case TRIXI_B200_IC_EOC_TEST_COUPLED_EULER_GRAVITY: { /* compressible_euler_3d.jl:196-215, _2d.jl:212-230 */
            double s = 0.0;
            for (int d = 0; d < nd; ++d) s += x[d];
            double ini = 2 + 0.1 * sin(M_PI * fmod(s - t, 2.0));
            double p = nd == 3 ? ini * ini * 1 * 2 / (3 * M_PI) : ini * ini * 1 / M_PI;
            u[0] = ini;
            for (int d = 0; d < nd; ++d) u[1 + d] = ini * 1.0;
            u[nd + 1] = p * eq->inv_gm1 + 0.5 * (nd * (ini * 1.0) * 1.0);
            return;
        }
        default:
            break;
        }
    } else if (is_mhd(eq)) {
        double prim[9];
        if (ic == TRIXI_B200_IC_CONSTANT) { /* ideal_glm_mhd_3d.jl:101-113 */
            static const double c[9] = {1.0, 0.1, -0.2, -0.5, 50.0, 3.0, -1.2, 0.5, 0.0};
            for (int v = 0; v < 9; ++v) u[v] = c[v];
            return;
        }
        if (ic == TRIXI_B200_IC_CONVERGENCE_TEST) { /* :124-151: Alfven wave, gamma = 5/3 */
            double p = 1, omega = 2 * M_PI, r = 2, e = 0.2;
            double nx = 1 / sqrt(r * r + 1), ny = r / sqrt(r * r + 1), sqr = 1;
            double Va = omega / (ny * sqr);
            double phi_alv = omega / ny * (nx * (x[0] - 0.5 * r) + ny * (x[1] - 0.5 * r)) - Va * t;
            double rho = 1;
            prim[0] = rho;
            prim[1] = -e * ny * cos(phi_alv) / rho;
            prim[2] = e * nx * cos(phi_alv) / rho;
            prim[3] = e * sin(phi_alv) / rho;
            prim[4] = p;
            prim[5] = nx - rho * prim[1] * sqr;
            prim[6] = ny - rho * prim[2] * sqr;
            prim[7] = -rho * prim[3] * sqr;
            prim[8] = 0;
            /* prim2cons :1273-1284 */
            u[0] = prim[0];
            u[1] = prim[0] * prim[1];
            u[2] = prim[0] * prim[2];
            u[3] = prim[0] * prim[3];
            u[4] = prim[4] * eq->inv_gm1 + 0.5 * (u[1] * prim[1] + u[2] * prim[2] + u[3] * prim[3]) +
                   0.5 * (prim[5] * prim[5] + prim[6] * prim[6] + prim[7] * prim[7]) + 0.5 * prim[8] * prim[8];
            for (int v = 5; v < 9; ++v) u[v] = prim[v];
            return;
        }
    } else if (ic == TRIXI_B200_IC_CONVERGENCE_TEST) { /* linear_scalar_advection_2d.jl:67-80 */
        double s = 0.0;
        for (int d = 0; d < nd; ++d) s += x[d] - eq->a[d] * t;
        u[0] = 1 + 0.5 * sin(2 * M_PI * 0.5 * s);
        return;
    } else if (ic == TRIXI_B200_IC_CONSTANT) {
        u[0] = 2.0;
        return;
    }
    for (int v = 0; v < eq->nv; ++v) u[v] = NAN;
}

/* boundary_condition_slip_wall compressible_euler_3d.jl:315-366 (normal version) specialised to the
 * unit normal of `orientation` as done by :374-389, then :398-414 for the sign by direction */
static void euler_slip_wall_normal(const eqn_t *eq, const double *u_inner, const double *nrm, double *f) {
    int nd = eq->nd;
    double norm_ = 0.0;
    for (int d = 0; d < nd; ++d) norm_ += nrm[d] * nrm[d];
    norm_ = sqrt(norm_);
    double normal[3] = {0, 0, 0};
    for (int d = 0; d < nd; ++d) normal[d] = nrm[d] / norm_;
    /* rotate_to_x: only the normal velocity and rho, p are needed */
    double rho = u_inner[0];
    double rv_n = 0.0;
    for (int d = 0; d < nd; ++d) rv_n += normal[d] * u_inner[1 + d];
    /* u_local = (rho, rho v_n, rho v_t1, rho v_t2, E): kinetic energy is rotation invariant */
    double kin = 0.0;
    for (int d = 0; d < nd; ++d) kin += u_inner[1 + d] * (u_inner[1 + d] / rho);
    double v_normal = rv_n / rho;
    double p_local = (eq->gamma - 1) * (u_inner[nd + 1] - 0.5 * kin);
    double p_star;
    if (v_normal <= 0) {
        double sound_speed = sqrt(eq->gamma * p_local / rho);
        double base = 1 + 0.5 * (eq->gamma - 1) * v_normal / sound_speed;
        if (base >= 0)
            p_star = p_local * pow(base, 2 * eq->gamma * eq->inv_gm1);
        else
            p_star = 0.0;
    } else {
        double A = 2 / ((eq->gamma + 1) * rho);
        double B = p_local * (eq->gamma - 1) / (eq->gamma + 1);
        p_star = p_local + 0.5 * v_normal / A * (v_normal + sqrt(v_normal * v_normal + 4 * A * (p_local + B)));
    }
    f[0] = 0.0;
    for (int d = 0; d < nd; ++d) f[1 + d] = p_star * normal[d] * norm_;
    f[nd + 1] = 0.0;
}

/* boundary flux for TreeMesh: bc(u_inner, orientation, direction, x, t, surface_flux, eq)
 * (dg_3d.jl:757); direction is 1-based like the reference, o 0-based */
static void boundary_flux(const eqn_t *eq, int bc, int ic, int surface_flux, const double *u_inner, int o,
                          int direction, const double *x, double t, double *f) {
    if (bc == TRIXI_B200_BC_DIRICHLET) { /* equations.jl:164-183 */
        double ub[MAXV];
        ic_eval(eq, ic, x, t, ub);
        if (direction % 2 == 0)
            numflux(eq, surface_flux, u_inner, ub, o, f);
        else
            numflux(eq, surface_flux, ub, u_inner, o, f);
    } else if (bc == TRIXI_B200_BC_SLIP_WALL) { /* compressible_euler_3d.jl:374-414 */
        double nrm[3] = {0, 0, 0};
        nrm[o] = 1.0;
        if (direction % 2 == 1) {
            double mn[3] = {-nrm[0], -nrm[1], -nrm[2]};
            euler_slip_wall_normal(eq, u_inner, mn, f);
            for (int v = 0; v < eq->nv; ++v) f[v] = -f[v];
        } else {
            euler_slip_wall_normal(eq, u_inner, nrm, f);
        }
    } else {
        for (int v = 0; v < eq->nv; ++v) f[v] = NAN;
    }
}

/* boundary flux for curved meshes: bc(u_inner, normal, direction, x, t, surface_flux, eq)
 * (dgsem_structured/dg.jl:124-165) */
static void boundary_flux_normal(const eqn_t *eq, int bc, int ic, int surface_flux, const double *u_inner,
                                 const double *n, int direction, const double *x, double t, double *f) {
    if (bc == TRIXI_B200_BC_DIRICHLET) { /* equations.jl:164-183 */
        double ub[MAXV];
        ic_eval(eq, ic, x, t, ub);
        if (direction % 2 == 0)
            numflux_normal(eq, surface_flux, u_inner, ub, n, f);
        else
            numflux_normal(eq, surface_flux, ub, u_inner, n, f);
    } else if (bc == TRIXI_B200_BC_SLIP_WALL) { /* compressible_euler_3d.jl:398-414 */
        if (direction % 2 == 1) {
            double mn[3] = {-n[0], -n[1], eq->nd == 3 ? -n[2] : 0.0};
            euler_slip_wall_normal(eq, u_inner, mn, f);
            for (int v = 0; v < eq->nv; ++v) f[v] = -f[v];
        } else {
            euler_slip_wall_normal(eq, u_inner, n, f);
        }
    } else {
        for (int v = 0; v < eq->nv; ++v) f[v] = NAN;
    }
}

/* ---- source terms ---------------------------------------------------------------------------------- */
static void source_terms(const eqn_t *eq, int src, const double *u, const double *x, double t, double *du) {
    int nd = eq->nd;
    (void)u;
    double s = 0.0;
    for (int d = 0; d < nd; ++d) s += x[d];
    switch (src) {
    case TRIXI_B200_SRC_CONVERGENCE_TEST: {
        double omega = 2 * M_PI * 0.5, g = eq->gamma;
        double si = sin(omega * (s - t)), co = cos(omega * (s - t));
        double rho = 2 + 0.1 * si;
        double rho_x = omega * 0.1 * co;
        if (nd == 3) { /* compressible_euler_3d.jl:127-153 */
            double tmp = (2 * rho - 1.5) * (g - 1);
            du[0] = 2 * rho_x;
            du[1] = du[2] = du[3] = rho_x * (2 + tmp);
            du[4] = rho_x * (4 * rho + 3 * tmp);
        } else { /* compressible_euler_2d.jl:120-145 */
            double tmp = (2 * rho - 1) * (g - 1);
            du[0] = rho_x;
            du[1] = du[2] = rho_x * (1 + tmp);
            du[3] = 2 * rho_x * (rho + tmp);
        }
        return;
    }
    case TRIXI_B200_SRC_EOC_TEST_EULER: {
        /* sincospi(x) */
        double r = fmod(s - t, 2.0);
        double si = sin(M_PI * r), co = cos(M_PI * r);
        double rhox = 0.1 * M_PI * co, rho = 2 + 0.1 * si;
        if (nd == 3) { /* compressible_euler_3d.jl:265-284 */
            double C_grav = -4.0 * 1 / (3 * M_PI);
            du[0] = rhox * 2;
            du[1] = du[2] = du[3] = rhox * (2 - C_grav * rho);
            du[4] = rhox * (3 - 5 * C_grav * rho);
        } else { /* compressible_euler_2d.jl:272-292 */
            double C_grav = -2.0 * 1 / M_PI;
            du[0] = rhox;
            du[1] = du[2] = rhox * (1 - C_grav * rho);
            du[3] = rhox * (1 - 3 * C_grav * rho);
        }
        return;
    }
    case TRIXI_B200_SRC_EOC_TEST_COUPLED_EULER_GRAVITY: {
        double r = fmod(s - t, 2.0);
        double si = sin(M_PI * r), co = cos(M_PI * r);
        double rhox = 0.1 * M_PI * co, rho = 2 + 0.1 * si;
        if (nd == 3) { /* compressible_euler_3d.jl:228-249 */
            double C_grav = -4.0 * 1 / (3 * M_PI);
            du[0] = du[1] = du[2] = du[3] = 2 * rhox;
            du[4] = 2 * rhox * (1.5 - C_grav * rho);
        } else { /* compressible_euler_2d.jl:241-261 */
            double C_grav = -2.0 * 1 / M_PI;
            du[0] = du[1] = du[2] = rhox;
            du[3] = (1 - C_grav * rho) * rhox;
        }
        return;
    }
    default:
        for (int v = 0; v < eq->nv; ++v) du[v] = 0.0;
    }
}

/* ---- index helpers ------------------------------------------------------------------------------------ */
/* volume node (0-based linear) of face node (a, b) on the layer `s` normal to orientation o
 * (face node order: x-faces (j,k), y-faces (i,k), z-faces (i,j), dg_3d.jl:540-563) */
static inline int face_to_volume_node(int nd, int n, int o, int s, int a, int b) {
    if (nd == 2) return o == 0 ? s + n * a : a + n * s;
    if (o == 0) return s + n * (a + n * b);
    if (o == 1) return a + n * (s + n * b);
    return a + n * (b + n * s);
}

/* ---- stages: src/solvers/dgsem_tree/dg_3d.jl, dg_2d.jl --------------------------------------------- */
/* set_zero! solvers.jl:8-25 */
void oracle_set_zero(const trixi_b200_desc *d, double *du) {
    int64_t len = (int64_t)d->nvars * ipow(d->nnodes, d->ndims) * d->nelements;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < len; ++i) du[i] = 0.0;
}

/* weak_form_kernel! dg_3d.jl:133-164 / dg_2d.jl:195-221 */
static void weak_form_kernel(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1;
    const double *Dhat = d->derivative_hat;
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                const double *un = u + (int64_t)nv * (i + n * (j + n * k));
                double f[MAXV];
                phys_flux(eq, un, 0, f);
                for (int ii = 0; ii < n; ++ii) {
                    double w = Dhat[ii + n * i];
                    double *t = du + (int64_t)nv * (ii + n * (j + n * k));
                    for (int v = 0; v < nv; ++v) t[v] = t[v] + w * f[v];
                }
                phys_flux(eq, un, 1, f);
                for (int jj = 0; jj < n; ++jj) {
                    double w = Dhat[jj + n * j];
                    double *t = du + (int64_t)nv * (i + n * (jj + n * k));
                    for (int v = 0; v < nv; ++v) t[v] = t[v] + w * f[v];
                }
                if (nd == 3) {
                    phys_flux(eq, un, 2, f);
                    for (int kk = 0; kk < n; ++kk) {
                        double w = Dhat[kk + n * k];
                        double *t = du + (int64_t)nv * (i + n * (j + n * kk));
                        for (int v = 0; v < nv; ++v) t[v] = t[v] + w * f[v];
                    }
                }
            }
}

/* flux_differencing_kernel! dg_3d.jl:166-214 / dg_2d.jl:223-259 */
static void flux_differencing_kernel(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u,
                                     double alpha) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1;
    const double *Ds = d->derivative_split;
    int vf = d->volume_flux;
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                int64_t node = i + n * (j + n * k);
                const double *un = u + nv * node;
                double f[MAXV];
                for (int ii = i + 1; ii < n; ++ii) {
                    int64_t node2 = ii + n * (j + n * k);
                    numflux(eq, vf, un, u + nv * node2, 0, f);
                    double w1 = alpha * Ds[i + n * ii], w2 = alpha * Ds[ii + n * i];
                    for (int v = 0; v < nv; ++v) du[nv * node + v] = du[nv * node + v] + w1 * f[v];
                    for (int v = 0; v < nv; ++v) du[nv * node2 + v] = du[nv * node2 + v] + w2 * f[v];
                }
                for (int jj = j + 1; jj < n; ++jj) {
                    int64_t node2 = i + n * (jj + n * k);
                    numflux(eq, vf, un, u + nv * node2, 1, f);
                    double w1 = alpha * Ds[j + n * jj], w2 = alpha * Ds[jj + n * j];
                    for (int v = 0; v < nv; ++v) du[nv * node + v] = du[nv * node + v] + w1 * f[v];
                    for (int v = 0; v < nv; ++v) du[nv * node2 + v] = du[nv * node2 + v] + w2 * f[v];
                }
                if (nd == 3)
                    for (int kk = k + 1; kk < n; ++kk) {
                        int64_t node2 = i + n * (j + n * kk);
                        numflux(eq, vf, un, u + nv * node2, 2, f);
                        double w1 = alpha * Ds[k + n * kk], w2 = alpha * Ds[kk + n * k];
                        for (int v = 0; v < nv; ++v) du[nv * node + v] = du[nv * node + v] + w1 * f[v];
                        for (int v = 0; v < nv; ++v) du[nv * node2 + v] = du[nv * node2 + v] + w2 * f[v];
                    }
            }
}


/* flux_differencing_kernel! for flux_ranocha_turbo on TreeMesh{3} (dgsem_tree/dg_3d_compressible_euler.jl:265-617):
 * primitive variables and log(rho), log(p) once per node, the logarithmic means inlined with
 * z = (y - x)^2 / (x + y)^2 and the branch as a select, SIMD along the 16 lines of a direction.  The reference
 * permutes its temporaries so that the SIMD index is contiguous; here the temporaries are SoA [7][64] and the
 * inner loop runs over the n^2 lines with their stride (same arithmetic, same summation order per node: x pairs,
 * then y pairs, then z pairs, each in the order (i, ii) of the triangular loop). */
static void flux_differencing_kernel_turbo(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u,
                                           double alpha) {
    enum { N = 4, NN = 64 };
    const double *Ds = d->derivative_split;
    double prim[7][NN], acc[5][NN];
    for (int q = 0; q < NN; ++q) {
        double rho = u[5 * q], rv1 = u[5 * q + 1], rv2 = u[5 * q + 2], rv3 = u[5 * q + 3], rho_e = u[5 * q + 4];
        double v1 = rv1 / rho, v2 = rv2 / rho, v3 = rv3 / rho;
        double p = (eq->gamma - 1) * (rho_e - 0.5 * (rv1 * v1 + rv2 * v2 + rv3 * v3));
        prim[0][q] = rho;
        prim[1][q] = v1;
        prim[2][q] = v2;
        prim[3][q] = v3;
        prim[4][q] = p;
        prim[5][q] = log(rho);
        prim[6][q] = log(p);
        for (int v = 0; v < 5; ++v) acc[v][q] = 0.0;
    }
    const int stride[3] = {1, N, N * N};
    for (int o = 0; o < 3; ++o) {
        /* the 16 lines of direction o start at the nodes whose o-th index is 0 */
        int line0[16], nl = 0;
        for (int q = 0; q < NN; ++q)
            if ((q / stride[o]) % N == 0) line0[nl++] = q;
        for (int i = 0; i < N; ++i)
            for (int ii = i + 1; ii < N; ++ii) {
                const double factor_i = alpha * Ds[i + N * ii], factor_ii = alpha * Ds[ii + N * i];
#pragma omp simd
                for (int l = 0; l < 16; ++l) {
                    const int a = line0[l] + i * stride[o], b = line0[l] + ii * stride[o];
                    const double rho_ll = prim[0][a], p_ll = prim[4][a], log_rho_ll = prim[5][a], log_p_ll = prim[6][a];
                    const double rho_rr = prim[0][b], p_rr = prim[4][b], log_rho_rr = prim[5][b], log_p_rr = prim[6][b];
                    const double v1_ll = prim[1][a], v2_ll = prim[2][a], v3_ll = prim[3][a];
                    const double v1_rr = prim[1][b], v2_rr = prim[2][b], v3_rr = prim[3][b];
                    const double x1_plus_y1 = rho_ll + rho_rr, y1_minus_x1 = rho_rr - rho_ll;
                    const double z1 = (y1_minus_x1 * y1_minus_x1) / (x1_plus_y1 * x1_plus_y1);
                    const double special_path1 = x1_plus_y1 / (2 + z1 * (2.0 / 3 + z1 * (2.0 / 5 + 2.0 / 7 * z1)));
                    const double regular_path1 = y1_minus_x1 / (log_rho_rr - log_rho_ll);
                    const double rho_mean = z1 < 1.0e-4 ? special_path1 : regular_path1;
                    const double x2 = rho_ll * p_rr, log_x2 = log_rho_ll + log_p_rr;
                    const double y2 = rho_rr * p_ll, log_y2 = log_rho_rr + log_p_ll;
                    const double x2_plus_y2 = x2 + y2, y2_minus_x2 = y2 - x2;
                    const double z2 = (y2_minus_x2 * y2_minus_x2) / (x2_plus_y2 * x2_plus_y2);
                    const double special_path2 = (2 + z2 * (2.0 / 3 + z2 * (2.0 / 5 + 2.0 / 7 * z2))) / x2_plus_y2;
                    const double regular_path2 = (log_y2 - log_x2) / y2_minus_x2;
                    const double inv_rho_p_mean = p_ll * p_rr * (z2 < 1.0e-4 ? special_path2 : regular_path2);
                    const double v1_avg = 0.5 * (v1_ll + v1_rr), v2_avg = 0.5 * (v2_ll + v2_rr), v3_avg = 0.5 * (v3_ll + v3_rr);
                    const double p_avg = 0.5 * (p_ll + p_rr);
                    const double velocity_square_avg = 0.5 * (v1_ll * v1_rr + v2_ll * v2_rr + v3_ll * v3_rr);
                    const double vn_avg = o == 0 ? v1_avg : (o == 1 ? v2_avg : v3_avg);
                    const double vn_ll = o == 0 ? v1_ll : (o == 1 ? v2_ll : v3_ll);
                    const double vn_rr = o == 0 ? v1_rr : (o == 1 ? v2_rr : v3_rr);
                    const double f1 = rho_mean * vn_avg;
                    const double f2 = f1 * v1_avg + (o == 0 ? p_avg : 0.0);
                    const double f3 = f1 * v2_avg + (o == 1 ? p_avg : 0.0);
                    const double f4 = f1 * v3_avg + (o == 2 ? p_avg : 0.0);
                    const double f5 = f1 * (velocity_square_avg + inv_rho_p_mean * eq->inv_gm1) +
                                      0.5 * (p_ll * vn_rr + p_rr * vn_ll);
                    acc[0][a] += factor_i * f1;
                    acc[1][a] += factor_i * f2;
                    acc[2][a] += factor_i * f3;
                    acc[3][a] += factor_i * f4;
                    acc[4][a] += factor_i * f5;
                    acc[0][b] += factor_ii * f1;
                    acc[1][b] += factor_ii * f2;
                    acc[2][b] += factor_ii * f3;
                    acc[3][b] += factor_ii * f4;
                    acc[4][b] += factor_ii * f5;
                }
            }
    }
    for (int q = 0; q < NN; ++q)
        for (int v = 0; v < 5; ++v) du[5 * q + v] += acc[v][q];
}

/* flux_differencing_kernel! with nonconservative terms dg_3d.jl:216-266: the nonsymmetric part */
static void flux_differencing_noncons(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u,
                                      double alpha) {
    int n = d->nnodes, nv = d->nvars;
    const double *Ds = d->derivative_split;
    int stride[3] = {1, n, n * n};
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                int idx[3] = {i, j, k};
                int64_t node = i + n * (j + n * k);
                const double *un = u + nv * node;
                double integral_contribution[MAXV] = {0};
                for (int a = 0; a < 3; ++a)
                    for (int ii = 0; ii < n; ++ii) {
                        int64_t node2 = node + (ii - idx[a]) * stride[a];
                        double g[MAXV];
                        mhd_noncons_powell(eq, un, u + nv * node2, a, g);
                        double w = Ds[idx[a] + n * ii];
                        for (int v = 0; v < nv; ++v) integral_contribution[v] = integral_contribution[v] + w * g[v];
                    }
                /* multiply_add_to_node_vars!(du, alpha * 0.5, integral_contribution, ...) (dg_3d.jl:259-262) */
                for (int v = 0; v < nv; ++v)
                    du[nv * node + v] = du[nv * node + v] + (alpha * 0.5) * integral_contribution[v];
            }
}


/* ---- VolumeIntegralShockCapturingHG -------------------------------------------------------------- */
/* density_pressure / density / pressure (compressible_euler_3d.jl:1937-1956, compressible_euler_2d.jl analogues) */
static inline double indicator_variable(const trixi_b200_desc *d, const eqn_t *eq, const double *u) {
    int nd = d->ndims;
    if (d->equation == TRIXI_B200_EQ_MHD_3D) {
        /* density, pressure, density_pressure of the ideal GLM-MHD equations (ideal_glm_mhd_3d.jl:1318-1347) */
        double rho = u[0], mom2 = u[1] * u[1] + u[2] * u[2] + u[3] * u[3];
        double mag = u[5] * u[5] + u[6] * u[6] + u[7] * u[7], psi2 = u[8] * u[8];
        switch (d->indicator_variable) {
        case TRIXI_B200_INDVAR_DENSITY: return rho;
        case TRIXI_B200_INDVAR_PRESSURE: return (eq->gamma - 1) * (u[4] - 0.5 * (mom2 / rho + mag + psi2));
        default: return (eq->gamma - 1) * (rho * u[4] - 0.5 * (mom2 + rho * (mag + psi2)));
        }
    }
    double rho = u[0], rho_e = u[nd + 1], q = 0.0;
    for (int a = 0; a < nd; ++a) q = q + u[1 + a] * u[1 + a];
    switch (d->indicator_variable) {
    case TRIXI_B200_INDVAR_DENSITY: return rho;
    case TRIXI_B200_INDVAR_PRESSURE: return (eq->gamma - 1) * (rho_e - 0.5 * q / rho);
    default: return (eq->gamma - 1) * (rho * rho_e - 0.5 * q);
    }
}

/* calc_indicator_hennemann_gassner! (dgsem_tree/indicators_3d.jl:41-131, indicators_2d.jl:26-98) for one element */
static double indicator_hg_element(const trixi_b200_desc *d, const eqn_t *eq, const double *u, double threshold,
                                   double parameter_s) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1, nn = ipow(n, nd);
    const double *V = d->inverse_vandermonde_legendre; /* [n, n] column-major */
    double ind[512], tmp1[512], tmp2[512], modal_buf[512];
    double *modal = modal_buf;
    for (int q = 0; q < nn; ++q) ind[q] = indicator_variable(d, eq, u + (int64_t)nv * q);
    /* multiply_scalar_dimensionwise! (interpolation.jl:207-234, 348-389): x, then y, then z */
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double res = 0.0;
                for (int ii = 0; ii < n; ++ii) res = res + V[i + n * ii] * ind[ii + n * (j + n * k)];
                tmp1[i + n * (j + n * k)] = res;
            }
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double res = 0.0;
                for (int jj = 0; jj < n; ++jj) res = res + V[j + n * jj] * tmp1[i + n * (jj + n * k)];
                tmp2[i + n * (j + n * k)] = res;
            }
    if (nd == 3) {
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    double res = 0.0;
                    for (int kk = 0; kk < n; ++kk) res = res + V[k + n * kk] * tmp2[i + n * (j + n * kk)];
                    modal[i + n * (j + n * k)] = res;
                }
    } else
        modal = tmp2;
#define M3(i, j, k) modal[(i) + n * ((j) + n * (k))]
    double clip2 = 0.0, clip1, total;
    if (nd == 3) {
        for (int k = 0; k < n - 2; ++k)
            for (int j = 0; j < n - 2; ++j)
                for (int i = 0; i < n - 2; ++i) clip2 += M3(i, j, k) * M3(i, j, k);
        clip1 = clip2;
        for (int j = 0; j < n - 1; ++j)
            for (int i = 0; i < n - 1; ++i) clip1 += M3(i, j, n - 2) * M3(i, j, n - 2);
        for (int k = 0; k < n - 2; ++k)
            for (int i = 0; i < n - 1; ++i) clip1 += M3(i, n - 2, k) * M3(i, n - 2, k);
        for (int k = 0; k < n - 2; ++k)
            for (int j = 0; j < n - 2; ++j) clip1 += M3(n - 2, j, k) * M3(n - 2, j, k);
        total = clip1;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) total += M3(i, j, n - 1) * M3(i, j, n - 1);
        for (int k = 0; k < n - 1; ++k)
            for (int i = 0; i < n; ++i) total += M3(i, n - 1, k) * M3(i, n - 1, k);
        for (int k = 0; k < n - 1; ++k)
            for (int j = 0; j < n - 1; ++j) total += M3(n - 1, j, k) * M3(n - 1, j, k);
    } else {
        for (int j = 0; j < n - 2; ++j)
            for (int i = 0; i < n - 2; ++i) clip2 += M3(i, j, 0) * M3(i, j, 0);
        clip1 = clip2;
        for (int i = 0; i < n - 1; ++i) clip1 += M3(i, n - 2, 0) * M3(i, n - 2, 0);
        for (int j = 0; j < n - 2; ++j) clip1 += M3(n - 2, j, 0) * M3(n - 2, j, 0);
        total = clip1;
        for (int i = 0; i < n; ++i) total += M3(i, n - 1, 0) * M3(i, n - 1, 0);
        for (int j = 0; j < n - 1; ++j) total += M3(n - 1, j, 0) * M3(n - 1, j, 0);
    }
#undef M3
    double frac1 = total != 0.0 ? (total - clip1) / total : 0.0;
    double frac2 = clip1 != 0.0 ? (clip1 - clip2) / clip1 : 0.0;
    double energy = frac1 > frac2 ? frac1 : frac2;
    double alpha = 1 / (1 + exp(-parameter_s / threshold * (energy - threshold)));
    if (alpha < d->indicator_alpha_min) alpha = 0.0;
    if (alpha > 1 - d->indicator_alpha_min) alpha = 1.0;
    return alpha < d->indicator_alpha_max ? alpha : d->indicator_alpha_max;
}

static inline double max3(double a, double b, double c) {
    double m = a > b ? a : b;
    return m > c ? m : c;
}

/* (indicator_hg::IndicatorHennemannGassner)(u, mesh, equations, dg, cache) (dgsem/indicators.jl:114-148) with
 * apply_smoothing! (indicators_3d.jl:133-186, indicators_2d.jl:101-138); alpha [nelements] */
void oracle_calc_indicator_hg(const trixi_b200_desc *d, double *alpha, const double *u) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims;
    int64_t esz = (int64_t)d->nvars * ipow(n, nd);
    double threshold = 0.5 * pow(10.0, -1.8 * pow((double)n, 0.25));
    double parameter_s = log((1 - 0.0001) / 0.0001);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e) alpha[e] = indicator_hg_element(d, &eq, u + e * esz, threshold, parameter_s);
    if (!d->indicator_alpha_smooth) return;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(d->nelements > 0 ? d->nelements : 1));
    memcpy(tmp, alpha, sizeof(double) * (size_t)d->nelements);
    if (d->mesh_kind == TRIXI_B200_MESH_STRUCTURED) {
        /* apply_smoothing! dgsem_structured/indicators_2d.jl:8-37, indicators_3d.jl:8-40: over the elements and
         * their left neighbours (periodic meshes only, as the reference asserts) */
        for (int64_t e = 0; e < d->nelements; ++e)
            for (int a = 0; a < nd; ++a) {
                int64_t l = d->left_neighbors[a + (int64_t)nd * e] - 1;
                if (l < 0) continue;
                alpha[l] = max3(tmp[l], 0.5 * tmp[e], alpha[l]);
                alpha[e] = max3(tmp[e], 0.5 * tmp[l], alpha[e]);
            }
    }
    for (int64_t I = 0; I < d->ninterfaces; ++I) {
        int64_t l = d->interface_neighbor_ids[2 * I] - 1, r = d->interface_neighbor_ids[2 * I + 1] - 1;
        alpha[l] = max3(tmp[l], 0.5 * tmp[r], alpha[l]);
        alpha[r] = max3(tmp[r], 0.5 * tmp[l], alpha[r]);
    }
    int ns = 1 << (nd - 1);
    for (int64_t M = 0; M < d->nmortars; ++M) {
        const int64_t *ids = d->mortar_neighbor_ids + (int64_t)(ns + 1) * M;
        int64_t large = ids[ns] - 1;
        for (int q = 0; q < ns; ++q) {
            int64_t sm = ids[q] - 1;
            alpha[sm] = max3(tmp[sm], 0.5 * tmp[large], alpha[sm]);
        }
        for (int q = 0; q < ns; ++q) alpha[large] = max3(tmp[large], 0.5 * tmp[ids[q] - 1], alpha[large]);
    }
    free(tmp);
}

/* fv_kernel! + calcflux_fv! (dg_3d.jl:268-306,352-384; dg_2d.jl analogues): first-order subcell finite volumes,
 * fstar = 0 on the element boundary */
static void fv_kernel(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u, double alpha) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1;
    const double *iw = d->inverse_weights;
    int stride[3] = {1, n, n * n};
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                int idx[3] = {i, j, k};
                int64_t node = i + n * (j + n * k);
                double sum[MAXV] = {0};
                for (int a = 0; a < nd; ++a) {
                    double fl[MAXV] = {0}, fr[MAXV] = {0}; /* fstar_R[idx], fstar_L[idx + 1] */
                    if (idx[a] > 0) numflux(eq, d->volume_flux_fv, u + nv * (node - stride[a]), u + nv * node, a, fl);
                    if (idx[a] < n - 1) numflux(eq, d->volume_flux_fv, u + nv * node, u + nv * (node + stride[a]), a, fr);
                    if (flux_has_noncons(d->volume_flux_fv)) {
                        /* calcflux_fv! with nonconservative terms (dg_3d.jl:391-452): fstar_R[idx] = flux + 0.5
                         * g(u_rr, u_ll), fstar_L[idx + 1] = flux + 0.5 g(u_ll, u_rr): the node's own state comes first */
                        double g[MAXV];
                        if (idx[a] > 0) {
                            mhd_noncons_powell(eq, u + nv * node, u + nv * (node - stride[a]), a, g);
                            for (int v = 0; v < nv; ++v) fl[v] = fl[v] + 0.5 * g[v];
                        }
                        if (idx[a] < n - 1) {
                            mhd_noncons_powell(eq, u + nv * node, u + nv * (node + stride[a]), a, g);
                            for (int v = 0; v < nv; ++v) fr[v] = fr[v] + 0.5 * g[v];
                        }
                    }
                    for (int v = 0; v < nv; ++v) sum[v] = sum[v] + iw[idx[a]] * (fr[v] - fl[v]);
                }
                for (int v = 0; v < nv; ++v) du[nv * node + v] = du[nv * node + v] + alpha * sum[v];
            }
}

/* calc_volume_integral! calc_volume_integral.jl:180-191 (+ dispatch :11-33); shock capturing :231-272 */
void oracle_calc_volume_integral(const trixi_b200_desc *d, double *du, const double *u) {
    eqn_t eq = make_eqn(d);
    int64_t esz = (int64_t)d->nvars * ipow(d->nnodes, d->ndims);
    if (d->volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG) {
        double *alpha = (double *)malloc(sizeof(double) * (size_t)(d->nelements > 0 ? d->nelements : 1));
        oracle_calc_indicator_hg(d, alpha, u);
        const double atol = 1.8189894035458565e-12; /* max(100 eps, eps^0.75) for Float64 */
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < d->nelements; ++e) {
            const int noncons = flux_has_noncons(d->volume_flux);
            if (fabs(alpha[e]) <= atol) { /* isapprox(alpha, 0, atol): pure DG */
                flux_differencing_kernel(d, &eq, du + e * esz, u + e * esz, 1.0);
                if (noncons) flux_differencing_noncons(d, &eq, du + e * esz, u + e * esz, 1.0);
            } else {
                flux_differencing_kernel(d, &eq, du + e * esz, u + e * esz, 1 - alpha[e]);
                if (noncons) flux_differencing_noncons(d, &eq, du + e * esz, u + e * esz, 1 - alpha[e]);
                fv_kernel(d, &eq, du + e * esz, u + e * esz, alpha[e]);
            }
        }
        free(alpha);
        return;
    }
    if (d->volume_integral == TRIXI_B200_VOLINT_PURE_LGL_FV) { /* fv_kernel! with alpha = true on every element */
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < d->nelements; ++e) fv_kernel(d, &eq, du + e * esz, u + e * esz, 1.0);
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e) {
        if (d->volume_integral == TRIXI_B200_VOLINT_WEAK_FORM)
            weak_form_kernel(d, &eq, du + e * esz, u + e * esz);
        else if (d->volume_flux == TRIXI_B200_FLUX_RANOCHA_TURBO && d->equation == TRIXI_B200_EQ_EULER_3D &&
                 d->nnodes == 4 && d->mesh_kind == TRIXI_B200_MESH_TREE)
            /* the reference's performance specialization (dispatch on typeof(flux_ranocha_turbo)) */
            flux_differencing_kernel_turbo(d, &eq, du + e * esz, u + e * esz, 1.0);
        else {
            flux_differencing_kernel(d, &eq, du + e * esz, u + e * esz, 1.0);
            if (flux_has_noncons(d->volume_flux)) flux_differencing_noncons(d, &eq, du + e * esz, u + e * esz, 1.0);
        }
    }
}

/* prolong2interfaces! dg_3d.jl:530-567 / dg_2d.jl:515-541; interfaces_u[2, nv, nf, I] */
void oracle_prolong2interfaces(const trixi_b200_desc *d, double *iu, const double *u) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->ninterfaces; ++I) {
        int64_t left = d->interface_neighbor_ids[2 * I] - 1, right = d->interface_neighbor_ids[2 * I + 1] - 1;
        int o = (int)d->interface_orientations[I] - 1;
        for (int b = 0; b < nb; ++b)
            for (int a = 0; a < n; ++a) {
                int fn = a + n * b;
                int nl = face_to_volume_node(nd, n, o, n - 1, a, b);
                int nr = face_to_volume_node(nd, n, o, 0, a, b);
                for (int v = 0; v < nv; ++v) {
                    iu[0 + 2 * (v + nv * (fn + (int64_t)nf * I))] = u[left * esz + nv * nl + v];
                    iu[1 + 2 * (v + nv * (fn + (int64_t)nf * I))] = u[right * esz + nv * nr + v];
                }
            }
    }
}

/* calc_interface_flux! dg_3d.jl:569-602 / dg_2d.jl:543-576; sfv[nv, nf, 2*nd, nelem] */
void oracle_calc_interface_flux(const trixi_b200_desc *d, double *sfv, const double *iu) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1);
    int64_t fsz = (int64_t)nv * nf * 2 * nd;
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->ninterfaces; ++I) {
        int64_t left = d->interface_neighbor_ids[2 * I] - 1, right = d->interface_neighbor_ids[2 * I + 1] - 1;
        int o = (int)d->interface_orientations[I] - 1;
        int left_direction = 2 * (o + 1) - 1, right_direction = 2 * (o + 1) - 2; /* 0-based directions */
        for (int fn = 0; fn < nf; ++fn) {
            double ul[MAXV], ur[MAXV], f[MAXV];
            for (int v = 0; v < nv; ++v) {
                ul[v] = iu[0 + 2 * (v + nv * (fn + (int64_t)nf * I))];
                ur[v] = iu[1 + 2 * (v + nv * (fn + (int64_t)nf * I))];
            }
            numflux(&eq, d->surface_flux, ul, ur, o, f);
            if (flux_has_noncons(d->surface_flux)) { /* dg_3d.jl:604-649 */
                double nl[MAXV], nr[MAXV];
                mhd_noncons_powell(&eq, ul, ur, o, nl);
                mhd_noncons_powell(&eq, ur, ul, o, nr);
                for (int v = 0; v < nv; ++v) {
                    sfv[left * fsz + v + nv * (fn + nf * left_direction)] = f[v] + 0.5 * nl[v];
                    sfv[right * fsz + v + nv * (fn + nf * right_direction)] = f[v] + 0.5 * nr[v];
                }
                continue;
            }
            for (int v = 0; v < nv; ++v) {
                sfv[left * fsz + v + nv * (fn + nf * left_direction)] = f[v];
                sfv[right * fsz + v + nv * (fn + nf * right_direction)] = f[v];
            }
        }
    }
}

/* prolong2boundaries! dg_3d.jl:651-701; boundaries_u[2, nv, nf, B] */
void oracle_prolong2boundaries(const trixi_b200_desc *d, double *bu, const double *u) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd);
#pragma omp parallel for schedule(static)
    for (int64_t B = 0; B < d->nboundaries; ++B) {
        int64_t element = d->boundary_neighbor_ids[B] - 1;
        int o = (int)d->boundary_orientations[B] - 1;
        int side = (int)d->boundary_neighbor_sides[B];
        for (int b = 0; b < nb; ++b)
            for (int a = 0; a < n; ++a) {
                int fn = a + n * b;
                int vn = face_to_volume_node(nd, n, o, side == 1 ? n - 1 : 0, a, b);
                for (int v = 0; v < nv; ++v)
                    bu[(side == 1 ? 0 : 1) + 2 * (v + nv * (fn + (int64_t)nf * B))] = u[element * esz + nv * vn + v];
            }
    }
}

/* calc_boundary_flux! dg_3d.jl:703-768 */
void oracle_calc_boundary_flux(const trixi_b200_desc *d, double *sfv, const double *bu, double t) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1);
    int64_t fsz = (int64_t)nv * nf * 2 * nd;
    int64_t first = 0;
    for (int direction = 1; direction <= 2 * nd; ++direction) {
        int64_t cnt = d->n_boundaries_per_direction[direction - 1];
        int bc = d->boundary_conditions[direction - 1], ic = d->boundary_ic[direction - 1];
#pragma omp parallel for schedule(static)
        for (int64_t B = first; B < first + cnt; ++B) {
            int64_t neighbor = d->boundary_neighbor_ids[B] - 1;
            int o = (int)d->boundary_orientations[B] - 1;
            int side = (int)d->boundary_neighbor_sides[B];
            for (int fn = 0; fn < nf; ++fn) {
                double ui[MAXV], f[MAXV];
                for (int v = 0; v < nv; ++v) ui[v] = bu[(side == 1 ? 0 : 1) + 2 * (v + nv * (fn + (int64_t)nf * B))];
                const double *x = d->boundary_node_coordinates + (int64_t)nd * (fn + (int64_t)nf * B);
                boundary_flux(&eq, bc, ic, d->surface_flux, ui, o, direction, x, t, f);
                for (int v = 0; v < nv; ++v) sfv[neighbor * fsz + v + nv * (fn + nf * (direction - 1))] = f[v];
            }
        }
        first += cnt;
    }
}

/* prolong2mpiinterfaces! dgsem_tree/dg_2d_parallel.jl:565-598 (p4est: dg_3d_parallel.jl:119-165):
 * the local side of mpi_interfaces_u[2, nv, nf, MI] is filled from u; the remote side arrives through
 * the halo exchange (dg_parallel.jl:66-182) */
void oracle_prolong2mpiinterfaces(const trixi_b200_desc *d, double *mu, const double *u) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->nmpiinterfaces; ++I) {
        int64_t element = d->mpi_local_neighbor_ids[I] - 1;
        int o = (int)d->mpi_orientations[I] - 1;
        int side = (int)d->mpi_local_sides[I]; /* 1: local element is the left (-) one */
        for (int b = 0; b < nb; ++b)
            for (int a = 0; a < n; ++a) {
                int fn = a + n * b;
                int vn = face_to_volume_node(nd, n, o, side == 1 ? n - 1 : 0, a, b);
                for (int v = 0; v < nv; ++v)
                    mu[(side - 1) + 2 * (v + nv * (fn + (int64_t)nf * I))] = u[element * esz + nv * vn + v];
            }
    }
}

/* calc_mpi_interface_flux! dgsem_tree/dg_2d_parallel.jl:700-740 (p4est: dg_3d_parallel.jl:167-242): the
 * shared flux is computed redundantly on both ranks, only the local element's storage is written */
void oracle_calc_mpi_interface_flux(const trixi_b200_desc *d, double *sfv, const double *mu) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1);
    int64_t fsz = (int64_t)nv * nf * 2 * nd;
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->nmpiinterfaces; ++I) {
        if (d->mpi_is_mortar_piece && d->mpi_is_mortar_piece[I]) continue; /* exchange only: see MPI mortars */
        int64_t element = d->mpi_local_neighbor_ids[I] - 1;
        int o = (int)d->mpi_orientations[I] - 1;
        int side = (int)d->mpi_local_sides[I];
        /* left element stores in direction 2*orientation, right element in 2*orientation - 1 (1-based) */
        int direction0 = side == 1 ? 2 * o + 1 : 2 * o;
        for (int fn = 0; fn < nf; ++fn) {
            double ul[MAXV], ur[MAXV], f[MAXV];
            for (int v = 0; v < nv; ++v) {
                ul[v] = mu[0 + 2 * (v + nv * (fn + (int64_t)nf * I))];
                ur[v] = mu[1 + 2 * (v + nv * (fn + (int64_t)nf * I))];
            }
            numflux(&eq, d->surface_flux, ul, ur, o, f);
            if (flux_has_noncons(d->surface_flux)) { /* dgsem_p4est/dg_3d_parallel.jl:302-337 (TreeMesh MPI is 2D-only upstream) */
                double g[MAXV];
                if (side == 1)
                    mhd_noncons_powell(&eq, ul, ur, o, g);
                else
                    mhd_noncons_powell(&eq, ur, ul, o, g);
                for (int v = 0; v < nv; ++v) f[v] = f[v] + 0.5 * g[v];
            }
            for (int v = 0; v < nv; ++v) sfv[element * fsz + v + nv * (fn + nf * direction0)] = f[v];
        }
    }
}

/* calc_surface_integral! dg_3d.jl:1337-1394 / dg_2d.jl:1245-1300 */
void oracle_calc_surface_integral(const trixi_b200_desc *d, double *du, const double *sfv) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd), fsz = (int64_t)nv * nf * 2 * nd;
    double factor = d->inverse_weights[0];
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e) {
        double *due = du + e * esz;
        const double *s = sfv + e * fsz;
        for (int m = 0; m < nb; ++m)
            for (int l = 0; l < n; ++l) {
                int fn = l + n * m;
                for (int v = 0; v < nv; ++v) {
                    for (int o = 0; o < nd; ++o) {
                        int lo = face_to_volume_node(nd, n, o, 0, l, m);
                        int hi = face_to_volume_node(nd, n, o, n - 1, l, m);
                        due[nv * lo + v] = due[nv * lo + v] - s[v + nv * (fn + nf * (2 * o))] * factor;
                        due[nv * hi + v] = due[nv * hi + v] + s[v + nv * (fn + nf * (2 * o + 1))] * factor;
                    }
                }
            }
    }
}

/* apply_jacobian! dg_3d.jl:1396-1414 */
void oracle_apply_jacobian(const trixi_b200_desc *d, double *du) {
    int64_t esz = (int64_t)d->nvars * ipow(d->nnodes, d->ndims);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e) {
        double factor = -d->inverse_jacobian[e];
        for (int64_t i = 0; i < esz; ++i) du[e * esz + i] *= factor;
    }
}

/* calc_sources! dg_3d.jl:1417-1437 */
void oracle_calc_sources(const trixi_b200_desc *d, double *du, const double *u, double t) {
    if (d->source_terms == TRIXI_B200_SRC_NONE) return;
    eqn_t eq = make_eqn(d);
    int nv = d->nvars, nd = d->ndims;
    int64_t nn = ipow(d->nnodes, nd);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e)
        for (int64_t q = 0; q < nn; ++q) {
            double s[MAXV];
            source_terms(&eq, d->source_terms, u + nv * (q + nn * e), d->node_coordinates + nd * (q + nn * e), t, s);
            for (int v = 0; v < nv; ++v) du[nv * (q + nn * e) + v] += s[v];
        }
}

/* ---- curved meshes: src/solvers/dgsem_structured/ ------------------------------------------------ */
/* contravariant vector Ja^index at a node (dgsem_structured/dg.jl:27-30); storage [dim, index, node, elem] */
static inline void get_contravariant_vector(const trixi_b200_desc *d, int index, int64_t node, int64_t e,
                                            double *ja) {
    int nd = d->ndims;
    int64_t nn = ipow(d->nnodes, nd);
    const double *p = d->contravariant_vectors + (int64_t)nd * nd * (node + nn * e) + (int64_t)nd * index;
    for (int dim = 0; dim < nd; ++dim) ja[dim] = p[dim];
}

/* weak_form_kernel! dgsem_structured/dg_3d.jl:36-89 / dg_2d.jl:85-124 */
static void weak_form_kernel_curved(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u,
                                    int64_t e) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1;
    const double *Dhat = d->derivative_hat;
    int stride[3] = {1, n, n * n};
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                int idx[3] = {i, j, k};
                int64_t node = i + n * (j + n * k);
                const double *un = u + (int64_t)nv * node;
                double fl[3][MAXV];
                for (int o = 0; o < nd; ++o) phys_flux(eq, un, o, fl[o]);
                for (int a = 0; a < nd; ++a) {
                    double ja[3], cf[MAXV];
                    get_contravariant_vector(d, a, node, e, ja);
                    for (int v = 0; v < nv; ++v) {
                        double s = ja[0] * fl[0][v] + ja[1] * fl[1][v];
                        if (nd == 3) s += ja[2] * fl[2][v];
                        cf[v] = s;
                    }
                    for (int ii = 0; ii < n; ++ii) {
                        double w = Dhat[ii + n * idx[a]];
                        double *t = du + (int64_t)nv * (node + (ii - idx[a]) * stride[a]);
                        for (int v = 0; v < nv; ++v) t[v] = t[v] + w * cf[v];
                    }
                }
            }
}

/* flux_differencing_kernel! dgsem_structured/dg_3d.jl:94-175 / dg_2d.jl:126-190 */
static void flux_differencing_kernel_curved(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u,
                                            int64_t e, double alpha) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1;
    const double *Ds = d->derivative_split;
    int stride[3] = {1, n, n * n};
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                int idx[3] = {i, j, k};
                int64_t node = i + n * (j + n * k);
                const double *un = u + nv * node;
                for (int a = 0; a < nd; ++a) {
                    double ja_node[3];
                    get_contravariant_vector(d, a, node, e, ja_node);
                    for (int ii = idx[a] + 1; ii < n; ++ii) {
                        int64_t node2 = node + (ii - idx[a]) * stride[a];
                        double ja2[3], ja_avg[3], f[MAXV];
                        get_contravariant_vector(d, a, node2, e, ja2);
                        for (int dim = 0; dim < nd; ++dim) ja_avg[dim] = 0.5 * (ja_node[dim] + ja2[dim]);
                        numflux_normal(eq, d->volume_flux, un, u + nv * node2, ja_avg, f);
                        double w1 = alpha * Ds[idx[a] + n * ii], w2 = alpha * Ds[ii + n * idx[a]];
                        for (int v = 0; v < nv; ++v) du[nv * node + v] = du[nv * node + v] + w1 * f[v];
                        for (int v = 0; v < nv; ++v) du[nv * node2 + v] = du[nv * node2 + v] + w2 * f[v];
                    }
                }
            }
}

/* nonconservative volume terms on curved meshes (dgsem_structured/dg_3d.jl:177-283): for every node
 * 0.5 sum_d sum_ii D_split[i, ii] g(u_node, u_ii, 0.5 (Ja^d_node + Ja^d_ii)) */
static void flux_differencing_noncons_curved(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u,
                                             int64_t e, double alpha) {
    int n = d->nnodes, nv = d->nvars;
    const double *Ds = d->derivative_split;
    int stride[3] = {1, n, n * n};
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                int idx[3] = {i, j, k};
                int64_t node = i + n * (j + n * k);
                const double *un = u + nv * node;
                double integral_contribution[MAXV] = {0};
                for (int a = 0; a < 3; ++a) {
                    double ja_node[3];
                    get_contravariant_vector(d, a, node, e, ja_node);
                    for (int ii = 0; ii < n; ++ii) {
                        int64_t node2 = node + (ii - idx[a]) * stride[a];
                        double ja2[3], ja_avg[3], g[MAXV];
                        get_contravariant_vector(d, a, node2, e, ja2);
                        for (int dim = 0; dim < 3; ++dim) ja_avg[dim] = 0.5 * (ja_node[dim] + ja2[dim]);
                        mhd_noncons_powell_normal(eq, un, u + nv * node2, ja_avg, g);
                        double w = Ds[idx[a] + n * ii];
                        for (int v = 0; v < nv; ++v) integral_contribution[v] = integral_contribution[v] + w * g[v];
                    }
                }
                for (int v = 0; v < nv; ++v)
                    du[nv * node + v] = du[nv * node + v] + (alpha * 0.5) * integral_contribution[v];
            }
}

/* fv_kernel! (dg_3d.jl:268-306, shared by all meshes) with calcflux_fv! for curved meshes (dgsem_structured/
 * dg_2d.jl, dg_3d.jl:377-436): first-order subcell finite volumes along the precomputed free-stream preserving
 * normal vectors (NormalVectorContainer, containers_3d.jl:352-541); fstar = 0 on the element boundary */
static void fv_kernel_curved(const trixi_b200_desc *d, const eqn_t *eq, double *du, const double *u, int64_t e,
                             double alpha) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1;
    const double *iw = d->inverse_weights;
    int stride[3] = {1, n, n * n};
    for (int k = 0; k < n3; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                int idx[3] = {i, j, k};
                int64_t node = i + n * (j + n * k);
                double sum[MAXV] = {0};
                for (int a = 0; a < nd; ++a) {
                    /* normal_vectors_a [nd, dims.., nelements] with n - 1 entries along direction a */
                    int dims[3] = {n, n, n3};
                    dims[a] = n - 1;
                    int64_t per_elem = (int64_t)dims[0] * dims[1] * (nd == 3 ? dims[2] : 1);
                    const double *nvec = d->subcell_normal_vectors[a] + (int64_t)nd * per_elem * e;
                    double fl[MAXV] = {0}, fr[MAXV] = {0}; /* fstar_R[idx], fstar_L[idx + 1] */
                    for (int side = 0; side < 2; ++side) {
                        int pos[3] = {i, j, k};
                        if (side == 0) {
                            if (idx[a] == 0) continue;
                            pos[a] = idx[a] - 1;
                        } else if (idx[a] == n - 1)
                            continue;
                        int64_t q = pos[0] + (int64_t)dims[0] * (pos[1] + (int64_t)dims[1] * pos[2]);
                        double nrm[3] = {0, 0, 0};
                        for (int c = 0; c < nd; ++c) nrm[c] = nvec[c + nd * q];
                        if (side == 0)
                            numflux_normal(eq, d->volume_flux_fv, u + nv * (node - stride[a]), u + nv * node, nrm, fl);
                        else
                            numflux_normal(eq, d->volume_flux_fv, u + nv * node, u + nv * (node + stride[a]), nrm, fr);
                        if (flux_has_noncons(d->volume_flux_fv)) {
                            /* calcflux_fv! with nonconservative terms (dgsem_structured/dg_3d.jl:438-530): ftilde_R =
                             * ftilde + 0.5 g(u_rr, u_ll, n), ftilde_L = ftilde + 0.5 g(u_ll, u_rr, n) */
                            double g[MAXV];
                            double *f = side == 0 ? fl : fr;
                            mhd_noncons_powell_normal(eq, u + nv * node,
                                                      u + nv * (side == 0 ? node - stride[a] : node + stride[a]), nrm, g);
                            for (int v = 0; v < nv; ++v) f[v] = f[v] + 0.5 * g[v];
                        }
                    }
                    for (int v = 0; v < nv; ++v) sum[v] = sum[v] + iw[idx[a]] * (fr[v] - fl[v]);
                }
                for (int v = 0; v < nv; ++v) du[nv * node + v] = du[nv * node + v] + alpha * sum[v];
            }
}

void oracle_calc_indicator_hg(const trixi_b200_desc *d, double *alpha, const double *u);
void oracle_calc_volume_integral_curved(const trixi_b200_desc *d, double *du, const double *u) {
    eqn_t eq = make_eqn(d);
    int64_t esz = (int64_t)d->nvars * ipow(d->nnodes, d->ndims);
    if (d->volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG) { /* calc_volume_integral.jl:231-272 */
        double *alpha = (double *)malloc(sizeof(double) * (size_t)(d->nelements > 0 ? d->nelements : 1));
        oracle_calc_indicator_hg(d, alpha, u);
        const double atol = 1.8189894035458565e-12; /* max(100 eps, eps^0.75) for Float64 */
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < d->nelements; ++e) {
            const int noncons = flux_has_noncons(d->volume_flux);
            if (fabs(alpha[e]) <= atol) {
                flux_differencing_kernel_curved(d, &eq, du + e * esz, u + e * esz, e, 1.0);
                if (noncons) flux_differencing_noncons_curved(d, &eq, du + e * esz, u + e * esz, e, 1.0);
            } else {
                flux_differencing_kernel_curved(d, &eq, du + e * esz, u + e * esz, e, 1 - alpha[e]);
                if (noncons) flux_differencing_noncons_curved(d, &eq, du + e * esz, u + e * esz, e, 1 - alpha[e]);
                fv_kernel_curved(d, &eq, du + e * esz, u + e * esz, e, alpha[e]);
            }
        }
        free(alpha);
        return;
    }
    if (d->volume_integral == TRIXI_B200_VOLINT_PURE_LGL_FV) {
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < d->nelements; ++e) fv_kernel_curved(d, &eq, du + e * esz, u + e * esz, e, 1.0);
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e) {
        if (d->volume_integral == TRIXI_B200_VOLINT_WEAK_FORM)
            weak_form_kernel_curved(d, &eq, du + e * esz, u + e * esz, e);
        else {
            flux_differencing_kernel_curved(d, &eq, du + e * esz, u + e * esz, e, 1.0);
            if (flux_has_noncons(d->volume_flux))
                flux_differencing_noncons_curved(d, &eq, du + e * esz, u + e * esz, e, 1.0);
        }
    }
}

/* prolong2interfaces! dgsem_structured/dg_3d.jl:619-655: interfaces_u[nv, nf, 2nd, nelem] */
void oracle_prolong2interfaces_structured(const trixi_b200_desc *d, double *iu, const double *u) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd), fsz = (int64_t)nv * nf * 2 * nd;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e)
        for (int b = 0; b < nb; ++b)
            for (int a = 0; a < n; ++a)
                for (int o = 0; o < nd; ++o)
                    for (int side = 0; side < 2; ++side) {
                        int vn = face_to_volume_node(nd, n, o, side ? n - 1 : 0, a, b);
                        for (int v = 0; v < nv; ++v)
                            iu[e * fsz + v + nv * ((a + n * b) + nf * (2 * o + side))] = u[e * esz + nv * vn + v];
                    }
}

/* calc_interface_flux! dgsem_structured/dg_3d.jl:657-753 */
void oracle_calc_interface_flux_structured(const trixi_b200_desc *d, double *sfv, const double *iu) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t nn = ipow(n, nd), fsz = (int64_t)nv * nf * 2 * nd;
#pragma omp parallel for schedule(static)
    for (int64_t right = 0; right < d->nelements; ++right)
        for (int o = 0; o < nd; ++o) {
            int64_t left = d->left_neighbors[o + nd * right] - 1;
            if (left < 0) continue;
            int right_direction = 2 * o + 1, left_direction = 2 * o; /* 0-based */
            for (int b = 0; b < nb; ++b)
                for (int a = 0; a < n; ++a) {
                    int fn = a + n * b;
                    double ul[MAXV], ur[MAXV], f[MAXV], ja[3], nrm[3] = {0, 0, 0};
                    for (int v = 0; v < nv; ++v) {
                        ul[v] = iu[left * fsz + v + nv * (fn + nf * right_direction)];
                        ur[v] = iu[right * fsz + v + nv * (fn + nf * left_direction)];
                    }
                    int vn = face_to_volume_node(nd, n, o, 0, a, b); /* first layer of the right element */
                    double ij = d->inverse_jacobian[vn + nn * right];
                    double sign_jacobian = (ij > 0) - (ij < 0);
                    get_contravariant_vector(d, o, vn, right, ja);
                    for (int dim = 0; dim < nd; ++dim) nrm[dim] = sign_jacobian * ja[dim];
                    numflux_normal(&eq, d->surface_flux, ul, ur, nrm, f);
                    if (flux_has_noncons(d->surface_flux)) { /* dgsem_structured/dg_3d.jl:755-828 */
                        double gl[MAXV], gr[MAXV];
                        mhd_noncons_powell_normal(&eq, ul, ur, nrm, gl);
                        mhd_noncons_powell_normal(&eq, ur, ul, nrm, gr);
                        for (int v = 0; v < nv; ++v) {
                            double fv = sign_jacobian * f[v];
                            sfv[left * fsz + v + nv * (fn + nf * right_direction)] = fv + 0.5 * (sign_jacobian * gl[v]);
                            sfv[right * fsz + v + nv * (fn + nf * left_direction)] = fv + 0.5 * (sign_jacobian * gr[v]);
                        }
                        continue;
                    }
                    for (int v = 0; v < nv; ++v) {
                        double fv = sign_jacobian * f[v];
                        sfv[left * fsz + v + nv * (fn + nf * right_direction)] = fv;
                        sfv[right * fsz + v + nv * (fn + nf * left_direction)] = fv;
                    }
                }
        }
}

/* calc_boundary_flux! dgsem_structured/dg_3d.jl:755-935 + calc_boundary_flux_by_direction! dg.jl:124-165;
 * the boundary faces come as the direction-sorted list of the descriptor */
void oracle_calc_boundary_flux_structured(const trixi_b200_desc *d, double *sfv, const double *iu, double t) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t nn = ipow(n, nd), fsz = (int64_t)nv * nf * 2 * nd;
    int64_t first = 0;
    for (int direction = 1; direction <= 2 * nd; ++direction) {
        int64_t cnt = d->n_boundaries_per_direction[direction - 1];
        int bc = d->boundary_conditions[direction - 1], ic = d->boundary_ic[direction - 1];
        int o = (direction - 1) / 2;
#pragma omp parallel for schedule(static)
        for (int64_t B = first; B < first + cnt; ++B) {
            int64_t e = d->boundary_neighbor_ids[B] - 1;
            for (int b = 0; b < nb; ++b)
                for (int a = 0; a < n; ++a) {
                    int fn = a + n * b;
                    int vn = face_to_volume_node(nd, n, o, direction % 2 == 1 ? 0 : n - 1, a, b);
                    double ui[MAXV], f[MAXV], ja[3], nrm[3] = {0, 0, 0};
                    for (int v = 0; v < nv; ++v) ui[v] = iu[e * fsz + v + nv * (fn + nf * (direction - 1))];
                    const double *x = d->node_coordinates + (int64_t)nd * (vn + nn * e);
                    double ij = d->inverse_jacobian[vn + nn * e];
                    double sign_jacobian = (ij > 0) - (ij < 0);
                    get_contravariant_vector(d, o, vn, e, ja);
                    for (int dim = 0; dim < nd; ++dim) nrm[dim] = sign_jacobian * ja[dim];
                    boundary_flux_normal(&eq, bc, ic, d->surface_flux, ui, nrm, direction, x, t, f);
                    for (int v = 0; v < nv; ++v) sfv[e * fsz + v + nv * (fn + nf * (direction - 1))] = sign_jacobian * f[v];
                }
        }
        first += cnt;
    }
}

/* apply_jacobian! dgsem_structured/dg_3d.jl:937-956 */
void oracle_apply_jacobian_curved(const trixi_b200_desc *d, double *du) {
    int nv = d->nvars;
    int64_t nn = ipow(d->nnodes, d->ndims);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e)
        for (int64_t q = 0; q < nn; ++q) {
            double factor = -d->inverse_jacobian[q + nn * e];
            for (int v = 0; v < nv; ++v) du[nv * (q + nn * e) + v] *= factor;
        }
}

/* rhs_hyperbolic! dgsem_structured/dg.jl:41-94; interfaces_u doubles as the work array [nv,nf,2nd,nelem] */
void oracle_rhs_structured(const trixi_b200_desc *d, double *du, const double *u, double t, double *interfaces_u,
                           double *sfv) {
    oracle_set_zero(d, du);
    oracle_calc_volume_integral_curved(d, du, u);
    oracle_prolong2interfaces_structured(d, interfaces_u, u);
    oracle_calc_interface_flux_structured(d, sfv, interfaces_u);
    if (d->nboundaries > 0) oracle_calc_boundary_flux_structured(d, sfv, interfaces_u, t);
    oracle_calc_surface_integral(d, du, sfv); /* shared with TreeMesh (dg_3d.jl:1337) */
    oracle_apply_jacobian_curved(d, du);
    oracle_calc_sources(d, du, u, t);
}

/* max_dt for curved meshes stepsize_dg3d.jl:79-123 (constant speed: :125-160), stepsize_dg2d.jl */
double oracle_max_dt_curved(const trixi_b200_desc *d, const double *u) {
    eqn_t eq = make_eqn(d);
    int nv = d->nvars, nd = d->ndims;
    int64_t nn = ipow(d->nnodes, nd);
    double max_lambda = 0.0;
    int nanflag = 0;
#pragma omp parallel for schedule(static) reduction(max : max_lambda) reduction(| : nanflag)
    for (int64_t e = 0; e < d->nelements; ++e) {
        double ml[3] = {0, 0, 0}, ml_const = 0.0;
        const int constant_speed = !is_euler(&eq) && !is_mhd(&eq); /* have_constant_speed: linear advection */
        for (int64_t q = 0; q < nn; ++q) {
            double lam[3] = {0, 0, 0};
            if (is_euler(&eq)) {
                double rho, v[3], p;
                euler_cons2prim(&eq, u + nv * (q + nn * e), &rho, v, &p);
                double c = sqrt(eq.gamma * p / rho);
                for (int dd = 0; dd < nd; ++dd) lam[dd] = fabs(v[dd]) + c;
            } else if (is_mhd(&eq)) { /* max_abs_speeds ideal_glm_mhd_3d.jl:1218-1228 */
                const double *un = u + nv * (q + nn * e);
                for (int dd = 0; dd < 3; ++dd) lam[dd] = fabs(un[1 + dd] / un[0]) + mhd_fast_wavespeed(&eq, un, dd);
            } else {
                for (int dd = 0; dd < nd; ++dd) lam[dd] = fabs(eq.a[dd]);
            }
            double inv_jacobian = fabs(d->inverse_jacobian[q + nn * e]);
            double node_sum = 0.0;
            for (int a = 0; a < nd; ++a) {
                double ja[3], s = 0.0;
                get_contravariant_vector(d, a, q, e, ja);
                for (int dim = 0; dim < nd; ++dim) s += ja[dim] * lam[dim];
                double val = inv_jacobian * fabs(s);
                if (isnan(val)) nanflag = 1;
                ml[a] = fmax(ml[a], val);
                node_sum += fabs(s);
            }
            /* constant_speed::True (max_scaled_speed_per_element, stepsize_dg3d.jl:176-210, stepsize_dg2d.jl:207-235):
             * the maximum over the nodes of inv_jacobian * (sum of the transformed speeds), not the sum of the
             * per-direction maxima */
            if (constant_speed) ml_const = fmax(ml_const, inv_jacobian * node_sum);
        }
        if (constant_speed) {
            ml[0] = ml_const;
            ml[1] = ml[2] = 0.0;
        }
        double s = 0.0;
        for (int a = 0; a < nd; ++a) s += ml[a];
        if (s > max_lambda) max_lambda = s;
    }
    if (nanflag) return NAN;
    double max_scaled_speed = fmax(DBL_TRUE_MIN, max_lambda);
    return 2 / (d->nnodes * max_scaled_speed);
}

/* ---- P4estMesh: src/solvers/dgsem_p4est/ ------------------------------------------------------------- */
/* node index along one axis from a symbolic index (index_to_start_step_3d dg_3d.jl:61-80 in closed form);
 * encoding :begin 0, :end 1, :i_forward 2, :i_backward 3, :j_forward 4, :j_backward 5 */
static inline int p4_index(int sym, int n, int i, int j) {
    switch (sym) {
    case 0: return 0;
    case 1: return n - 1;
    case 2: return i;
    case 3: return n - 1 - i;
    case 4: return j;
    default: return n - 1 - j;
    }
}
/* indices2direction dgsem_p4est/containers.jl (direction 1..6 of a face from its index tuple), 0-based here */
static inline int p4_direction(int nd, const int64_t *idx) {
    for (int c = 0; c < nd; ++c) {
        if (idx[c] == 0) return 2 * c;
        if (idx[c] == 1) return 2 * c + 1;
    }
    return -1;
}
static inline int64_t p4_volume_node(int nd, int n, const int64_t *idx, int i, int j) {
    int64_t node = 0, stride = 1;
    for (int c = 0; c < nd; ++c) {
        node += stride * p4_index((int)idx[c], n, i, j);
        stride *= n;
    }
    return node;
}
/* surface_indices dg_3d.jl:82-92: the two (3D) / one (2D) varying symbols of an index tuple */
static inline void p4_surface_node(int nd, int n, const int64_t *idx, int i, int j, int *fn) {
    int s[2] = {0, 0}, k = 0;
    for (int c = 0; c < nd; ++c)
        if (idx[c] > 1) s[k++] = p4_index((int)idx[c], n, i, j);
    *fn = nd == 3 ? s[0] + n * s[1] : s[0];
}

/* get_normal_direction dgsem_p4est/dg.jl:74-86: outward normal = +-Ja^orientation */
static inline void p4_normal(const trixi_b200_desc *d, int direction0, int64_t node, int64_t e, double *nrm) {
    double ja[3] = {0, 0, 0};
    get_contravariant_vector(d, direction0 / 2, node, e, ja);
    double sgn = direction0 % 2 == 0 ? -1.0 : 1.0;
    for (int dim = 0; dim < d->ndims; ++dim) nrm[dim] = sgn * ja[dim];
}

/* prolong2interfaces! dgsem_p4est/dg_3d.jl:94-183: interfaces_u[2, nv, nf, I], both sides stored at the
 * primary's face node (i, j) */
void oracle_prolong2interfaces_p4est(const trixi_b200_desc *d, double *iu, const double *u) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->ninterfaces; ++I)
        for (int side = 0; side < 2; ++side) {
            int64_t e = d->interface_neighbor_ids[2 * I + side] - 1;
            const int64_t *idx = d->interface_node_indices + (int64_t)nd * (side + 2 * I);
            for (int j = 0; j < nb; ++j)
                for (int i = 0; i < n; ++i) {
                    int64_t vn = p4_volume_node(nd, n, idx, i, j);
                    for (int v = 0; v < nv; ++v)
                        iu[side + 2 * (v + nv * ((i + n * j) + (int64_t)nf * I))] = u[e * esz + nv * vn + v];
                }
        }
}

/* calc_interface_flux! dgsem_p4est/dg_3d.jl:185-314 */
void oracle_calc_interface_flux_p4est(const trixi_b200_desc *d, double *sfv, const double *iu) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t fsz = (int64_t)nv * nf * 2 * nd;
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->ninterfaces; ++I) {
        int64_t primary = d->interface_neighbor_ids[2 * I] - 1, secondary = d->interface_neighbor_ids[2 * I + 1] - 1;
        const int64_t *pidx = d->interface_node_indices + (int64_t)nd * (0 + 2 * I);
        const int64_t *sidx = d->interface_node_indices + (int64_t)nd * (1 + 2 * I);
        int pdir = p4_direction(nd, pidx), sdir = p4_direction(nd, sidx);
        for (int j = 0; j < nb; ++j)
            for (int i = 0; i < n; ++i) {
                double ul[MAXV], ur[MAXV], f[MAXV], nrm[3] = {0, 0, 0};
                int fn = i + n * j, fn_sec;
                for (int v = 0; v < nv; ++v) {
                    ul[v] = iu[0 + 2 * (v + nv * (fn + (int64_t)nf * I))];
                    ur[v] = iu[1 + 2 * (v + nv * (fn + (int64_t)nf * I))];
                }
                p4_normal(d, pdir, p4_volume_node(nd, n, pidx, i, j), primary, nrm);
                numflux_normal(&eq, d->surface_flux, ul, ur, nrm, f);
                p4_surface_node(nd, n, sidx, i, j, &fn_sec);
                if (flux_has_noncons(d->surface_flux)) { /* dgsem_p4est/dg_3d.jl:340-375 */
                    double gp[MAXV], gs[MAXV];
                    mhd_noncons_powell_normal(&eq, ul, ur, nrm, gp);
                    mhd_noncons_powell_normal(&eq, ur, ul, nrm, gs);
                    for (int v = 0; v < nv; ++v) {
                        sfv[primary * fsz + v + nv * (fn + nf * pdir)] = f[v] + 0.5 * gp[v];
                        sfv[secondary * fsz + v + nv * (fn_sec + nf * sdir)] = -(f[v] + 0.5 * gs[v]);
                    }
                    continue;
                }
                for (int v = 0; v < nv; ++v) {
                    sfv[primary * fsz + v + nv * (fn + nf * pdir)] = f[v];
                    sfv[secondary * fsz + v + nv * (fn_sec + nf * sdir)] = -f[v];
                }
            }
    }
}

/* prolong2boundaries! + calc_boundary_flux! dgsem_p4est/dg_3d.jl:412-548; Dirichlet for unstructured
 * meshes (equations.jl:206-228): flux(u_inner, u_boundary, outward normal) */
void oracle_calc_boundary_flux_p4est(const trixi_b200_desc *d, double *sfv, const double *u, double t) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t nn = ipow(n, nd), esz = nv * nn, fsz = (int64_t)nv * nf * 2 * nd;
    int64_t first = 0;
    for (int name = 0; name < 2 * nd; ++name) {
        int64_t cnt = d->n_boundaries_per_direction[name];
        int bc = d->boundary_conditions[name], ic = d->boundary_ic[name];
#pragma omp parallel for schedule(static)
        for (int64_t B = first; B < first + cnt; ++B) {
            int64_t e = d->boundary_neighbor_ids[B] - 1;
            const int64_t *idx = d->boundary_node_indices + (int64_t)nd * B;
            int dir = p4_direction(nd, idx);
            for (int j = 0; j < nb; ++j)
                for (int i = 0; i < n; ++i) {
                    int64_t vn = p4_volume_node(nd, n, idx, i, j);
                    double ui[MAXV], f[MAXV], nrm[3] = {0, 0, 0};
                    for (int v = 0; v < nv; ++v) ui[v] = u[e * esz + nv * vn + v];
                    p4_normal(d, dir, vn, e, nrm);
                    const double *x = d->node_coordinates + (int64_t)nd * (vn + nn * e);
                    if (bc == TRIXI_B200_BC_DIRICHLET) {
                        double ub[MAXV];
                        ic_eval(&eq, ic, x, t, ub);
                        numflux_normal(&eq, d->surface_flux, ui, ub, nrm, f);
                        if (flux_has_noncons(d->surface_flux)) { /* equations.jl:232-247, dgsem_p4est/dg_3d.jl:550-590 */
                            double g[MAXV];
                            mhd_noncons_powell_normal(&eq, ui, ub, nrm, g);
                            for (int v = 0; v < nv; ++v) f[v] = f[v] + 0.5 * g[v];
                        }
                    } else if (bc == TRIXI_B200_BC_SLIP_WALL) { /* compressible_euler_3d.jl:315-366 */
                        euler_slip_wall_normal(&eq, ui, nrm, f);
                    } else {
                        for (int v = 0; v < nv; ++v) f[v] = NAN;
                    }
                    for (int v = 0; v < nv; ++v) sfv[e * fsz + v + nv * ((i + n * j) + nf * dir)] = f[v];
                }
        }
        first += cnt;
    }
}

/* calc_surface_integral! dgsem_p4est/dg_3d.jl:976-1034: outward normals everywhere => "+" on all six faces */
void oracle_calc_surface_integral_p4est(const trixi_b200_desc *d, double *du, const double *sfv) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd), fsz = (int64_t)nv * nf * 2 * nd;
    double factor = d->inverse_weights[0];
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < d->nelements; ++e) {
        double *due = du + e * esz;
        const double *s = sfv + e * fsz;
        for (int m = 0; m < nb; ++m)
            for (int l = 0; l < n; ++l) {
                int fn = l + n * m;
                for (int v = 0; v < nv; ++v)
                    for (int o = 0; o < nd; ++o) {
                        int lo = face_to_volume_node(nd, n, o, 0, l, m);
                        int hi = face_to_volume_node(nd, n, o, n - 1, l, m);
                        due[nv * lo + v] = due[nv * lo + v] + s[v + nv * (fn + nf * (2 * o))] * factor;
                        due[nv * hi + v] = due[nv * hi + v] + s[v + nv * (fn + nf * (2 * o + 1))] * factor;
                    }
            }
    }
}

/* rhs_hyperbolic! dgsem_tree/dg_2d.jl:113-186 dispatched for P4estMesh */
void oracle_calc_mortar_flux_p4est(const trixi_b200_desc *d, double *sfv, const double *u);
void oracle_calc_mpi_mortar_flux(const trixi_b200_desc *d, double *sfv, const double *u, const double *mpi_u);
void oracle_rhs_p4est(const trixi_b200_desc *d, double *du, const double *u, double t, double *interfaces_u,
                      double *sfv) {
    oracle_set_zero(d, du);
    oracle_calc_volume_integral_curved(d, du, u); /* shared kernels dgsem_structured/dg_3d.jl:36-175 */
    oracle_prolong2interfaces_p4est(d, interfaces_u, u);
    oracle_calc_interface_flux_p4est(d, sfv, interfaces_u);
    if (d->nboundaries > 0) oracle_calc_boundary_flux_p4est(d, sfv, u, t);
    oracle_calc_mortar_flux_p4est(d, sfv, u); /* prolong2mortars! + calc_mortar_flux! (no-op without mortars) */
    oracle_calc_surface_integral_p4est(d, du, sfv);
    oracle_apply_jacobian_curved(d, du);
    oracle_calc_sources(d, du, u, t);
}

/* ---- L2 mortars on TreeMesh ----------------------------------------------------------------------------
 * prolong2mortars! (dg_2d.jl:899-996, dg_3d.jl:770-957), calc_mortar_flux! (dg_2d.jl:1000-1036,
 * dg_3d.jl:959-1010) and mortar_fluxes_to_elements! (dg_2d.jl:1169-1243, dg_3d.jl:1236-1335) for
 * conservative equations.  Positions p = 0..2^(d-1)-1: bit 0 = upper half along the first face coordinate
 * ("right" in 3D, "upper" in 2D), bit 1 = upper half along the second ("upper" in 3D). */
static void mortar_apply_1d(const double *A, int n, int nv, int nb, int dim, const double *in, double *out, int add) {
    /* multiply_dimensionwise! (interpolation.jl): out[v, i, j] (+)= sum_ii A[i, ii] in[v, ii, j] along dim 0,
     * or out[v, i, j] (+)= sum_jj A[j, jj] in[v, i, jj] along dim 1; face data [nv, n, nb] */
    for (int j = 0; j < nb; ++j)
        for (int i = 0; i < n; ++i)
            for (int v = 0; v < nv; ++v) {
                double acc = 0.0;
                for (int q = 0; q < n; ++q) {
                    double a = dim == 0 ? A[i + n * q] : A[j + n * q];
                    double x = dim == 0 ? in[v + nv * (q + n * j)] : in[v + nv * (i + n * q)];
                    acc += a * x;
                }
                if (add)
                    out[v + nv * (i + n * j)] += acc;
                else
                    out[v + nv * (i + n * j)] = acc;
            }
}

void oracle_calc_mortar_flux(const trixi_b200_desc *d, double *sfv, const double *u) {
    if (d->nmortars <= 0) return;
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1, np = 1 << (nd - 1);
    int64_t esz = (int64_t)nv * ipow(n, nd), fsz = (int64_t)nv * nf * 2 * nd;
    const double *fwd[2] = {d->mortar_forward_lower, d->mortar_forward_upper};
    const double *rev[2] = {d->mortar_reverse_lower, d->mortar_reverse_upper};
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < d->nmortars; ++m) {
        const int64_t *ids = d->mortar_neighbor_ids + (int64_t)(np + 1) * m;
        int64_t large = ids[np] - 1;
        int o = (int)d->mortar_orientations[m] - 1;
        int large_side = (int)d->mortar_large_sides[m]; /* 1: large element on the negative (left) side */
        double ularge[MAXV * 64], tmp[MAXV * 64], uproj[MAXV * 64], fstar[4][MAXV * 64], usmall[MAXV * 64];
        double fprim[4][MAXV * 64];
        /* face of the large element that touches the mortar: its +face if it sits on the left */
        int lidx = large_side == 1 ? n - 1 : 0, sidx = large_side == 1 ? 0 : n - 1;
        for (int b = 0; b < nb; ++b)
            for (int a = 0; a < n; ++a) {
                int vn = face_to_volume_node(nd, n, o, lidx, a, b);
                for (int v = 0; v < nv; ++v) ularge[v + nv * (a + n * b)] = u[large * esz + nv * vn + v];
            }
        for (int p = 0; p < np; ++p) {
            int64_t small = ids[p] - 1;
            /* element_solutions_to_mortars!: forward interpolation of the large face to sub-face p */
            if (nd == 2) {
                mortar_apply_1d(fwd[p & 1], n, nv, 1, 0, ularge, uproj, 0);
            } else {
                mortar_apply_1d(fwd[p & 1], n, nv, nb, 0, ularge, tmp, 0);
                mortar_apply_1d(fwd[(p >> 1) & 1], n, nv, nb, 1, tmp, uproj, 0);
            }
            for (int b = 0; b < nb; ++b)
                for (int a = 0; a < n; ++a) {
                    int vn = face_to_volume_node(nd, n, o, sidx, a, b);
                    for (int v = 0; v < nv; ++v) usmall[v + nv * (a + n * b)] = u[small * esz + nv * vn + v];
                }
            /* calc_fstar!: left state = the side on the negative side of the mortar */
            for (int fn = 0; fn < nf; ++fn) {
                const double *ul = large_side == 1 ? uproj + nv * fn : usmall + nv * fn;
                const double *ur = large_side == 1 ? usmall + nv * fn : uproj + nv * fn;
                numflux(&eq, d->surface_flux, ul, ur, o, fstar[p] + nv * fn);
                for (int v = 0; v < nv; ++v) fprim[p][v + nv * fn] = fstar[p][v + nv * fn];
                if (flux_has_noncons(d->surface_flux)) {
                    /* dg_3d.jl:1012-1233: fstar_primary (kept by the small elements) takes
                     * nonconservative_flux(u_large, u_small), fstar_secondary (projected to the large element)
                     * nonconservative_flux(u_small, u_large), each weighted 0.5 -- for either large side */
                    const double *ularge_p = uproj + nv * fn, *usmall_p = usmall + nv * fn;
                    double np_[MAXV], ns_[MAXV];
                    mhd_noncons_powell(&eq, ularge_p, usmall_p, o, np_);
                    mhd_noncons_powell(&eq, usmall_p, ularge_p, o, ns_);
                    for (int v = 0; v < nv; ++v) {
                        fprim[p][v + nv * fn] += 0.5 * np_[v];
                        fstar[p][v + nv * fn] += 0.5 * ns_[v];
                    }
                }
            }
            /* small elements take the (primary) flux as it is: direction facing the large element */
            int dir_small = large_side == 1 ? 2 * o : 2 * o + 1;
            for (int q = 0; q < nv * nf; ++q) sfv[small * fsz + q + (int64_t)nv * nf * dir_small] = fprim[p][q];
        }
        /* L2 projection of the small fluxes onto the large face */
        int dir_large = large_side == 1 ? 2 * o + 1 : 2 * o;
        double *out = sfv + large * fsz + (int64_t)nv * nf * dir_large;
        if (nd == 2) {
            /* multiply_dimensionwise!(out, reverse_upper, f_upper, reverse_lower, f_lower) (dg_2d.jl:1238-1240) */
            for (int i = 0; i < n; ++i)
                for (int v = 0; v < nv; ++v) {
                    double acc = 0.0;
                    for (int q = 0; q < n; ++q)
                        acc += rev[1][i + n * q] * fstar[1][v + nv * q] + rev[0][i + n * q] * fstar[0][v + nv * q];
                    out[v + nv * i] = acc;
                }
        } else {
            /* upper_left, upper_right, lower_left, lower_right in this order (dg_3d.jl:1314-1331) */
            static const int order[4] = {2, 3, 0, 1};
            for (int k = 0; k < 4; ++k) {
                int p = order[k];
                mortar_apply_1d(rev[p & 1], n, nv, nb, 0, fstar[p], tmp, 0);
                mortar_apply_1d(rev[(p >> 1) & 1], n, nv, nb, 1, tmp, out, k > 0);
            }
        }
    }
}

/* ---- L2 mortars on P4estMesh -----------------------------------------------------------------------------
 * prolong2mortars! (dgsem_p4est/dg_2d.jl:802-868, dg_3d.jl:651-748), calc_mortar_flux! (dg_2d.jl:870-961,
 * dg_3d.jl:750-858, conservative equations) and mortar_fluxes_to_elements! (dg_2d.jl:997-1055, dg_3d.jl:890-974).
 * neighbor_ids [2^(d-1)+1, M]: small elements by position, then the large element; node_indices [nd, 2, M]:
 * 1 = small side (always forward), 2 = large side.  The flux uses the outward normal of the small element and
 * goes to the large element projected, with the sign switched and scaled by 2^(d-1). */
void oracle_calc_mortar_flux_p4est(const trixi_b200_desc *d, double *sfv, const double *u) {
    if (d->nmortars <= 0) return;
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1, np = 1 << (nd - 1);
    int64_t esz = (int64_t)nv * ipow(n, nd), fsz = (int64_t)nv * nf * 2 * nd;
    const double *fwd[2] = {d->mortar_forward_lower, d->mortar_forward_upper};
    const double *rev[2] = {d->mortar_reverse_lower, d->mortar_reverse_upper};
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < d->nmortars; ++m) {
        const int64_t *ids = d->mortar_neighbor_ids + (int64_t)(np + 1) * m;
        const int64_t *small_idx = d->mortar_node_indices + (int64_t)nd * (0 + 2 * m);
        const int64_t *large_idx = d->mortar_node_indices + (int64_t)nd * (1 + 2 * m);
        int64_t large = ids[np] - 1;
        int small_dir = p4_direction(nd, small_idx), large_dir = p4_direction(nd, large_idx);
        double u_buffer[MAXV * 64], tmp[MAXV * 64], u_large[4][MAXV * 64], fstar[4][MAXV * 64], fprim[4][MAXV * 64];
        /* prolong2mortars!: the large face in the orientation of the small side ... */
        for (int j = 0; j < nb; ++j)
            for (int i = 0; i < n; ++i) {
                int64_t vn = p4_volume_node(nd, n, large_idx, i, j);
                for (int v = 0; v < nv; ++v) u_buffer[v + nv * (i + n * j)] = u[large * esz + nv * vn + v];
            }
        /* ... interpolated to the 2^(d-1) small faces */
        for (int p = 0; p < np; ++p) {
            if (nd == 2) {
                mortar_apply_1d(fwd[p & 1], n, nv, 1, 0, u_buffer, u_large[p], 0);
            } else {
                mortar_apply_1d(fwd[p & 1], n, nv, nb, 0, u_buffer, tmp, 0);
                mortar_apply_1d(fwd[(p >> 1) & 1], n, nv, nb, 1, tmp, u_large[p], 0);
            }
        }
        /* calc_mortar_flux!: u_ll = small element, u_rr = interpolated large element, normal of the small element */
        for (int p = 0; p < np; ++p) {
            int64_t small = ids[p] - 1;
            for (int j = 0; j < nb; ++j)
                for (int i = 0; i < n; ++i) {
                    int64_t vn = p4_volume_node(nd, n, small_idx, i, j);
                    double nrm[3] = {0, 0, 0};
                    p4_normal(d, small_dir, vn, small, nrm);
                    const double *us = u + small * esz + nv * vn, *ul_ = u_large[p] + nv * (i + n * j);
                    double *fsec = fstar[p] + nv * (i + n * j), *fpri = fprim[p] + nv * (i + n * j);
                    numflux_normal(&eq, d->surface_flux, us, ul_, nrm, fsec);
                    for (int v = 0; v < nv; ++v) fpri[v] = fsec[v];
                    if (flux_has_noncons(d->surface_flux)) { /* dg_3d.jl:860-888: 0.5 g(u_ll, u_rr) / 0.5 g(u_rr, u_ll) */
                        double gp[MAXV], gs[MAXV];
                        mhd_noncons_powell_normal(&eq, us, ul_, nrm, gp);
                        mhd_noncons_powell_normal(&eq, ul_, us, nrm, gs);
                        for (int v = 0; v < nv; ++v) {
                            fpri[v] = fsec[v] + 0.5 * gp[v];
                            fsec[v] = fsec[v] + 0.5 * gs[v];
                        }
                    }
                }
            /* mortar_fluxes_to_elements!: small to small */
            for (int q = 0; q < nv * nf; ++q) sfv[small * fsz + q + (int64_t)nv * nf * small_dir] = fprim[p][q];
        }
        /* project the small fluxes to the large element */
        if (nd == 2) {
            /* multiply_dimensionwise!(u_buffer, reverse_upper, fstar[2], reverse_lower, fstar[1]) (dg_2d.jl:1020-1022) */
            for (int i = 0; i < n; ++i)
                for (int v = 0; v < nv; ++v) {
                    double acc = 0.0;
                    for (int q = 0; q < n; ++q)
                        acc += rev[1][i + n * q] * fstar[1][v + nv * q] + rev[0][i + n * q] * fstar[0][v + nv * q];
                    u_buffer[v + nv * i] = acc;
                }
        } else {
            /* positions 1..4 in this order (dg_3d.jl:914-929) */
            for (int p = 0; p < 4; ++p) {
                mortar_apply_1d(rev[p & 1], n, nv, nb, 0, fstar[p], tmp, 0);
                mortar_apply_1d(rev[(p >> 1) & 1], n, nv, nb, 1, tmp, u_buffer, p > 0);
            }
        }
        /* sign switch and scaling by the area ratio (dg_2d.jl:1024-1031, dg_3d.jl:931-939), then the copy in the
         * orientation of the large face (surface_indices of the large side) */
        double scale = nd == 2 ? -2.0 : -4.0;
        for (int j = 0; j < nb; ++j)
            for (int i = 0; i < n; ++i) {
                int fn_large;
                p4_surface_node(nd, n, large_idx, i, j, &fn_large);
                for (int v = 0; v < nv; ++v)
                    sfv[large * fsz + v + nv * (fn_large + nf * large_dir)] = u_buffer[v + nv * (i + n * j)] * scale;
            }
    }
}

/* ---- MPI mortars ---------------------------------------------------------------------------------------------
 * calc_mpi_mortar_flux! + mpi_mortar_fluxes_to_elements! (dgsem_tree/dg_2d_parallel.jl:742-860, dgsem_p4est/
 * dg_2d_parallel.jl:270-420, dg_3d_parallel.jl:382-560): the arithmetic of the serial mortars above, with the faces
 * of remote elements taken from the exchanged buffer (the remote side of the MPI-interface entry a negative
 * neighbor id points to) and fluxes stored for local elements only.  The reference ships all faces of a mortar in
 * one buffer; here every (large, small) pair of different ranks is one exchange-only MPI-interface entry. */
static void mm_fetch_face(const trixi_b200_desc *d, int64_t id, const double *u, const double *mpi_u, int o, int fidx,
                          const int64_t *idx, double *out) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd);
    if (id > 0) {
        for (int b = 0; b < nb; ++b)
            for (int a = 0; a < n; ++a) {
                int64_t vn = idx ? p4_volume_node(nd, n, idx, a, b) : face_to_volume_node(nd, n, o, fidx, a, b);
                for (int v = 0; v < nv; ++v) out[v + nv * (a + n * b)] = u[(id - 1) * esz + nv * vn + v];
            }
    } else {
        int64_t I = -id - 1;
        int remote = 1 - ((int)d->mpi_local_sides[I] - 1);
        for (int fn = 0; fn < nf; ++fn)
            for (int v = 0; v < nv; ++v) out[v + nv * fn] = mpi_u[remote + 2 * (v + nv * (fn + (int64_t)nf * I))];
    }
}

void oracle_calc_mpi_mortar_flux(const trixi_b200_desc *d, double *sfv, const double *u, const double *mpi_u) {
    if (d->nmpimortars <= 0) return;
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1, np = 1 << (nd - 1);
    int64_t fsz = (int64_t)nv * nf * 2 * nd;
    int p4 = d->mesh_kind == TRIXI_B200_MESH_P4EST;
    const double *fwd[2] = {d->mortar_forward_lower, d->mortar_forward_upper};
    const double *rev[2] = {d->mortar_reverse_lower, d->mortar_reverse_upper};
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < d->nmpimortars; ++m) {
        const int64_t *ids = d->mpi_mortar_neighbor_ids + (int64_t)(np + 1) * m;
        const int64_t *small_idx = p4 ? d->mpi_mortar_node_indices + (int64_t)nd * (0 + 2 * m) : NULL;
        const int64_t *large_idx = p4 ? d->mpi_mortar_node_indices + (int64_t)nd * (1 + 2 * m) : NULL;
        int o = p4 ? 0 : (int)d->mpi_mortar_orientations[m] - 1;
        int large_side = p4 ? 0 : (int)d->mpi_mortar_large_sides[m];
        int lidx = large_side == 1 ? n - 1 : 0, sidx = large_side == 1 ? 0 : n - 1;
        int small_dir = p4 ? p4_direction(nd, small_idx) : (large_side == 1 ? 2 * o : 2 * o + 1);
        int large_dir = p4 ? p4_direction(nd, large_idx) : (large_side == 1 ? 2 * o + 1 : 2 * o);
        double ularge[MAXV * 64], tmp[MAXV * 64], uproj[MAXV * 64], usmall[MAXV * 64], fstar[4][MAXV * 64],
            fprim[MAXV * 64], out[MAXV * 64];
        if (ids[np] == 0) continue; /* cannot happen: a rank with a small element receives the large face */
        mm_fetch_face(d, ids[np], u, mpi_u, o, lidx, large_idx, ularge);
        for (int p = 0; p < np; ++p) {
            if (ids[p] == 0) continue; /* neither local nor needed */
            if (nd == 2) {
                mortar_apply_1d(fwd[p & 1], n, nv, 1, 0, ularge, uproj, 0);
            } else {
                mortar_apply_1d(fwd[p & 1], n, nv, nb, 0, ularge, tmp, 0);
                mortar_apply_1d(fwd[(p >> 1) & 1], n, nv, nb, 1, tmp, uproj, 0);
            }
            mm_fetch_face(d, ids[p], u, mpi_u, o, sidx, small_idx, usmall);
            for (int fn = 0; fn < nf; ++fn) {
                double *f = fstar[p] + nv * fn;
                if (p4) {
                    const double *nrm = d->mpi_mortar_normal_directions + (int64_t)nd * (fn + (int64_t)nf * (p + (int64_t)np * m));
                    double nn[3] = {0, 0, 0};
                    for (int k = 0; k < nd; ++k) nn[k] = nrm[k];
                    numflux_normal(&eq, d->surface_flux, usmall + nv * fn, uproj + nv * fn, nn, f);
                    if (flux_has_noncons(d->surface_flux)) {
                        double gp[MAXV], gs[MAXV];
                        mhd_noncons_powell_normal(&eq, usmall + nv * fn, uproj + nv * fn, nn, gp);
                        mhd_noncons_powell_normal(&eq, uproj + nv * fn, usmall + nv * fn, nn, gs);
                        for (int v = 0; v < nv; ++v) fprim[v + nv * fn] = f[v] + 0.5 * gp[v];
                        for (int v = 0; v < nv; ++v) f[v] = f[v] + 0.5 * gs[v];
                        continue;
                    }
                } else {
                    const double *ul = large_side == 1 ? uproj + nv * fn : usmall + nv * fn;
                    const double *ur = large_side == 1 ? usmall + nv * fn : uproj + nv * fn;
                    numflux(&eq, d->surface_flux, ul, ur, o, f);
                }
                for (int v = 0; v < nv; ++v) fprim[v + nv * fn] = f[v];
                if (!p4 && flux_has_noncons(d->surface_flux)) {
                    double np_[MAXV], ns_[MAXV];
                    mhd_noncons_powell(&eq, uproj + nv * fn, usmall + nv * fn, o, np_);
                    mhd_noncons_powell(&eq, usmall + nv * fn, uproj + nv * fn, o, ns_);
                    for (int v = 0; v < nv; ++v) {
                        fprim[v + nv * fn] += 0.5 * np_[v];
                        f[v] += 0.5 * ns_[v];
                    }
                }
            }
            if (ids[p] > 0)
                for (int q = 0; q < nv * nf; ++q) sfv[(ids[p] - 1) * fsz + q + (int64_t)nv * nf * small_dir] = fprim[q];
        }
        if (ids[np] < 0) continue; /* the large element belongs to another rank */
        int64_t large = ids[np] - 1;
        if (p4) {
            if (nd == 2) {
                for (int i = 0; i < n; ++i)
                    for (int v = 0; v < nv; ++v) {
                        double acc = 0.0;
                        for (int q = 0; q < n; ++q)
                            acc += rev[1][i + n * q] * fstar[1][v + nv * q] + rev[0][i + n * q] * fstar[0][v + nv * q];
                        out[v + nv * i] = acc;
                    }
            } else {
                for (int p = 0; p < 4; ++p) {
                    mortar_apply_1d(rev[p & 1], n, nv, nb, 0, fstar[p], tmp, 0);
                    mortar_apply_1d(rev[(p >> 1) & 1], n, nv, nb, 1, tmp, out, p > 0);
                }
            }
            double scale = nd == 2 ? -2.0 : -4.0;
            for (int j = 0; j < nb; ++j)
                for (int i = 0; i < n; ++i) {
                    int fn_large;
                    p4_surface_node(nd, n, large_idx, i, j, &fn_large);
                    for (int v = 0; v < nv; ++v)
                        sfv[large * fsz + v + nv * (fn_large + nf * large_dir)] = out[v + nv * (i + n * j)] * scale;
                }
        } else {
            double *dst = sfv + large * fsz + (int64_t)nv * nf * large_dir;
            if (nd == 2) {
                for (int i = 0; i < n; ++i)
                    for (int v = 0; v < nv; ++v) {
                        double acc = 0.0;
                        for (int q = 0; q < n; ++q)
                            acc += rev[1][i + n * q] * fstar[1][v + nv * q] + rev[0][i + n * q] * fstar[0][v + nv * q];
                        dst[v + nv * i] = acc;
                    }
            } else {
                static const int order[4] = {2, 3, 0, 1};
                for (int k = 0; k < 4; ++k) {
                    int p = order[k];
                    mortar_apply_1d(rev[p & 1], n, nv, nb, 0, fstar[p], tmp, 0);
                    mortar_apply_1d(rev[(p >> 1) & 1], n, nv, nb, 1, tmp, dst, k > 0);
                }
            }
        }
    }
}

/* rhs_hyperbolic! dgsem_tree/dg_2d.jl:113-186.  Work arrays: interfaces_u [2,nv,nf,I],
 * boundaries_u [2,nv,nf,B], sfv [nv,nf,2nd,nelem] (owned by the caller = the cache). */
void oracle_rhs(const trixi_b200_desc *d, double *du, const double *u, double t, double *interfaces_u,
                double *boundaries_u, double *sfv) {
    if (d->mesh_kind == TRIXI_B200_MESH_STRUCTURED) {
        oracle_rhs_structured(d, du, u, t, interfaces_u, sfv);
        return;
    }
    if (d->mesh_kind == TRIXI_B200_MESH_P4EST) {
        oracle_rhs_p4est(d, du, u, t, interfaces_u, sfv);
        return;
    }
    oracle_set_zero(d, du);
    oracle_calc_volume_integral(d, du, u);
    oracle_prolong2interfaces(d, interfaces_u, u);
    oracle_calc_interface_flux(d, sfv, interfaces_u);
    if (d->nboundaries > 0) {
        oracle_prolong2boundaries(d, boundaries_u, u);
        oracle_calc_boundary_flux(d, sfv, boundaries_u, t);
    }
    oracle_calc_mortar_flux(d, sfv, u); /* prolong2mortars! + calc_mortar_flux! (no-op without mortars) */
    oracle_calc_surface_integral(d, du, sfv);
    oracle_apply_jacobian(d, du);
    oracle_calc_sources(d, du, u, t);
}

/* distributed rhs_hyperbolic! (dgsem_tree/dg_2d_parallel.jl:453-563, p4est dg_3d_parallel.jl:8-117) in
 * two halves around the halo exchange the caller performs: part 1 = prolong2mpiinterfaces + all local
 * work up to the boundary fluxes; part 2 = calc_mpi_interface_flux!, surface integral, Jacobian, sources */
/* prolong2mpiinterfaces! dgsem_p4est/dg_3d_parallel.jl:119-165: mpi_u[2, nv, nf, MI], the local side filled
 * through the local element's node_indices (interface aligned at the primary element) */
void oracle_prolong2mpiinterfaces_p4est(const trixi_b200_desc *d, double *mu, const double *u) {
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t esz = (int64_t)nv * ipow(n, nd);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->nmpiinterfaces; ++I) {
        int64_t e = d->mpi_local_neighbor_ids[I] - 1;
        int side = (int)d->mpi_local_sides[I] - 1;
        const int64_t *idx = d->mpi_node_indices + (int64_t)nd * I;
        for (int j = 0; j < nb; ++j)
            for (int i = 0; i < n; ++i) {
                int64_t vn = p4_volume_node(nd, n, idx, i, j);
                for (int v = 0; v < nv; ++v)
                    mu[side + 2 * (v + nv * ((i + n * j) + (int64_t)nf * I))] = u[e * esz + nv * vn + v];
            }
    }
}

/* calc_mpi_interface_flux! dgsem_p4est/dg_3d_parallel.jl:167-273: normal of the LOCAL element; the secondary
 * side stores -surface_flux(u_ll, u_rr, -normal) (:262-266) at its own surface node */
void oracle_calc_mpi_interface_flux_p4est(const trixi_b200_desc *d, double *sfv, const double *mu) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int nf = ipow(n, nd - 1), nb = nd == 3 ? n : 1;
    int64_t fsz = (int64_t)nv * nf * 2 * nd;
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < d->nmpiinterfaces; ++I) {
        if (d->mpi_is_mortar_piece && d->mpi_is_mortar_piece[I]) continue; /* exchange only: see MPI mortars */
        int64_t e = d->mpi_local_neighbor_ids[I] - 1;
        int side = (int)d->mpi_local_sides[I];
        const int64_t *idx = d->mpi_node_indices + (int64_t)nd * I;
        int dir = p4_direction(nd, idx);
        for (int j = 0; j < nb; ++j)
            for (int i = 0; i < n; ++i) {
                double ul[MAXV], ur[MAXV], f[MAXV], nrm[3] = {0, 0, 0};
                int fn = i + n * j, fn_loc;
                for (int v = 0; v < nv; ++v) {
                    ul[v] = mu[0 + 2 * (v + nv * (fn + (int64_t)nf * I))];
                    ur[v] = mu[1 + 2 * (v + nv * (fn + (int64_t)nf * I))];
                }
                p4_normal(d, dir, p4_volume_node(nd, n, idx, i, j), e, nrm);
                if (side == 2)
                    for (int dim = 0; dim < nd; ++dim) nrm[dim] = -nrm[dim];
                numflux_normal(&eq, d->surface_flux, ul, ur, nrm, f);
                p4_surface_node(nd, n, idx, i, j, &fn_loc);
                if (flux_has_noncons(d->surface_flux)) { /* dg_3d_parallel.jl:302-337 */
                    double g[MAXV];
                    if (side == 1)
                        mhd_noncons_powell_normal(&eq, ul, ur, nrm, g);
                    else
                        mhd_noncons_powell_normal(&eq, ur, ul, nrm, g);
                    for (int v = 0; v < nv; ++v)
                        sfv[e * fsz + v + nv * (fn_loc + nf * dir)] = side == 1 ? f[v] + 0.5 * g[v] : -f[v] + 0.5 * (-g[v]);
                    continue;
                }
                for (int v = 0; v < nv; ++v)
                    sfv[e * fsz + v + nv * (fn_loc + nf * dir)] = side == 1 ? f[v] : -f[v];
            }
    }
}

void oracle_rhs_parallel_part1(const trixi_b200_desc *d, double *du, const double *u, double t,
                               double *interfaces_u, double *boundaries_u, double *sfv, double *mpi_u) {
    if (d->mesh_kind == TRIXI_B200_MESH_P4EST) { /* dgsem_p4est/dg_3d_parallel.jl:8-117 */
        oracle_prolong2mpiinterfaces_p4est(d, mpi_u, u);
        oracle_set_zero(d, du);
        oracle_calc_volume_integral_curved(d, du, u);
        oracle_prolong2interfaces_p4est(d, interfaces_u, u);
        oracle_calc_interface_flux_p4est(d, sfv, interfaces_u);
        if (d->nboundaries > 0) oracle_calc_boundary_flux_p4est(d, sfv, u, t);
        oracle_calc_mortar_flux_p4est(d, sfv, u);
        return;
    }
    oracle_prolong2mpiinterfaces(d, mpi_u, u);
    oracle_set_zero(d, du);
    oracle_calc_volume_integral(d, du, u);
    oracle_prolong2interfaces(d, interfaces_u, u);
    oracle_calc_interface_flux(d, sfv, interfaces_u);
    if (d->nboundaries > 0) {
        oracle_prolong2boundaries(d, boundaries_u, u);
        oracle_calc_boundary_flux(d, sfv, boundaries_u, t);
    }
    oracle_calc_mortar_flux(d, sfv, u);
}
void oracle_rhs_parallel_part2(const trixi_b200_desc *d, double *du, const double *u, double t, double *sfv,
                               const double *mpi_u) {
    if (d->mesh_kind == TRIXI_B200_MESH_P4EST) {
        oracle_calc_mpi_interface_flux_p4est(d, sfv, mpi_u);
        oracle_calc_mpi_mortar_flux(d, sfv, u, mpi_u);
        oracle_calc_surface_integral_p4est(d, du, sfv);
        oracle_apply_jacobian_curved(d, du);
        oracle_calc_sources(d, du, u, t);
        return;
    }
    oracle_calc_mpi_interface_flux(d, sfv, mpi_u);
    oracle_calc_mpi_mortar_flux(d, sfv, u, mpi_u);
    oracle_calc_surface_integral(d, du, sfv);
    oracle_apply_jacobian(d, du);
    oracle_calc_sources(d, du, u, t);
}

/* integrate_via_indices (callbacks_step/analysis_dg3d.jl:364-425, analysis_dg2d.jl analogues) with the integrands of
 * analysis_integrals: integrate(u) (cons2cons), entropy (compressible_euler_3d.jl:1959-2009), energy_total/kinetic/
 * internal (:2012-2023), entropy_timederivative = cons2entropy(u) . du (:1796-1817, analysis_dg3d.jl:506-517).
 * out: [nvars] for TRIXI_B200_INTEGRAL_CONS, [1] otherwise; NOT normalised; *volume = quadrature of the volume. */
void oracle_integrate(const trixi_b200_desc *d, int quantity, const double *u, const double *du, double *out,
                      double *volume) {
    eqn_t eq = make_eqn(d);
    int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    int n3 = nd == 3 ? n : 1;
    int64_t nn = ipow(n, nd);
    int curved = d->mesh_kind != TRIXI_B200_MESH_TREE;
    double sums[MAXV] = {0}, vol = 0.0;
    for (int64_t e = 0; e < d->nelements; ++e)
        for (int k = 0; k < n3; ++k)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    int64_t node = i + n * (j + n * k);
                    double w = (1 / d->inverse_weights[i]) * (1 / d->inverse_weights[j]) * (nd == 3 ? 1 / d->inverse_weights[k] : 1.0);
                    if (curved)
                        w *= fabs(1 / d->inverse_jacobian[node + nn * e]);
                    else
                        for (int a = 0; a < nd; ++a) w *= 1 / d->inverse_jacobian[e];
                    const double *un = u + nv * (node + nn * e);
                    vol += w;
                    if (quantity == TRIXI_B200_INTEGRAL_CONS) {
                        for (int v = 0; v < nv; ++v) sums[v] += w * un[v];
                        continue;
                    }
                    double val = NAN;
                    if (is_euler(&eq)) {
                        double rho = un[0], msq = 0.0;
                        for (int a = 0; a < nd; ++a) msq += un[1 + a] * un[1 + a];
                        if (quantity == TRIXI_B200_INTEGRAL_ENERGY_TOTAL)
                            val = un[nd + 1];
                        else if (quantity == TRIXI_B200_INTEGRAL_ENERGY_KINETIC)
                            val = 0.5 * msq / rho;
                        else if (quantity == TRIXI_B200_INTEGRAL_ENERGY_INTERNAL)
                            val = un[nd + 1] - 0.5 * msq / rho;
                        else if (quantity == TRIXI_B200_INTEGRAL_ENTROPY) {
                            double p = (eq.gamma - 1) * (un[nd + 1] - 0.5 * msq / rho);
                            double s_ = log(p) - eq.gamma * log(rho);
                            val = -s_ * rho * eq.inv_gm1;
                        } else if (quantity == TRIXI_B200_INTEGRAL_ENTROPY_TIMEDERIVATIVE) {
                            const double *dn = du + nv * (node + nn * e);
                            double v[3] = {0, 0, 0}, v_square = 0.0;
                            for (int a = 0; a < nd; ++a) {
                                v[a] = un[1 + a] / rho;
                                v_square += v[a] * v[a];
                            }
                            double p = (eq.gamma - 1) * (un[nd + 1] - 0.5 * rho * v_square);
                            double s_ = log(p) - eq.gamma * log(rho);
                            double rho_p = rho / p;
                            val = ((eq.gamma - s_) * eq.inv_gm1 - 0.5 * rho_p * v_square) * dn[0];
                            for (int a = 0; a < nd; ++a) val += rho_p * v[a] * dn[1 + a];
                            val += -rho_p * dn[nd + 1];
                        }
                    }
                    sums[0] += w * val;
                }
    int nvals = quantity == TRIXI_B200_INTEGRAL_CONS ? nv : 1;
    for (int v = 0; v < nvals; ++v) out[v] = sums[v];
    *volume = vol;
}

/* max_dt stepsize_dg3d.jl:8-32 (constant_speed False), stepsize_dg2d.jl:60-75 (True) */
double oracle_max_dt(const trixi_b200_desc *d, const double *u) {
    if (d->mesh_kind != TRIXI_B200_MESH_TREE) return oracle_max_dt_curved(d, u);
    eqn_t eq = make_eqn(d);
    int nv = d->nvars, nd = d->ndims;
    int64_t nn = ipow(d->nnodes, nd);
    double max_scaled_speed = DBL_TRUE_MIN; /* nextfloat(zero(t)) */
    int nanflag = 0;
#pragma omp parallel for schedule(static) reduction(max : max_scaled_speed) reduction(| : nanflag)
    for (int64_t e = 0; e < d->nelements; ++e) {
        double ml[3] = {0, 0, 0};
        if (is_euler(&eq)) {
            for (int64_t q = 0; q < nn; ++q) {
                double rho, v[3], p;
                euler_cons2prim(&eq, u + nv * (q + nn * e), &rho, v, &p);
                double c = sqrt(eq.gamma * p / rho); /* max_abs_speeds compressible_euler_3d.jl:1770-1775 */
                for (int dd = 0; dd < nd; ++dd) {
                    double l = fabs(v[dd]) + c;
                    if (isnan(l)) nanflag = 1;
                    ml[dd] = fmax(ml[dd], l);
                }
            }
        } else if (is_mhd(&eq)) { /* max_abs_speeds ideal_glm_mhd_3d.jl:1218-1228 */
            for (int64_t q = 0; q < nn; ++q) {
                const double *un = u + nv * (q + nn * e);
                for (int dd = 0; dd < 3; ++dd) {
                    double l = fabs(un[1 + dd] / un[0]) + mhd_fast_wavespeed(&eq, un, dd);
                    if (isnan(l)) nanflag = 1;
                    ml[dd] = fmax(ml[dd], l);
                }
            }
        } else {
            for (int dd = 0; dd < nd; ++dd) ml[dd] = fabs(eq.a[dd]);
        }
        double s = 0.0;
        for (int dd = 0; dd < nd; ++dd) s += ml[dd];
        double val = d->inverse_jacobian[e] * s;
        if (val > max_scaled_speed) max_scaled_speed = val;
    }
    if (nanflag) return NAN; /* Base.max propagates NaN */
    return 2 / (d->nnodes * max_scaled_speed);
}

/* stage loop of step!(::SimpleIntegrator2N) methods_2N.jl:144-159 */
void oracle_step_2n(const trixi_b200_desc *d, double *u, double *du, double *u_tmp, double t, double dt,
                    const double *a, const double *b, const double *c, int nstages, double *interfaces_u,
                    double *boundaries_u, double *sfv) {
    int64_t len = (int64_t)d->nvars * ipow(d->nnodes, d->ndims) * d->nelements;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < len; ++i) u_tmp[i] = 0.0;
    for (int s = 0; s < nstages; ++s) {
        double t_stage = t + dt * c[s];
        oracle_rhs(d, du, u, t_stage, interfaces_u, boundaries_u, sfv);
        double a_stage = a[s], b_stage_dt = b[s] * dt;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < len; ++i) {
            u_tmp[i] = du[i] - u_tmp[i] * a_stage;
            u[i] += u_tmp[i] * b_stage_dt;
        }
    }
}

/* pointwise two-point flux for unit tests (flux consistency, test/test_unit.jl:1590-1641) */
void oracle_numflux(const trixi_b200_desc *d, int flux_id, const double *ul, const double *ur, int orientation,
                    double *f) {
    eqn_t eq = make_eqn(d);
    numflux(&eq, flux_id, ul, ur, orientation - 1, f);
}
void oracle_numflux_normal(const trixi_b200_desc *d, int flux_id, const double *ul, const double *ur,
                           const double *normal_direction, double *f) {
    eqn_t eq = make_eqn(d);
    numflux_normal(&eq, flux_id, ul, ur, normal_direction, f);
}
void oracle_flux_normal(const trixi_b200_desc *d, const double *u, const double *normal_direction, double *f) {
    eqn_t eq = make_eqn(d);
    phys_flux_normal(&eq, u, normal_direction, f);
}
double oracle_ln_mean(double x, double y) { return ln_mean(x, y); }
double oracle_inv_ln_mean(double x, double y) { return inv_ln_mean(x, y); }
void oracle_flux(const trixi_b200_desc *d, const double *u, int orientation, double *f) {
    eqn_t eq = make_eqn(d);
    phys_flux(&eq, u, orientation - 1, f);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
