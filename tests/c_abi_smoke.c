/* C-ABI smoke test: a plain C host (no Python, no ctypes, no torch) drives libtrixi_b200.so through the calls the
 * Julia binding makes -- create -> upload -> rhs_host -> max_dt -> step_2n -> download -> destroy -- on a small
 * periodic 3D compressible Euler problem (flux-differencing DGSEM, flux_ranocha, polydeg 3, the headline
 * configuration) and checks every result against the CPU oracle called with the SAME descriptor.
 *
 *   gcc -O2 -std=c11 -Iinclude tests/c_abi_smoke.c -o c_abi_smoke \
 *       trixi.jl_b200/libtrixi_b200.so oracle/libtrixi_oracle.so -lm -Wl,-rpath,...
 *
 * TEST INFRASTRUCTURE: links the oracle as the checker (tests/test_c_abi_smoke.py builds and runs it, -m gpu).
 * The containers are built here the way create_cache does (dgsem_tree/containers_3d.jl: elements 9-18, interfaces
 * 136-144) for a uniform Cartesian box; the basis follows basis_lobatto_legendre.jl:17-31,570-664 for polydeg 3. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "trixi_b200.h"

/* oracle/trixi_oracle.c (the checker) */
void oracle_rhs(const trixi_b200_desc *d, double *du, const double *u, double t, double *interfaces_u,
                double *boundaries_u, double *sfv);
double oracle_max_dt(const trixi_b200_desc *d, const double *u);
void oracle_step_2n(const trixi_b200_desc *d, double *u, double *du, double *u_tmp, double t, double dt,
                    const double *a, const double *b, const double *c, int nstages, double *interfaces_u,
                    double *boundaries_u, double *sfv);

#define N 4       /* nnodes = polydeg + 1 */
#define NV 5      /* compressible Euler 3D */
#define CELLS 4   /* elements per direction */

static double rel_err(const double *a, const double *b, int64_t n) {
    double num = 0.0, den = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double d = fabs(a[i] - b[i]);
        if (!(d <= num)) num = d; /* NaN-propagating max */
        if (fabs(b[i]) > den) den = fabs(b[i]);
    }
    return num / den;
}

#define CHECK_RC(call)                                                                            \
    do {                                                                                          \
        int rc__ = (call);                                                                        \
        if (rc__ != TRIXI_B200_OK) {                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, trixi_b200_last_error(h));             \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

int main(void) {
    /* ---- LobattoLegendreBasis(3): nodes, weights, derivative matrices (column-major [n, n]) ---- */
    const double nodes[N] = {-1.0, -sqrt(1.0 / 5.0), sqrt(1.0 / 5.0), 1.0};
    const double weights[N] = {1.0 / 6.0, 5.0 / 6.0, 5.0 / 6.0, 1.0 / 6.0};
    double wbary[N], D[N][N], dsplit[N * N], dhat[N * N], inv_weights[N];
    for (int j = 0; j < N; ++j) {
        double w = 1.0;
        for (int k = 0; k < N; ++k)
            if (k != j) w *= nodes[j] - nodes[k];
        wbary[j] = 1.0 / w;
    }
    for (int i = 0; i < N; ++i) {
        D[i][i] = 0.0;
        for (int j = 0; j < N; ++j)
            if (j != i) {
                D[i][j] = (wbary[j] / wbary[i]) / (nodes[i] - nodes[j]);
                D[i][i] -= D[i][j];
            }
    }
    for (int i = 0; i < N; ++i) {
        inv_weights[i] = 1.0 / weights[i];
        for (int j = 0; j < N; ++j) {
            dsplit[i + N * j] = 2.0 * D[i][j];                        /* calc_dsplit */
            dhat[i + N * j] = -D[j][i] * weights[j] / weights[i];     /* calc_dhat */
        }
    }
    dsplit[0] += 1.0 / weights[0];
    dsplit[N * N - 1] -= 1.0 / weights[N - 1];

    /* ---- containers of a periodic CELLS^3 box on [-2, 2]^3 ---- */
    const int64_t nel = (int64_t)CELLS * CELLS * CELLS, nn = N * N * N, nif = 3 * nel;
    const double dx = 4.0 / CELLS;
    double *inverse_jacobian = malloc(sizeof(double) * nel);
    double *coords = malloc(sizeof(double) * 3 * nn * nel);
    int64_t *if_ids = malloc(sizeof(int64_t) * 2 * nif), *if_orient = malloc(sizeof(int64_t) * nif);
    for (int64_t e = 0; e < nel; ++e) {
        const int c[3] = {(int)(e % CELLS), (int)((e / CELLS) % CELLS), (int)(e / (CELLS * CELLS))};
        inverse_jacobian[e] = 2.0 / dx;
        for (int k = 0; k < N; ++k)
            for (int j = 0; j < N; ++j)
                for (int i = 0; i < N; ++i) {
                    const int idx[3] = {i, j, k};
                    for (int d = 0; d < 3; ++d)
                        coords[d + 3 * ((i + N * (j + N * k)) + nn * e)] = -2.0 + dx * (c[d] + 0.5) + 0.5 * dx * nodes[idx[d]];
                }
        for (int d = 0; d < 3; ++d) { /* the interface towards the +d neighbour; 1-based ids */
            int cn[3] = {c[0], c[1], c[2]};
            cn[d] = (cn[d] + 1) % CELLS;
            const int64_t I = 3 * e + d;
            if_ids[2 * I] = e + 1;
            if_ids[2 * I + 1] = cn[0] + CELLS * (cn[1] + CELLS * cn[2]) + 1;
            if_orient[I] = d + 1;
        }
    }

    trixi_b200_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.abi_version = TRIXI_B200_ABI_VERSION;
    desc.device = -1;
    desc.ndims = 3;
    desc.nvars = NV;
    desc.nnodes = N;
    desc.mesh_kind = TRIXI_B200_MESH_TREE;
    desc.nelements = nel;
    desc.equation = TRIXI_B200_EQ_EULER_3D;
    desc.volume_integral = TRIXI_B200_VOLINT_FLUX_DIFFERENCING;
    desc.volume_flux = TRIXI_B200_FLUX_RANOCHA;
    desc.surface_flux = TRIXI_B200_FLUX_RANOCHA;
    desc.source_terms = TRIXI_B200_SRC_NONE;
    desc.eq_params[0] = 1.4;
    desc.eq_params[1] = 1.0 / (1.4 - 1.0);
    desc.derivative_split = dsplit;
    desc.derivative_hat = dhat;
    desc.inverse_weights = inv_weights;
    desc.inverse_jacobian = inverse_jacobian;
    desc.node_coordinates = coords;
    desc.ninterfaces = nif;
    desc.interface_neighbor_ids = if_ids;
    desc.interface_orientations = if_orient;
    desc.rank = 0;
    desc.world_size = 1;

    /* ---- a smooth periodic state, and one with a 17% density/pressure jump (both ln_mean branches) ---- */
    const int64_t len = NV * nn * nel;
    double *u0 = malloc(sizeof(double) * len), *u = malloc(sizeof(double) * len), *du = malloc(sizeof(double) * len);
    double *u_ref = malloc(sizeof(double) * len), *du_ref = malloc(sizeof(double) * len), *ut_ref = malloc(sizeof(double) * len);
    double *iu = malloc(sizeof(double) * 2 * NV * N * N * nif), *sfv = malloc(sizeof(double) * NV * N * N * 6 * nel);
    for (int64_t q = 0; q < nn * nel; ++q) {
        const double x = coords[3 * q], y = coords[3 * q + 1], z = coords[3 * q + 2];
        const double r = sqrt(x * x + y * y + z * z);
        const double rho = (r < 0.9 ? 1.17 : 1.0) + 0.05 * sin(0.5 * M_PI * x) * cos(0.5 * M_PI * y);
        const double v1 = 0.1 + 0.05 * sin(0.5 * M_PI * z), v2 = -0.2, v3 = 0.15 * cos(0.5 * M_PI * x);
        const double p = (r < 0.9 ? 1.245 : 1.0) + 0.02 * cos(0.5 * M_PI * (y + z));
        u0[NV * q] = rho;
        u0[NV * q + 1] = rho * v1;
        u0[NV * q + 2] = rho * v2;
        u0[NV * q + 3] = rho * v3;
        u0[NV * q + 4] = p / 0.4 + 0.5 * rho * (v1 * v1 + v2 * v2 + v3 * v3);
    }

    trixi_b200_handle *h = NULL;
    if (trixi_b200_abi_version() != TRIXI_B200_ABI_VERSION) {
        fprintf(stderr, "ABI version mismatch\n");
        return 1;
    }
    int rc = trixi_b200_create(&desc, &h);
    if (rc != TRIXI_B200_OK) {
        fprintf(stderr, "trixi_b200_create -> %d: %s\n", rc, trixi_b200_last_error(NULL));
        return rc == TRIXI_B200_ENODEVICE ? 77 : 1; /* 77: skipped, no GPU (there is no CPU fallback) */
    }

    /* 1. rhs_hyperbolic! with host buffers */
    CHECK_RC(trixi_b200_rhs_host(h, du, u0, 0.25));
    oracle_rhs(&desc, du_ref, u0, 0.25, iu, NULL, sfv);
    const double err_rhs = rel_err(du, du_ref, len);

    /* 2. max_dt on the device-resident u */
    double dt_gpu = 0.0;
    CHECK_RC(trixi_b200_upload(h, 0, u0));
    CHECK_RC(trixi_b200_max_dt(h, 0.0, &dt_gpu));
    const double dt_ref = oracle_max_dt(&desc, u0);
    const double err_dt = fabs(dt_gpu - dt_ref) / dt_ref;

    /* 3. two CarpenterKennedy2N54 steps (methods_2N.jl:47-64), fused CFL on: max_dt after the step costs no pass */
    const double a[5] = {0.0, 567301805773.0 / 1357537059087.0, 2404267990393.0 / 2016746695238.0,
                         3550918686646.0 / 2091501179385.0, 1275806237668.0 / 842570457699.0};
    const double b[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                         1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                         2277821191437.0 / 14882151754819.0};
    const double c[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                         2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
    CHECK_RC(trixi_b200_set_option(h, TRIXI_B200_OPT_FUSED_CFL, 1));
    memcpy(u_ref, u0, sizeof(double) * len);
    double t = 0.0, dt = 1.3 * dt_gpu, err_dt2 = 0.0;
    for (int step = 0; step < 2; ++step) {
        CHECK_RC(trixi_b200_step_2n(h, t, dt, a, b, c, 5));
        oracle_step_2n(&desc, u_ref, du_ref, ut_ref, t, dt, a, b, c, 5, iu, NULL, sfv);
        t += dt;
        double dt_next = 0.0;
        CHECK_RC(trixi_b200_max_dt(h, t, &dt_next));
        const double e = fabs(dt_next - oracle_max_dt(&desc, u_ref)) / dt_next;
        if (e > err_dt2) err_dt2 = e;
        dt = 1.3 * dt_next;
    }
    CHECK_RC(trixi_b200_download(h, 0, u));
    const double err_u = rel_err(u, u_ref, len);
    const long long launches = (long long)trixi_b200_launch_count(h);
    trixi_b200_destroy(h);

    printf("c_abi_smoke: rhs %.3e  max_dt %.3e  step_2n u %.3e  max_dt after steps %.3e  launches %lld\n", err_rhs,
           err_dt, err_u, err_dt2, launches);
    /* tolerances: BASELINE.json (1e-12 per RHS); max_dt 1e-14 (tests/test_gpu_parity.py); two steps 1e-13 */
    const int ok = err_rhs <= 1e-12 && err_dt <= 1e-14 && err_u <= 1e-13 && err_dt2 <= 1e-13 && launches > 0;
    printf(ok ? "c_abi_smoke: OK\n" : "c_abi_smoke: FAILED\n");
    return ok ? 0 : 1;
}
