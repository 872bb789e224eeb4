// Tuned element kernel: 3D compressible Euler, polydeg 3, weak-form volume integral (weak_form_kernel!
// dg_3d.jl:133-164), fused with surface integral, Jacobian, source terms and the 2N Runge-Kutta stage.
// BASELINE.json config 2 (elixir_euler_source_terms.jl: weak form + flux_lax_friedrichs + source terms).
//
// The kernel is HBM-bound (244 B/DOF with source terms against 149 flop/DOF); the generic one-thread-per-node
// kernel waits on its global loads (ncu: long-scoreboard 19 warp-cycles per issue, 48% of DRAM peak).  Here, as
// in the flux-differencing kernel (kernel_euler3d_fd_p3.cuh):
//  * one warp = one CTA = one element, 16-18 CTAs per SM, no block barriers;
//  * the element's u, u_tmp, surface_flux_values and node_coordinates records are contiguous and arrive by
//    cp.async.bulk on one mbarrier, results leave by cp.async.bulk stores: 8.9-10.5 KB in flight per warp with
//    no per-thread address arithmetic, and the next wave's records are prefetched into L2;
//  * each thread owns the nodes n = lane and lane + 32.  Per direction the nodal fluxes of the element go to one
//    shared tile [64][5] (natural order, odd record stride), then every node contracts its line with its row of
//    D_hat held in registers: du[:, node] += sum_l D_hat[idx_d, l] f_d[:, line(l)].
#pragma once
#include "tile_io.cuh"

namespace tb {

struct WeakCfg {
    static constexpr int THREADS = 32;
    static constexpr int CONS = 320, SFV = 480, XYZ = 192, JA = 576, IJ = 64;  // doubles per element
    // s_u, s_ut (TMA in/out), s_sfv, s_f (one direction's nodal fluxes), mbarrier, [s_ja, s_ij], [s_x]
    static constexpr size_t smem(bool with_sources, bool curved) {
        return sizeof(double) * (3 * CONS + SFV + (curved ? JA + IJ : 0) + (with_sources ? XYZ : 0)) + 16;
    }
    static constexpr int MIN_BLOCKS = 18;
    static constexpr int blocks_per_sm(bool with_sources, bool curved) {
        return (int)((227 * 1024 + 1024) / (smem(with_sources, curved) + 1024));  // 1 KB per CTA is reserved
    }
};

// CURVED: StructuredMesh / P4estMesh (weak_form_kernel! dgsem_structured/dg_3d.jl:36-89): contravariant fluxes
// Ja^a . f, nodal inverse Jacobian (apply_jacobian! :937-956), P4est's all-plus surface integral
// (dgsem_p4est/dg_3d.jl:976-1034); the element's contravariant_vectors [3, 3, 64] and inverse_jacobian [64]
// records are two more bulk loads.
// GEN: also the 3S* / SSP stage updates (KParams::mode 2, 3); see kernel_euler3d_fd_p3.cuh
template <bool WITH_SURFACE, bool CURVED, bool GEN = false>
__global__ void __launch_bounds__(WeakCfg::THREADS, CURVED ? 12 : WeakCfg::MIN_BLOCKS)
    k_element_euler3d_weak_p3(const KParams P) {
    using C = WeakCfg;
    constexpr int CONS = C::CONS, SFV = C::SFV, XYZ = C::XYZ, JA = C::JA, IJ = C::IJ;
    extern __shared__ __align__(128) double smem[];
    double *s_u = smem;           // [64][5] u in, updated u out
    double *s_ut = s_u + CONS;    // [64][5] u_tmp in, u_tmp (or du) out
    double *s_sfv = s_ut + CONS;  // [6][16][5]
    double *s_f = s_sfv + SFV;    // [64][5] nodal fluxes of the current direction
    const uint32_t bar = smem_u32(s_f + CONS);
    double *s_ja = s_f + CONS + 2;               // curved: [64][3 (index)][3 (dim)]
    double *s_ij = s_ja + JA;                    // curved: [64]
    double *s_x = CURVED ? s_ij + IJ : s_ja;     // [64][3] node coordinates (source terms only)

    const int lane = threadIdx.x;
    const long long e = P.elem_begin + blockIdx.x;
    const double gamma = P.eq.p[0];
    const bool rk = P.mode != 0;
    const bool need_ut = rk && P.rk_read_tmp;
    const bool have_src = WITH_SURFACE && P.source_terms != TRIXI_B200_SRC_NONE;

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        constexpr uint32_t bu = CONS * sizeof(double), bs = SFV * sizeof(double), bx = XYZ * sizeof(double);
        constexpr uint32_t bj = JA * sizeof(double), bi = IJ * sizeof(double);
        mbar_expect_tx(bar, bu + (need_ut ? bu : 0u) + (WITH_SURFACE ? bs : 0u) + (have_src ? bx : 0u) +
                                (CURVED ? bj + (WITH_SURFACE ? bi : 0u) : 0u));
        if constexpr (CURVED) {
            tma_load(smem_u32(s_ja), P.contravariant_vectors + e * JA, bj, bar);
            if (WITH_SURFACE) tma_load(smem_u32(s_ij), P.inverse_jacobian + e * IJ, bi, bar);
        }
        tma_load(smem_u32(s_u), P.u + e * CONS, bu, bar);
        if (need_ut) tma_load(smem_u32(s_ut), P.u_tmp + e * CONS, bu, bar);
        if (WITH_SURFACE) tma_load(smem_u32(s_sfv), P.sfv + e * SFV, bs, bar);
        if (have_src) tma_load(smem_u32(s_x), P.node_coordinates + e * XYZ, bx, bar);
        const long long en = e + P.prefetch_distance;
        if (P.prefetch_distance > 0 && en < P.nelements) {
            tma_prefetch_l2(P.u + en * CONS, bu);
            if (need_ut) tma_prefetch_l2(P.u_tmp + en * CONS, bu);
            if (WITH_SURFACE) tma_prefetch_l2(P.sfv + en * SFV, bs);
            if (have_src) tma_prefetch_l2(P.node_coordinates + en * XYZ, bx);
            if constexpr (CURVED) {
                tma_prefetch_l2(P.contravariant_vectors + en * JA, bj);
                if (WITH_SURFACE) tma_prefetch_l2(P.inverse_jacobian + en * IJ, bi);
            }
        }
    }
    // node n = lane + 32 r: i = lane & 3, j = (lane >> 2) & 3, k = (lane >> 4) + 2 r.  Rows of D_hat
    // (column-major [4, 4]) for the thread's line positions, fetched while the tiles are in flight.
    const int i = lane & 3, j = (lane >> 2) & 3, k0 = lane >> 4;
    double di[4], dj[4], dk[2][4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        di[l] = P.dhat[i + 4 * l];
        dj[l] = P.dhat[j + 4 * l];
        dk[0][l] = P.dhat[k0 + 4 * l];
        dk[1][l] = P.dhat[k0 + 2 + 4 * l];
    }
    while (!mbar_try_wait(bar, 0)) {
    }

    // cons2prim (compressible_euler_3d.jl:1783-1793) once per node; flux(u, orientation) (:420-447) per direction
    double mom[2][3], vel[2][3], pr[2], ep[2], acc[2][5];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double *c = s_u + (lane + 32 * r) * 5;
        const double rho = c[0];
        double kin = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mom[r][d] = c[1 + d];
            vel[r][d] = c[1 + d] / rho;
            kin += c[1 + d] * vel[r][d];
        }
        pr[r] = (gamma - 1) * (c[4] - 0.5 * kin);
        ep[r] = c[4] + pr[r];
#pragma unroll
        for (int v = 0; v < 5; ++v) acc[r][v] = 0.0;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            double *f = s_f + (lane + 32 * r) * 5;
            if constexpr (!CURVED) {
                const double rv = mom[r][d];
                f[0] = rv;
                f[1] = rv * vel[r][0] + (d == 0 ? pr[r] : 0.0);
                f[2] = rv * vel[r][1] + (d == 1 ? pr[r] : 0.0);
                f[3] = rv * vel[r][2] + (d == 2 ? pr[r] : 0.0);
                f[4] = ep[r] * vel[r][d];
            } else {
                // contravariant flux Ja^d . (f1, f2, f3) (dgsem_structured/dg_3d.jl:52-84), summed in the order
                // of the dimensions like the generic kernel
                const double *ja = s_ja + (lane + 32 * r) * 9 + 3 * d;
                double sum[5];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double rv = mom[r][q];
                    const double fq[5] = {rv, rv * vel[r][0] + (q == 0 ? pr[r] : 0.0),
                                          rv * vel[r][1] + (q == 1 ? pr[r] : 0.0),
                                          rv * vel[r][2] + (q == 2 ? pr[r] : 0.0), ep[r] * vel[r][q]};
#pragma unroll
                    for (int v = 0; v < 5; ++v) sum[v] = q == 0 ? ja[0] * fq[v] : sum[v] + ja[q] * fq[v];
                }
#pragma unroll
                for (int v = 0; v < 5; ++v) f[v] = sum[v];
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int n = lane + 32 * r;
            const int idx = d == 0 ? i : (d == 1 ? j : k0 + 2 * r);
            const int stride = d == 0 ? 1 : (d == 1 ? 4 : 16);
            const double *line = s_f + (n - idx * stride) * 5;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const double w = d == 0 ? di[l] : (d == 1 ? dj[l] : dk[r][l]);
                const double *f = line + l * stride * 5;
#pragma unroll
                for (int v = 0; v < 5; ++v) acc[r][v] = fma(w, f[v], acc[r][v]);
            }
        }
        __syncwarp();
    }

    const double factor_tree = (WITH_SURFACE && !CURVED) ? -P.inverse_jacobian[e] : 1.0;
    // Tree/Structured: "-" on the negative faces; P4est: "+" everywhere (fluxes along outward normals)
    const double w_neg = (CURVED && P.p4est) ? P.inv_weight0 : -P.inv_weight0;
    unsigned long long cfl0 = 0ull, cfl1 = 0ull, cfl2 = 0ull;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int k = k0 + 2 * r;
        const int n = lane + 32 * r;
        double(&val)[5] = acc[r];
        if constexpr (WITH_SURFACE) {
            // calc_surface_integral! (dg_3d.jl:1337-1394): directions 1..6 = -x,+x,-y,+y,-z,+z
            if (i == 0 || i == 3) {
                const double *sf = s_sfv + ((i == 0 ? 0 : 1) * 16 + j + 4 * k) * 5;
                const double w = i == 0 ? w_neg : P.inv_weight0;
#pragma unroll
                for (int v = 0; v < 5; ++v) val[v] = fma(sf[v], w, val[v]);
            }
            if (j == 0 || j == 3) {
                const double *sf = s_sfv + ((j == 0 ? 2 : 3) * 16 + i + 4 * k) * 5;
                const double w = j == 0 ? w_neg : P.inv_weight0;
#pragma unroll
                for (int v = 0; v < 5; ++v) val[v] = fma(sf[v], w, val[v]);
            }
            if (k == 0 || k == 3) {
                const double *sf = s_sfv + ((k == 0 ? 4 : 5) * 16 + i + 4 * j) * 5;
                const double w = k == 0 ? w_neg : P.inv_weight0;
#pragma unroll
                for (int v = 0; v < 5; ++v) val[v] = fma(sf[v], w, val[v]);
            }
            // apply_jacobian! (dg_3d.jl:1396-1414; curved: nodal, dgsem_structured/dg_3d.jl:937-956)
            const double factor = CURVED ? -s_ij[n] : factor_tree;
#pragma unroll
            for (int v = 0; v < 5; ++v) val[v] *= factor;
            // calc_sources! (dg_3d.jl:1417-1437)
            if (have_src) {
                const Euler<3> eq(P.eq);
                double un[5], x[3], sv[5];
#pragma unroll
                for (int v = 0; v < 5; ++v) un[v] = s_u[n * 5 + v];
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) x[dd] = s_x[n * 3 + dd];
                eq.source_terms(P.source_terms, un, x, P.t, sv);
#pragma unroll
                for (int v = 0; v < 5; ++v) val[v] += sv[v];
            }
        }
        double *out_t = s_ut + n * 5;
        if (!rk) {
#pragma unroll
            for (int v = 0; v < 5; ++v) out_t[v] = val[v];
        } else {
            // 2N stage (methods_2N.jl:152-158): u_tmp = du - u_tmp * a; u += u_tmp * (b * dt)
            double *out_u = s_u + n * 5;
            double un[5];
            if (!GEN || P.mode == 1) {
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const double tmp = need_ut ? val[v] - out_t[v] * P.rk_a : val[v];
                    out_t[v] = tmp;
                    un[v] = out_u[v] + tmp * P.rk_b_dt;
                    out_u[v] = un[v];
                }
            } else {
                // 3S* / SSP stage (KParams::mode 2, 3): u_tmp2 comes straight from global memory (40-byte node
                // records, consecutive lanes = consecutive records)
                const double *u2 = P.u_tmp2 + (e * 64 + n) * 5;
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    double xn;
                    un[v] = rk_stage_3s_ssp(P, val[v], need_ut ? out_t[v] : 0.0, out_u[v], P.mode == 2 ? u2[v] : 0.0, xn);
                    out_t[v] = xn;
                    out_u[v] = un[v];
                }
            }
            if (P.want_cfl) {
                // max_dt of the updated state (stepsize_dg3d.jl:8-32; curved :94-123), as in the flux-differencing kernel
                const double rho = un[0], inv_rho = fast_rcp(rho);
                double v1 = un[1] * inv_rho, v2 = un[2] * inv_rho, v3 = un[3] * inv_rho;
                v1 = fma(fma(-rho, v1, un[1]), inv_rho, v1);
                v2 = fma(fma(-rho, v2, un[2]), inv_rho, v2);
                v3 = fma(fma(-rho, v3, un[3]), inv_rho, v3);
                const double p_new = (gamma - 1) * (un[4] - 0.5 * (un[1] * v1 + un[2] * v2 + un[3] * v3));
                const double gp = gamma * p_new;
                double c2 = gp * inv_rho;
                c2 = fma(fma(-rho, c2, gp), inv_rho, c2);
                const double c = sqrt(c2);
                const double lam[3] = {fabs(v1) + c, fabs(v2) + c, fabs(v3) + c};
                if constexpr (!CURVED) {
                    cfl0 = max(cfl0, cfl_encode(lam[0]));
                    cfl1 = max(cfl1, cfl_encode(lam[1]));
                    cfl2 = max(cfl2, cfl_encode(lam[2]));
                } else {
                    // |inverse_jacobian| |Ja^a . lambda| per node and direction a (the nodal Jacobian enters here)
                    const double *ja = s_ja + n * 9;
                    const double aij = fabs(s_ij[n]);
                    cfl0 = max(cfl0, cfl_encode(aij * fabs(ja[0] * lam[0] + ja[1] * lam[1] + ja[2] * lam[2])));
                    cfl1 = max(cfl1, cfl_encode(aij * fabs(ja[3] * lam[0] + ja[4] * lam[1] + ja[5] * lam[2])));
                    cfl2 = max(cfl2, cfl_encode(aij * fabs(ja[6] * lam[0] + ja[7] * lam[1] + ja[8] * lam[2])));
                }
            }
        }
    }
    if (rk && P.want_cfl) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            cfl0 = max(cfl0, __shfl_xor_sync(0xffffffffu, cfl0, off));
            cfl1 = max(cfl1, __shfl_xor_sync(0xffffffffu, cfl1, off));
            cfl2 = max(cfl2, __shfl_xor_sync(0xffffffffu, cfl2, off));
        }
        if (lane == 0) {
            double sum = 0.0;
            sum += __longlong_as_double((long long)cfl0);
            sum += __longlong_as_double((long long)cfl1);
            sum += __longlong_as_double((long long)cfl2);
            atomicMax(P.cfl_key + (blockIdx.x & (kCflSlots - 1)), cfl_encode(CURVED ? sum : P.inverse_jacobian[e] * sum));
        }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        constexpr uint32_t bu = CONS * sizeof(double);
        if (!rk) {
            tma_store(P.du + e * CONS, smem_u32(s_ut), bu);
        } else {
            if (!GEN || P.rk_write_tmp) tma_store(P.u_tmp + e * CONS, smem_u32(s_ut), bu);
            tma_store(P.u_out + e * CONS, smem_u32(s_u), bu);
        }
        tma_store_commit_and_wait_read();
    }
}

cudaError_t preload_tuned_euler3d_weak() {
    cudaError_t e;
    if ((e = preload_kernel(k_element_euler3d_weak_p3<true, false>)) != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_weak_p3<false, false>)) != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_weak_p3<true, true>)) != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_weak_p3<true, false, true>)) != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_weak_p3<true, true, true>)) != cudaSuccess) return e;
    return preload_kernel(k_element_euler3d_weak_p3<false, true>);
}

template <bool WS, bool CURVED, bool GEN = false>
static cudaError_t launch_weak_variant(const KParams &P, cudaStream_t s) {
    using C = WeakCfg;
    static PerDeviceFlag configured;
    auto kern = k_element_euler3d_weak_p3<WS, CURVED, GEN>;
    if (!configured.test_and_set()) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                               cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
    }
    const unsigned blocks = (unsigned)(P.elem_end - P.elem_begin);
    const bool with_sources = WS && P.source_terms != TRIXI_B200_SRC_NONE;
    KParams Q = P;
    if (Q.prefetch_distance < 0) Q.prefetch_distance = C::blocks_per_sm(with_sources, CURVED) * Q.sm_count;
    kern<<<blocks, C::THREADS, C::smem(with_sources, CURVED), s>>>(Q);
    return cudaSuccess;
}

cudaError_t launch_element_euler3d_weak_p3(const KParams &P, bool with_surface, cudaStream_t s) {
    if (P.mode > 1)  // 3S* / SSP stage (always with the surface terms)
        return P.curved ? launch_weak_variant<true, true, true>(P, s) : launch_weak_variant<true, false, true>(P, s);
    if (P.curved)
        return with_surface ? launch_weak_variant<true, true>(P, s) : launch_weak_variant<false, true>(P, s);
    return with_surface ? launch_weak_variant<true, false>(P, s) : launch_weak_variant<false, false>(P, s);
}

}  // namespace tb
