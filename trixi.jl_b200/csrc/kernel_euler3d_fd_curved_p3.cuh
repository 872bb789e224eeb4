// Tuned element kernel for curved meshes (StructuredMesh and P4estMesh, 3D): compressible Euler, polydeg 3,
// flux-differencing volume integral with flux_ranocha along averaged contravariant vectors
// (flux_differencing_kernel! dgsem_structured/dg_3d.jl:94-175; flux_ranocha(u_ll, u_rr, normal_direction)
// compressible_euler_3d.jl:795-828), fused with the surface integral (dg_3d.jl:1337-1394 / dgsem_p4est/dg_3d.jl:
// 976-1034), the nodal Jacobian (dgsem_structured/dg_3d.jl:937-956), source terms and the 2N Runge-Kutta stage.
// This is the p = 3 instance of the one GPU kernel the reference itself has (dgsem_p4est/dg_3d_gpu.jl:30-136:
// one thread per node, every two-point flux evaluated twice, atomics-free but unfused).
//
// Same skeleton as the TreeMesh headline kernel (kernel_euler3d_fd_p3.cuh): one warp = two elements, a thread owns
// one whole line of 4 nodes per direction pass and evaluates its 6 node pairs once each, tiles take turns in one
// shared-memory region.  Differences:
//  * the node record carries the three contravariant vectors next to the hoisted primitive variables (16 doubles,
//    stride 17 at the swizzled node position: conflict-free in all three passes); a pass uses Ja^d of both ends,
//    and the flux is evaluated along Ja^d_a + Ja^d_b -- the factor 1/2 of the average joins the other powers of two
//    in the D_split weights (D/4 for the density flux, D/8 for the others: exact);
//  * the velocity is not rotated (the direction is a general vector), so the momentum accumulators stay in place;
//  * the Jacobian is nodal, and P4estMesh adds all six face fluxes with "+" (outward normals).
#pragma once
#include <cstdint>

#include "kernel_euler3d_fd_p3.cuh"

namespace tb {

struct CurvedCfg {
    static constexpr int EPB = 2, THREADS = 32;
    static constexpr int CONS = 320, REC = 64 * kNPC_STRIDE, SFV = 480;  // doubles per element
    // flux passes: [records 0..2176 | du tiles 2176..2816]; before the x pass the du tiles' storage holds u.
    // epilogue:    [surface_flux_values 0..960 | u_tmp 2176..2816 | b dt u_tmp 960..1600 (not resident)]
    static constexpr int OFF_DU = EPB * REC, REGION = OFF_DU + EPB * CONS;
    static constexpr size_t SMEM_STREAM = sizeof(double) * REGION + 32;
    static constexpr size_t SMEM_RESIDENT = SMEM_STREAM + sizeof(double) * EPB * CONS;
    static constexpr int MIN_BLOCKS = 9;
    static constexpr int blocks_per_sm(bool resident) { return resident ? 8 : 9; }
};

// GEN: also the 3S* / SSP stage updates (KParams::mode 2, 3); see kernel_euler3d_fd_p3.cuh
template <bool WITH_SURFACE, bool GEN = false>
__global__ void __launch_bounds__(CurvedCfg::THREADS, CurvedCfg::MIN_BLOCKS)
    k_element_euler3d_ranocha_curved_p3(const KParams P) {
    using C = CurvedCfg;
    constexpr int CONS = C::CONS, REC = C::REC, SFV = C::SFV, EPB = C::EPB;
    extern __shared__ __align__(128) double smem[];
    const bool have_src = WITH_SURFACE && P.source_terms != TRIXI_B200_SRC_NONE;
    const bool resident = tuned_u_resident(P, WITH_SURFACE);
    double *s_rec = smem;                 // [2][64][17] swizzled
    double *s_du = smem + C::OFF_DU;      // [2][64][5] swizzled; before the x pass: u, natural order (not resident)
    double *s_sfv = smem;                 // epilogue: [2][6][16][5] natural
    double *s_inc = smem + EPB * SFV;     // then b dt u_tmp [2][64][5] natural (not resident)
    double *s_ut = smem + C::OFF_DU;      // epilogue: [2][64][5] natural: u_tmp in, u_tmp (or du) out
    const uint32_t bar_u = smem_u32(smem + C::REGION), bar_s = bar_u + 8, bar_t = bar_u + 16;
    double *s_u = smem + C::REGION + 4;   // resident only

    const int lane = threadIdx.x;
    const int t = lane & 15;
    const long long e0 = (long long)EPB * blockIdx.x;
    const int nvalid = (int)min((long long)EPB, P.nelements - e0);
    const int eh = (lane >> 4) < nvalid ? (lane >> 4) : 0;  // odd tail: the second half-warp mirrors the first
    const long long e = e0 + eh;
    const double gamma = P.eq.p[0], igm1 = P.eq.p[1];
    const bool rk = P.mode != 0;
    const bool need_ut = rk && P.rk_read_tmp;
    const uint32_t bu = nvalid * CONS * sizeof(double), bs = nvalid * SFV * sizeof(double);

    if (lane == 0) {
        mbar_init(bar_u, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_t, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    double *const s_uin = resident ? s_u : s_du;
    if (lane == 0) {
        mbar_expect_tx(bar_u, bu);
        tma_load(smem_u32(s_uin), P.u + e0 * CONS, bu, bar_u);
        if (WITH_SURFACE) tma_prefetch_l2(P.sfv + e0 * SFV, bs);
        if (need_ut) tma_prefetch_l2(P.u_tmp + e0 * CONS, bu);
        const long long en = e0 + P.prefetch_distance;
        if (P.prefetch_distance > 0 && en + EPB <= P.nelements) {
            tma_prefetch_l2(P.u + en * CONS, EPB * CONS * sizeof(double));
            tma_prefetch_l2(P.contravariant_vectors + en * 576, EPB * 576 * sizeof(double));
        }
    }
    const double *const su = s_uin + eh * CONS;
    double *const sr = s_rec + eh * REC, *const sd = s_du + eh * CONS;
    const double *const ja_e = P.contravariant_vectors + e * 576;  // [3 (dim), 3 (index), 64]: 9 doubles per node
    while (!mbar_try_wait(bar_u, 0)) {
    }

    // 1. node records for nodes t, t + 16, t + 32, t + 48: hoisted primitive variables and the contravariant vectors
#pragma unroll 2
    for (int r = 0; r < 4; ++r) {
        const int n = t + 16 * r;
        const double *c = su + n * 5;
        const double rho = c[0], m1 = c[1], m2 = c[2], m3 = c[3];
        double ja[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) ja[q] = ja_e[n * 9 + q];
        const double inv_rho = fast_rcp(rho);
        double v1 = m1 * inv_rho, v2 = m2 * inv_rho, v3 = m3 * inv_rho;
        v1 = fma(fma(-rho, v1, m1), inv_rho, v1);
        v2 = fma(fma(-rho, v2, m2), inv_rho, v2);
        v3 = fma(fma(-rho, v3, m3), inv_rho, v3);
        const double pr = (gamma - 1) * (c[4] - 0.5 * (m1 * v1 + m2 * v2 + m3 * v3));
        const double lr = log_pos(rho);
        double *o = sr + (16 * r + (t ^ (5 * r))) * kNPC_STRIDE;  // swz_pos(t + 16 r)
        o[0] = rho;
        o[1] = v1;
        o[2] = v2;
        o[3] = v3;
        o[4] = pr + pr;
        o[5] = lr;
        o[6] = lr - log_pos(pr);
#pragma unroll
        for (int q = 0; q < 9; ++q) o[7 + q] = ja[q];
    }
    __syncwarp();

    // 2. direction passes: line (j, k) / (i, k) / (i, j) = (t & 3, t >> 2); node m at position B ^ (m * M)
    const int a0 = t & 3, a1 = t >> 2;
    const int B0 = 16 * a1 + 4 * (a0 ^ a1) + a1, B1 = 20 * a1 + (a0 ^ a1);
    double acc[4][5];
    int pos[4];
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        const int B = d == 0 ? B0 : (d == 1 ? B1 : t), M = d == 0 ? 1 : (d == 1 ? 4 : 21);
        double q[4][10];  // rho, v1, v2, v3, 2 p, log rho, log rho - log p, Ja^d
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            pos[m] = B ^ (m * M);
            const double *src = sr + pos[m] * kNPC_STRIDE;
#pragma unroll
            for (int c = 0; c < 7; ++c) q[m][c] = src[c];
#pragma unroll
            for (int c = 0; c < 3; ++c) q[m][7 + c] = src[7 + 3 * d + c];
        }
        if (WITH_SURFACE && d == 2) {
            // the records have been read for the last time: the surface fluxes land in their storage
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(bar_s, bs);
                tma_load(smem_u32(s_sfv), P.sfv + e0 * SFV, bs, bar_s);
            }
        }
        double g[5];
        // (the pass-local record q keeps Ja^d at slots 7..9: direction index 0 for ranocha_pair_normal)
#define TB_PAIR(a, b, FIRST_A, FIRST_B)                                                                        \
    ranocha_pair_normal(q[a], q[b], 0, igm1, g);                                                               \
    acc[a][0] = FIRST_A ? P.dsplit_q[a + 4 * b] * g[0] : fma(P.dsplit_q[a + 4 * b], g[0], acc[a][0]);          \
    acc[b][0] = FIRST_B ? P.dsplit_q[b + 4 * a] * g[0] : fma(P.dsplit_q[b + 4 * a], g[0], acc[b][0]);          \
    _Pragma("unroll") for (int v = 1; v < 5; ++v) {                                                            \
        acc[a][v] = FIRST_A ? P.dsplit_e[a + 4 * b] * g[v] : fma(P.dsplit_e[a + 4 * b], g[v], acc[a][v]);      \
        acc[b][v] = FIRST_B ? P.dsplit_e[b + 4 * a] * g[v] : fma(P.dsplit_e[b + 4 * a], g[v], acc[b][v]);      \
    }
        TB_PAIR(0, 1, true, true)
        TB_PAIR(2, 3, true, true)
        TB_PAIR(0, 2, false, false)
        TB_PAIR(1, 3, false, false)
        TB_PAIR(0, 3, false, false)
        TB_PAIR(1, 2, false, false)
#undef TB_PAIR
        if (d == 0) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                double *o = sd + pos[m] * 5;
#pragma unroll
                for (int v = 0; v < 5; ++v) o[v] = acc[m][v];
            }
            __syncwarp();
        } else if (d == 1) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                double *o = sd + pos[m] * 5;
#pragma unroll
                for (int v = 0; v < 5; ++v) o[v] += acc[m][v];
            }
            __syncwarp();
        }
    }

    // 3. finish the z line (i, j) = (a0, a1), nodes n = t + 16 k, in registers
    double val[4][5];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double *o = sd + pos[k] * 5;
#pragma unroll
        for (int v = 0; v < 5; ++v) val[k][v] = o[v] + acc[k][v];
    }
    fence_proxy_async();
    __syncwarp();
    if (need_ut && lane == 0) {
        mbar_expect_tx(bar_t, bu);
        tma_load(smem_u32(s_ut), P.u_tmp + e0 * CONS, bu, bar_t);
    }
    if (WITH_SURFACE) {
        while (!mbar_try_wait(bar_s, 0)) {
        }
    }
    {
        const int i = a0, j = a1;
        if constexpr (WITH_SURFACE) {
            // calc_surface_integral!: StructuredMesh "-" on the negative faces (dg_3d.jl:1337-1394), P4estMesh "+" on
            // all six (fluxes along outward normals, dgsem_p4est/dg_3d.jl:976-1034)
            const double wneg = P.p4est ? P.inv_weight0 : -P.inv_weight0;
            const double *ssf = s_sfv + eh * SFV;
            if (i == 0 || i == 3) {
                const double *sf = ssf + (i == 0 ? 0 : 80) + j * 5;
                const double w = i == 0 ? wneg : P.inv_weight0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[k][v] = fma(sf[20 * k + v], w, val[k][v]);
            }
            if (j == 0 || j == 3) {
                const double *sf = ssf + (j == 0 ? 160 : 240) + i * 5;
                const double w = j == 0 ? wneg : P.inv_weight0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[k][v] = fma(sf[20 * k + v], w, val[k][v]);
            }
            {
                const double *sf = ssf + 320 + t * 5;
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    val[0][v] = fma(sf[v], wneg, val[0][v]);
                    val[3][v] = fma(sf[80 + v], P.inv_weight0, val[3][v]);
                }
            }
            // apply_jacobian! with the nodal inverse Jacobian (dgsem_structured/dg_3d.jl:937-956)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double factor = -P.inverse_jacobian[e * 64 + t + 16 * k];
#pragma unroll
                for (int v = 0; v < 5; ++v) val[k][v] *= factor;
            }
            if (have_src) {  // calc_sources! (dg_3d.jl:1417-1437)
                const Euler<3> eq(P.eq);
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const int n = t + 16 * k;
                    double un[5], x[3], sv[5];
#pragma unroll
                    for (int v = 0; v < 5; ++v) un[v] = su[n * 5 + v];
#pragma unroll
                    for (int dd = 0; dd < 3; ++dd) x[dd] = P.node_coordinates[(e * 64 + n) * 3 + dd];
                    eq.source_terms(P.source_terms, un, x, P.t, sv);
#pragma unroll
                    for (int v = 0; v < 5; ++v) {
                        val[0][v] += k == 0 ? sv[v] : 0.0;
                        val[1][v] += k == 1 ? sv[v] : 0.0;
                        val[2][v] += k == 2 ? sv[v] : 0.0;
                        val[3][v] += k == 3 ? sv[v] : 0.0;
                    }
                }
            }
        }
        double *const sut = s_ut + eh * CONS;
        if (need_ut) {
            while (!mbar_try_wait(bar_t, 0)) {
            }
        }
        if (!rk) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int v = 0; v < 5; ++v) sut[(t + 16 * k) * 5 + v] = val[k][v];
        } else {
            // 2N stage (methods_2N.jl:152-158), arithmetic as in the TreeMesh kernel
            const bool rk2n = !GEN || P.mode == 1;  // (modes 2 and 3, the 3S* and SSP stages, always run resident)
            if (need_ut && rk2n) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[k][v] -= sut[(t + 16 * k) * 5 + v] * P.rk_a;
            }
            if (rk2n) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) sut[(t + 16 * k) * 5 + v] = val[k][v];
            }
            if (!resident) {
                double *const sinc = s_inc + eh * CONS;  // (behind the face tiles: no wait for their readers)
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) sinc[(t + 16 * k) * 5 + v] = __dmul_rn(val[k][v], P.rk_b_dt);
            } else {
                double *const suo = s_u + eh * CONS;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) {
                        double *out_u = suo + (t + 16 * k) * 5 + v;
                        if (rk2n) {
                            *out_u = __dadd_rn(*out_u, __dmul_rn(val[k][v], P.rk_b_dt));
                        } else {  // 3S* / SSP stage (KParams::mode 2, 3): u_tmp2 straight from global memory
                            double *out_t = sut + (t + 16 * k) * 5 + v;
                            double xn;
                            *out_u = rk_stage_3s_ssp(P, val[k][v], need_ut ? *out_t : 0.0, *out_u,
                                                     P.mode == 2 ? P.u_tmp2[e * CONS + (t + 16 * k) * 5 + v] : 0.0, xn);
                            *out_t = xn;
                        }
                    }
            }
        }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (!rk) {
            tma_store(P.du + e0 * CONS, smem_u32(s_ut), bu);
        } else {
            if (!GEN || P.rk_write_tmp) tma_store(P.u_tmp + e0 * CONS, smem_u32(s_ut), bu);
            if (resident)
                tma_store(P.u_out + e0 * CONS, smem_u32(s_u), bu);
            else
                tma_reduce_add_f64(P.u_out + e0 * CONS, smem_u32(s_inc), bu);
        }
        tma_store_commit_and_wait_read();
    }
}

cudaError_t preload_tuned_euler3d_curved() {
    cudaError_t e = preload_kernel(k_element_euler3d_ranocha_curved_p3<true>);
    if (e != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_ranocha_curved_p3<true, true>)) != cudaSuccess) return e;
    return preload_kernel(k_element_euler3d_ranocha_curved_p3<false>);
}

cudaError_t launch_element_euler3d_ranocha_curved_p3(const KParams &P, bool with_surface, cudaStream_t s) {
    using C = CurvedCfg;
    static PerDeviceFlag configured;
    if (!configured.test_and_set()) {
        cudaError_t err = cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_p3<true>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout,
                                               cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_p3<false>,
                                   cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_p3<true, true>,
                                   cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
    }
    const unsigned blocks = (unsigned)((P.nelements + C::EPB - 1) / C::EPB);
    KParams Q = P;
    Q.want_cfl = 0;  // (the CFL reduction of curved meshes stays in k_max_dt_curved: fuses_cfl() says so)
    const bool resident = tuned_u_resident(Q, with_surface);
    const size_t smem = resident ? C::SMEM_RESIDENT : C::SMEM_STREAM;
    if (Q.prefetch_distance < 0) Q.prefetch_distance = C::EPB * C::blocks_per_sm(resident) * Q.sm_count;
    if (Q.mode > 1)  // 3S* / SSP stage (always with the surface terms)
        k_element_euler3d_ranocha_curved_p3<true, true><<<blocks, C::THREADS, smem, s>>>(Q);
    else if (with_surface)
        k_element_euler3d_ranocha_curved_p3<true><<<blocks, C::THREADS, smem, s>>>(Q);
    else
        k_element_euler3d_ranocha_curved_p3<false><<<blocks, C::THREADS, smem, s>>>(Q);
    return cudaSuccess;
}

}  // namespace tb
