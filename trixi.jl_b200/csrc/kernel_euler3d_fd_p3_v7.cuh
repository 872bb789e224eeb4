// Tuned element kernel: 3D compressible Euler, polydeg 3 (4^3 nodes), flux-differencing volume integral
// with flux_ranocha, fused with surface integral, Jacobian, source terms and the 2N Runge-Kutta stage.
// This is the headline configuration (BASELINE.json: 3D Euler EC p=3).
//
// Work decomposition.
//  * One warp = one CTA = one element; everything is warp-synchronous (no block barriers), 14 CTAs
//    resident per SM so one warp's tile I/O latency and FP64 dependency chains hide behind the others.
//  * Tile I/O is TMA: three `cp.async.bulk` loads (u, u_tmp, surface_flux_values of the two elements are
//    contiguous 5/5/7.5 KB records) signalled on an mbarrier, results leave through `cp.async.bulk`
//    stores -- no per-thread address arithmetic, no register staging, fully coalesced HBM traffic.
//  * Per element and direction the 64 nodes form 16 lines of 4 nodes with 6 symmetric node pairs each.
//    Two threads share a line per direction pass, three pair fluxes each, so every two-point flux is
//    evaluated exactly once (288 per element, like the reference's symmetric loop dg_3d.jl:177-211);
//    D_split[a,b] f is accumulated into both end nodes, partial sums meet in a shared-memory du tile, and
//    after the z pass each thread finishes surface integral, Jacobian, sources and the RK update of its
//    two nodes in registers.
//  * flux_ranocha is evaluated in the hoisted form of the reference's own SIMD kernel
//    (dg_3d_compressible_euler.jl:289-309,360-385): primitive variables and log(rho), log(p) once per
//    node, so the logarithmic means need no log per pair.
//  * The prim and du tiles are AoS records at a swizzled node position pos(i,j,k) = 16k + 4(j^k) + (i^k):
//    in every direction pass the 16 lines of an element hit 16 distinct 8-byte banks (record strides 5
//    and 7 are odd, so the map stays bijective); the TMA-filled buffers keep the global (natural) order,
//    which is conflict-free for the per-node passes.
#pragma once
#include <cstdint>

#include "tile_io.cuh"

namespace tb {

// Node record in the prim tile: rho, v1, v2, v3, p, log(rho), log(p)
constexpr int kNPv7 = 7;

// flux_ranocha(u_ll, u_rr, orientation) (compressible_euler_3d.jl:746-793) on hoisted node records whose
// velocity components have been rotated so that slot 1 is the normal one: (rho, vn, vt1, vt2, p, log rho,
// log p).  The output is rotated the same way: (f_rho, f_n, f_t1, f_t2, f_E).
TB_DEV void ranocha_pair_rot_v7(const double (&L)[kNPv7], const double (&R)[kNPv7], double inv_gm1, double (&f)[5]) {
    const double rho_ll = L[0], p_ll = L[4], rho_rr = R[0], p_rr = R[4];
    const double dlog_rho = R[5] - L[5];  // log(rho_rr / rho_ll)
    // ln_mean(rho_ll, rho_rr) (math.jl:198-210); f^2 = (x-y)^2/(x+y)^2 as in the reference's SIMD kernel
    double rho_mean;
    {
        const double sum = rho_ll + rho_rr, dif = rho_rr - rho_ll;
        const double f2 = (dif * dif) * rcp_1nr(sum * sum);
        const bool series = f2 < 1.0e-4;
        const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        rho_mean = fast_div(series ? sum : dif, series ? poly : dlog_rho);
    }
    // inv_ln_mean(rho_ll * p_rr, rho_rr * p_ll) (math.jl:238-250)
    double inv_rho_p_mean;
    {
        const double x = rho_ll * p_rr, y = rho_rr * p_ll;
        const double sum = x + y, dif = y - x;
        const double f2 = (dif * dif) * rcp_1nr(sum * sum);
        const bool series = f2 < 1.0e-4;
        const double poly = fma(f2, fma(f2, fma(f2, 2.0 / 7.0, 2.0 / 5.0), 2.0 / 3.0), 2.0);
        // log(y / x) = log(rho_rr p_ll) - log(rho_ll p_rr)
        const double m = fast_div(series ? poly : dlog_rho + (L[6] - R[6]), series ? sum : dif);
        inv_rho_p_mean = p_ll * p_rr * m;
    }
    const double vn_avg = 0.5 * (L[1] + R[1]), vt1_avg = 0.5 * (L[2] + R[2]), vt2_avg = 0.5 * (L[3] + R[3]);
    const double p_avg = 0.5 * (p_ll + p_rr);
    const double velocity_square_avg = 0.5 * (L[1] * R[1] + L[2] * R[2] + L[3] * R[3]);
    const double f1 = rho_mean * vn_avg;
    f[0] = f1;
    f[1] = f1 * vn_avg + p_avg;
    f[2] = f1 * vt1_avg;
    f[3] = f1 * vt2_avg;
    f[4] = f1 * (velocity_square_avg + inv_rho_p_mean * inv_gm1) + 0.5 * (p_ll * R[1] + p_rr * L[1]);
}

struct TunedCfgV7 {
    static constexpr int EPB = 1, THREADS = 32;  // one warp, one element, two threads per line
    static constexpr int CONS = 320, PRIM = 64 * kNPv7, SFV = 480;  // doubles per element
    // s_u (natural order, TMA), s_sfv (natural, TMA), s_du, s_prim (swizzled), mbarrier, [s_ut (natural, TMA)].
    // Without source terms the u_tmp tile is not resident during the flux passes: it is loaded into the prim
    // tile's storage once the z pass has read it for the last time, and leaves from there.  12.6 KB instead of
    // 15.1 KB per element: 17 instead of 14 resident warps per SM (shared memory is the occupancy limiter,
    // 105 registers per thread would allow 18).
    static constexpr size_t SMEM_DEFERRED = sizeof(double) * EPB * (2 * CONS + SFV + PRIM) + 16;
    static constexpr size_t SMEM_RESIDENT = SMEM_DEFERRED + sizeof(double) * EPB * CONS;
    static constexpr int MIN_BLOCKS = 17;
    static constexpr int blocks_per_sm(bool deferred) { return deferred ? 17 : 14; }
};

template <bool WITH_SURFACE>
__global__ void __launch_bounds__(TunedCfgV7::THREADS, TunedCfgV7::MIN_BLOCKS)
    k_element_euler3d_ranocha_p3_v7(const KParams P) {
    using C = TunedCfgV7;
    constexpr int CONS = C::CONS, PRIM = C::PRIM, SFV = C::SFV;
    extern __shared__ __align__(128) double smem[];
    const bool have_src = WITH_SURFACE && P.source_terms != TRIXI_B200_SRC_NONE;
    const bool deferred = !have_src;  // must match the launch's dynamic shared memory size
    double *s_u = smem;            // [64][5] natural: u in, updated u out
    double *s_sfv = s_u + CONS;    // [6][16][5] natural
    double *s_du = s_sfv + SFV;    // [64][5] swizzled
    double *s_prim = s_du + CONS;  // [64][7] swizzled; after the flux passes: source terms or the u_tmp tile
    const uint32_t bar = smem_u32(s_prim + PRIM);
    double *s_ut = deferred ? s_prim : s_prim + PRIM + 2;  // [64][5] natural: u_tmp in, u_tmp (or du) out

    const int lane = threadIdx.x;
    const long long e = P.elem_begin + blockIdx.x;
    const double gamma = P.eq.p[0], inv_gm1 = P.eq.p[1];
    const bool rk = P.mode != 0;
    const bool need_ut = rk && P.rk_read_tmp;

    // 0. TMA loads of the contiguous element records
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        constexpr uint32_t bu = CONS * sizeof(double), bs = SFV * sizeof(double);
        const bool ut_now = need_ut && !deferred;
        mbar_expect_tx(bar, bu + (ut_now ? bu : 0u) + (WITH_SURFACE ? bs : 0u));
        tma_load(smem_u32(s_u), P.u + e * CONS, bu, bar);
        if (ut_now) tma_load(smem_u32(s_ut), P.u_tmp + e * CONS, bu, bar);
        if (WITH_SURFACE) tma_load(smem_u32(s_sfv), P.sfv + e * SFV, bs, bar);
        // warm L2 for the element that will occupy this CTA slot next (blocks are scheduled in index order:
        // one wave further on), so its tile loads see L2 instead of HBM latency
        const long long en = e + P.prefetch_distance;
        if (P.prefetch_distance > 0 && en < P.nelements) {
            tma_prefetch_l2(P.u + en * CONS, bu);
            if (need_ut) tma_prefetch_l2(P.u_tmp + en * CONS, bu);
            if (WITH_SURFACE) tma_prefetch_l2(P.sfv + en * SFV, bs);
        }
    }
    // Two threads (h = 0, 1) share line l16 of every direction pass.  In line-local node numbering
    // thread h owns nodes lm[0], lm[1] and sees lm[2], lm[3] as foreign: h = 0: (0,1 | 2,3), h = 1: (3,2 | 0,1).
    // Both evaluate the pairs (lm0,lm1), (lm0,lm2), (lm1,lm3): together all 6 pairs of the line, once each.
    const int h = lane >> 4, l16 = lane & 15;
    const int a0 = l16 & 3, a1 = l16 >> 2;
    const int lm[4] = {h ? 3 : 0, h ? 2 : 1, h ? 0 : 2, h ? 1 : 3};
    // D_split[a, b] for the three pairs in both directions (column-major n x n)
    const double w01 = P.dsplit_c[lm[0] + 4 * lm[1]], w10 = P.dsplit_c[lm[1] + 4 * lm[0]];
    const double w02 = P.dsplit_c[lm[0] + 4 * lm[2]], w20 = P.dsplit_c[lm[2] + 4 * lm[0]];
    const double w13 = P.dsplit_c[lm[1] + 4 * lm[3]], w31 = P.dsplit_c[lm[3] + 4 * lm[1]];
    while (!mbar_try_wait(bar, 0)) {
    }

    // 1. cons2prim + logs for the two nodes (i, j, k = lm[0], lm[1]) this thread also finishes in step 3;
    //    a half-warp reads 16 consecutive node records: conflict-free
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const int n = l16 + 16 * (r == 0 ? lm[0] : lm[1]);
        const double *c = s_u + n * 5;
        const double rho = c[0];
        const double inv_rho = fast_rcp(rho);
        // v = rho_v / rho with a residual correction (cons2prim, compressible_euler_3d.jl:1783-1793)
        double v1 = c[1] * inv_rho, v2 = c[2] * inv_rho, v3 = c[3] * inv_rho;
        v1 = fma(fma(-rho, v1, c[1]), inv_rho, v1);
        v2 = fma(fma(-rho, v2, c[2]), inv_rho, v2);
        v3 = fma(fma(-rho, v3, c[3]), inv_rho, v3);
        const double pr = (gamma - 1) * (c[4] - 0.5 * (c[1] * v1 + c[2] * v2 + c[3] * v3));
        double *o = s_prim + swz_pos(n) * kNPv7;
        o[0] = rho;
        o[1] = v1;
        o[2] = v2;
        o[3] = v3;
        o[4] = pr;
        o[5] = log(rho);
        o[6] = log(pr);
    }
    __syncwarp();

    // 2. direction passes x, y, z with ONE copy of the flux code; the direction only enters through
    // shared-memory offsets (velocity slots rotated while loading, momentum slots while storing)
    int pos[4];
    double own[2][5], frn[2][5];
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        const int stride = 1 << (2 * d);
        const int base = d == 0 ? 4 * l16 : (d == 1 ? a0 + 16 * a1 : l16);
        const int on = 1 + d, ot1 = d == 2 ? 1 : 2 + d, ot2 = d == 0 ? 3 : d;  // 1 + (d + {0,1,2}) % 3
        double q[4][kNPv7];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            pos[m] = swz_pos(base + lm[m] * stride);
            const double *src = s_prim + pos[m] * kNPv7;
            q[m][0] = src[0];
            q[m][1] = src[on];
            q[m][2] = src[ot1];
            q[m][3] = src[ot2];
            q[m][4] = src[4];
            q[m][5] = src[5];
            q[m][6] = src[6];
        }
        double f[5];
        ranocha_pair_rot_v7(q[0], q[1], inv_gm1, f);
#pragma unroll
        for (int v = 0; v < 5; ++v) {
            own[0][v] = w01 * f[v];
            own[1][v] = w10 * f[v];
        }
        ranocha_pair_rot_v7(q[0], q[2], inv_gm1, f);
#pragma unroll
        for (int v = 0; v < 5; ++v) {
            own[0][v] = fma(w02, f[v], own[0][v]);
            frn[0][v] = w20 * f[v];
        }
        ranocha_pair_rot_v7(q[1], q[3], inv_gm1, f);
#pragma unroll
        for (int v = 0; v < 5; ++v) {
            own[1][v] = fma(w13, f[v], own[1][v]);
            frn[1][v] = w31 * f[v];
        }
        // every node receives one own and one foreign partial per pass; they meet in the du tile
        if (d < 2) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                double *t = s_du + pos[m] * 5;
                if (d == 0) {
                    t[0] = own[m][0];
                    t[on] = own[m][1];
                    t[ot1] = own[m][2];
                    t[ot2] = own[m][3];
                    t[4] = own[m][4];
                } else {
                    t[0] += own[m][0];
                    t[on] += own[m][1];
                    t[ot1] += own[m][2];
                    t[ot2] += own[m][3];
                    t[4] += own[m][4];
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double *t = s_du + pos[2 + m] * 5;
            t[0] += frn[m][0];
            t[on] += frn[m][1];
            t[ot1] += frn[m][2];
            t[ot2] += frn[m][3];
            t[4] += frn[m][4];
        }
        __syncwarp();
    }

    // the prim tile is dead now: fetch the u_tmp tile into its storage (second phase of the mbarrier); the
    // surface integral and the Jacobian below run while it is in flight
    const bool ut_late = need_ut && deferred;
    if (ut_late) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            constexpr uint32_t bu = CONS * sizeof(double);
            mbar_expect_tx(bar, bu);
            tma_load(smem_u32(s_ut), P.u_tmp + e * CONS, bu, bar);
        }
    }

    // calc_sources! (dg_3d.jl:1417-1437): evaluated into the (now dead) prim tile, natural node order
    if (have_src) {
        const Euler<3> eq(P.eq);
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            const int n = l16 + 16 * (r == 0 ? lm[0] : lm[1]);
            double un[5], x[3], sv[5];
#pragma unroll
            for (int v = 0; v < 5; ++v) un[v] = s_u[n * 5 + v];
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) x[dd] = P.node_coordinates[(e * 64 + n) * 3 + dd];
            eq.source_terms(P.source_terms, un, x, P.t, sv);
#pragma unroll
            for (int v = 0; v < 5; ++v) s_prim[n * 5 + v] = sv[v];
        }
        __syncwarp();
    }

    // 3. finish the two own nodes (i, j, k = lm[0], lm[1]) of the z line in registers; the z-pass
    // accumulators are rotated: slots (1, 2, 3) hold the (v3, v1, v2) momentum components
    {
        const int i = a0, j = a1;
        const double factor = WITH_SURFACE ? -P.inverse_jacobian[e] : 1.0;
        unsigned long long cfl0 = 0ull, cfl1 = 0ull, cfl2 = 0ull;
        double vals[2][5];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = lm[r];
            const int n = l16 + 16 * k;
            const double *t = s_du + pos[r] * 5;
            double(&val)[5] = vals[r];
            val[0] = t[0] + own[r][0];
            val[1] = t[1] + own[r][2];
            val[2] = t[2] + own[r][3];
            val[3] = t[3] + own[r][1];
            val[4] = t[4] + own[r][4];
            if constexpr (WITH_SURFACE) {
                // calc_surface_integral! (dg_3d.jl:1337-1394): directions 1..6 = -x,+x,-y,+y,-z,+z
                if (i == 0 || i == 3) {
                    const double *sf = s_sfv + ((i == 0 ? 0 : 1) * 16 + j + 4 * k) * 5;
                    const double w = i == 0 ? -P.inv_weight0 : P.inv_weight0;
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[v] = fma(sf[v], w, val[v]);
                }
                if (j == 0 || j == 3) {
                    const double *sf = s_sfv + ((j == 0 ? 2 : 3) * 16 + i + 4 * k) * 5;
                    const double w = j == 0 ? -P.inv_weight0 : P.inv_weight0;
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[v] = fma(sf[v], w, val[v]);
                }
                if (r == 0) {  // k = lm[0] is 0 (h = 0) or 3 (h = 1): always a z face; lm[1] never is
                    const double *sf = s_sfv + ((h ? 5 : 4) * 16 + l16) * 5;
                    const double w = h ? P.inv_weight0 : -P.inv_weight0;
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[v] = fma(sf[v], w, val[v]);
                }
                // apply_jacobian! (dg_3d.jl:1396-1414)
#pragma unroll
                for (int v = 0; v < 5; ++v) val[v] *= factor;
                if (have_src) {
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[v] += s_prim[n * 5 + v];
                }
            }
        }
        if (ut_late) {
            while (!mbar_try_wait(bar, 1)) {
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int n = l16 + 16 * lm[r];
            double(&val)[5] = vals[r];
            double *out_t = s_ut + n * 5;
            if (!rk) {
#pragma unroll
                for (int v = 0; v < 5; ++v) out_t[v] = val[v];
            } else {
                // 2N stage (methods_2N.jl:152-158): u_tmp = du - u_tmp * a; u += u_tmp * (b * dt)
                double *out_u = s_u + n * 5;
                double un[5];
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    const double tmp = need_ut ? val[v] - out_t[v] * P.rk_a : val[v];
                    out_t[v] = tmp;
                    un[v] = out_u[v] + tmp * P.rk_b_dt;
                    out_u[v] = un[v];
                }
                if (P.want_cfl) {
                    // max_dt of the updated state (stepsize_dg3d.jl:8-32); max_abs_speeds
                    // (compressible_euler_3d.jl:1770-1775) with the divisions done as one Newton reciprocal plus
                    // a residual correction each (within 1 ulp of k_max_dt's IEEE divisions)
                    const double rho = un[0], inv_rho = fast_rcp(rho);
                    double v1 = un[1] * inv_rho, v2 = un[2] * inv_rho, v3 = un[3] * inv_rho;
                    v1 = fma(fma(-rho, v1, un[1]), inv_rho, v1);
                    v2 = fma(fma(-rho, v2, un[2]), inv_rho, v2);
                    v3 = fma(fma(-rho, v3, un[3]), inv_rho, v3);
                    const double pr = (gamma - 1) * (un[4] - 0.5 * (un[1] * v1 + un[2] * v2 + un[3] * v3));
                    const double gp = gamma * pr;
                    double c2 = gp * inv_rho;
                    c2 = fma(fma(-rho, c2, gp), inv_rho, c2);
                    const double c = sqrt(c2);
                    const double lam[3] = {fabs(v1) + c, fabs(v2) + c, fabs(v3) + c};
                    cfl0 = max(cfl0, cfl_encode(lam[0]));
                    cfl1 = max(cfl1, cfl_encode(lam[1]));
                    cfl2 = max(cfl2, cfl_encode(lam[2]));
                }
            }
        }
        if (P.want_cfl) {
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
                atomicMax(P.cfl_key + (blockIdx.x & (kCflSlots - 1)), cfl_encode(P.inverse_jacobian[e] * sum));
            }
        }
    }
    // 4. results leave through the async proxy
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        constexpr uint32_t bu = CONS * sizeof(double);
        if (!rk) {
            tma_store(P.du + e * CONS, smem_u32(s_ut), bu);
        } else {
            tma_store(P.u_tmp + e * CONS, smem_u32(s_ut), bu);
            tma_store(P.u_out + e * CONS, smem_u32(s_u), bu);
        }
        tma_store_commit_and_wait_read();
    }
}

cudaError_t preload_tuned_euler3d_v7() {
    cudaError_t e = preload_kernel(k_element_euler3d_ranocha_p3_v7<true>);
    if (e != cudaSuccess) return e;
    return preload_kernel(k_element_euler3d_ranocha_p3_v7<false>);
}

cudaError_t launch_element_euler3d_ranocha_p3_v7(const KParams &P, bool with_surface, cudaStream_t s) {
    using C = TunedCfgV7;
    static PerDeviceFlag configured;
    if (!configured.test_and_set()) {
        cudaError_t err = cudaFuncSetAttribute(k_element_euler3d_ranocha_p3_v7<true>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout,
                                               cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_p3_v7<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
    }
    const unsigned blocks = (unsigned)((P.elem_end - P.elem_begin + C::EPB - 1) / C::EPB);
    // (the kernel derives `deferred` from the same condition)
    const bool deferred = !(with_surface && P.source_terms != TRIXI_B200_SRC_NONE);
    const size_t smem = deferred ? C::SMEM_DEFERRED : C::SMEM_RESIDENT;
    KParams Q = P;
    if (Q.prefetch_distance < 0) Q.prefetch_distance = C::blocks_per_sm(deferred) * Q.sm_count;
    if (with_surface)
        k_element_euler3d_ranocha_p3_v7<true><<<blocks, C::THREADS, smem, s>>>(Q);
    else
        k_element_euler3d_ranocha_p3_v7<false><<<blocks, C::THREADS, smem, s>>>(Q);
    return cudaSuccess;
}

}  // namespace tb
