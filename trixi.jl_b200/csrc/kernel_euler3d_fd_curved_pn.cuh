// Tuned element kernel for curved meshes at higher polynomial degrees (instantiated for polydeg 5 = the reference's
// own GPU benchmark, benchmark/CUDA/elixir_euler_taylor_green_vortex.jl:29-44): compressible Euler 3D,
// flux-differencing volume integral with flux_ranocha along averaged contravariant vectors
// (dgsem_structured/dg_3d.jl:94-175, dgsem_p4est/dg_3d_gpu.jl:30-136), fused with the surface integral, the nodal
// Jacobian and the 2N Runge-Kutta stage.
//
// The reference's KernelAbstractions kernel gives every node a thread that evaluates all of its 3 (N - 1) two-point
// fluxes itself, i.e. every flux twice.  Here a thread owns one whole LINE of N nodes per direction pass and
// evaluates its N (N - 1) / 2 node pairs once each (for N = 6: 15 independent flux evaluations per thread and pass,
// which is where the latency hiding comes from), accumulating N x 5 results in registers.  N^2 lines per element
// do not fill warps evenly, so a block takes EPB = 4 elements (144 of 160 threads busy) and synchronises with
// block barriers between the passes.
//  * node records (rho, v, 2 p, log rho, log rho - log p), the du tile and the Ja^d tile sit at the padded node
//    position i + N j + (N^2 + 1) k with odd record strides (on average 1.8-way bank conflicts in the passes: 36
//    lines against 16-lane half-warps leave no conflict-free linear layout);
//  * the contravariant vectors Ja^d of a pass are copied into their tile with cp.async while the previous pass
//    computes (the first version read them with __ldg at the start of each pass: long-scoreboard stalls of 3 warps
//    per issue); keeping all three in the records would halve the resident blocks;
//  * tiles take turns as in the p = 3 kernels: u arrives (TMA) in the du tile's storage, the face fluxes are fetched
//    into the record tile's storage while the z pass computes, u_tmp into the du tile's storage afterwards, and
//    u += b dt u_tmp leaves as a bulk reduce-add.
#pragma once
#include <cstdint>

#include "ranocha_common.cuh"

namespace tb {

template <int N, int EPB_>
struct CurvedNCfg {
    static constexpr int NN = N * N * N, NL = N * N, NF = N * N;
    static constexpr int EPB = EPB_, THREADS = ((EPB * NL + 31) / 32) * 32;
    static constexpr int SJ = N, SK = N * N + 1;                           // padded node position strides
    static constexpr int NPOS = (N - 1) * (1 + SJ + SK) + 1;               // positions per element
    static constexpr int PRIM = ((NPOS * kNP + 1) / 2) * 2, DU = ((NPOS * 5 + 1) / 2) * 2;  // doubles per element
    static constexpr int JA = ((NPOS * 3 + 1) / 2) * 2;
    static constexpr int CONS = NN * 5, SFV = 6 * NF * 5;
    static_assert(DU >= CONS && PRIM >= SFV, "epilogue tiles must fit the flux-pass tiles");
    // [prim tiles EPB x PRIM | du tiles EPB x DU | Ja^d tiles EPB x JA | 3 mbarriers | resident u tiles]
    static constexpr int OFF_DU = EPB * PRIM, OFF_JA = OFF_DU + EPB * DU, REGION = OFF_JA + EPB * JA;
    static constexpr size_t SMEM_STREAM = sizeof(double) * REGION + 32;
    static constexpr size_t SMEM_RESIDENT = SMEM_STREAM + sizeof(double) * EPB * CONS;
};

// (two resident blocks per SM: a 5-warp block puts two warps on one scheduler, whose register file holds three warps
// of 168 registers; a 4-warp block leaves each scheduler two warps of up to 255)
// GEN: also the 3S* / SSP stage updates (KParams::mode 2, 3); see kernel_euler3d_fd_p3.cuh
template <int N, int EPB_, bool WITH_SURFACE, bool GEN = false>
__global__ void __launch_bounds__(CurvedNCfg<N, EPB_>::THREADS, 2) k_element_euler3d_ranocha_curved_pn(const KParams P) {
    using C = CurvedNCfg<N, EPB_>;
    constexpr int NN = C::NN, NL = C::NL, NF = C::NF, EPB = C::EPB, CONS = C::CONS, SFV = C::SFV;
    constexpr int SJ = C::SJ, SK = C::SK, PRIM = C::PRIM, DU = C::DU, JA = C::JA;
    extern __shared__ __align__(128) double smem[];
    const bool resident = tuned_u_resident(P, WITH_SURFACE);
    double *s_prim = smem;               // [EPB][NPOS][7] padded positions
    double *s_du = smem + C::OFF_DU;     // [EPB][NPOS][5]; before the x pass: u, natural order [EPB][NN][5] (streamed)
    double *s_sfv = smem;                // epilogue: [EPB][6][NF][5] natural, one tile per element at stride PRIM
    double *s_ut = smem + C::OFF_DU;     // epilogue: u_tmp in / out, natural, one tile per element at stride DU
    double *s_ja = smem + C::OFF_JA;     // [EPB][NPOS][3]: contravariant vector Ja^d of the current pass
    const uint32_t bar_u = smem_u32(smem + C::REGION), bar_s = bar_u + 8, bar_t = bar_u + 16;
    double *s_u = smem + C::REGION + 4;  // resident only: [EPB][NN][5]

    const int tid = threadIdx.x;
    const long long e0 = (long long)EPB * blockIdx.x;
    const int nvalid = (int)min((long long)EPB, P.nelements - e0);
    const int le = tid / NL, l = tid - le * NL;  // local element and line of this thread
    const bool active = le < nvalid;             // (threads beyond EPB * NL and elements beyond the mesh idle)
    const long long e = e0 + le;
    const double gamma = P.eq.p[0], igm1 = P.eq.p[1];
    const bool rk = P.mode != 0;
    const bool need_ut = rk && P.rk_read_tmp;
    constexpr uint32_t bu1 = CONS * sizeof(double), bs1 = SFV * sizeof(double);

    if (tid == 0) {
        mbar_init(bar_u, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_t, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ustride = resident ? CONS : DU;  // element stride of the natural-order u tiles
    double *const s_uin = resident ? s_u : s_du;
    if (tid == 0) {
        mbar_expect_tx(bar_u, nvalid * bu1);
        for (int q = 0; q < nvalid; ++q) tma_load(smem_u32(s_uin + q * ustride), P.u + (e0 + q) * CONS, bu1, bar_u);
        if (WITH_SURFACE) tma_prefetch_l2(P.sfv + e0 * SFV, nvalid * bs1);
        if (need_ut) tma_prefetch_l2(P.u_tmp + e0 * CONS, nvalid * bu1);
    }
    // Ja^d of all nodes of the block into the Ja tile (8-byte cp.async: no registers, completion awaited later)
    auto stage_ja = [&](int d) {
        for (int idx = tid; idx < nvalid * NN; idx += C::THREADS) {
            const int q = idx / NN, n = idx - q * NN;
            const int i = n % N, j = (n / N) % N, k = n / (N * N);
            const double *src = P.contravariant_vectors + ((e0 + q) * NN + n) * 9 + 3 * d;
            const uint32_t dst = smem_u32(s_ja + q * JA + (i + SJ * j + SK * k) * 3);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8 * c), "l"(src + c) : "memory");
        }
    };
    stage_ja(0);
    while (!mbar_try_wait(bar_u, 0)) {
    }

    // 1. node records: thread-strided over the block's nodes
    for (int idx = tid; idx < nvalid * NN; idx += C::THREADS) {
        const int q = idx / NN, n = idx - q * NN;
        const int i = n % N, j = (n / N) % N, k = n / (N * N);
        const double *c = s_uin + q * ustride + n * 5;
        const double rho = c[0], m1 = c[1], m2 = c[2], m3 = c[3];
        const double inv_rho = fast_rcp(rho);
        double v1 = m1 * inv_rho, v2 = m2 * inv_rho, v3 = m3 * inv_rho;
        v1 = fma(fma(-rho, v1, m1), inv_rho, v1);
        v2 = fma(fma(-rho, v2, m2), inv_rho, v2);
        v3 = fma(fma(-rho, v3, m3), inv_rho, v3);
        const double pr = (gamma - 1) * (c[4] - 0.5 * (m1 * v1 + m2 * v2 + m3 * v3));
        const double lr = log_pos(rho);
        double *o = s_prim + q * PRIM + (i + SJ * j + SK * k) * kNP;
        o[0] = rho;
        o[1] = v1;
        o[2] = v2;
        o[3] = v3;
        o[4] = pr + pr;
        o[5] = lr;
        o[6] = lr - log_pos(pr);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // (also: u has been read before the x pass overwrites the du tiles; Ja^1 has landed)

    // 2. direction passes; line l = a0 + N a1: x: (j, k) = (a0, a1), y: (i, k), z: (i, j)
    const int a0 = l % N, a1 = l / N;
    double *const sp = s_prim + le * PRIM, *const sd = s_du + le * DU;
    const double *const sja = s_ja + le * JA;
    double acc[N][5];
    int pbase = 0, pstep = 1;
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        // node m of the line: padded position pbase + m pstep, natural node nbase + m nstep
        pbase = d == 0 ? SJ * a0 + SK * a1 : (d == 1 ? a0 + SK * a1 : a0 + SJ * a1);
        pstep = d == 0 ? 1 : (d == 1 ? SJ : SK);
        double q[N][10];
        if (active) {
#pragma unroll
            for (int m = 0; m < N; ++m) {
                const double *src = sp + (pbase + m * pstep) * kNP;
#pragma unroll
                for (int c = 0; c < kNP; ++c) q[m][c] = src[c];
                const double *ja = sja + (pbase + m * pstep) * 3;
                q[m][7] = ja[0];
                q[m][8] = ja[1];
                q[m][9] = ja[2];
            }
        }
        if (d < 2) {
            __syncthreads();  // every thread has read Ja^d: the next pass's vectors arrive while this one computes
            stage_ja(d + 1);
        } else if (WITH_SURFACE) {
            // every thread has read the record tiles for the last time: the face fluxes land in their storage
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(bar_s, nvalid * bs1);
                for (int qq = 0; qq < nvalid; ++qq)
                    tma_load(smem_u32(s_sfv + qq * PRIM), P.sfv + (e0 + qq) * SFV, bs1, bar_s);
            }
        }
        if (active) {
            // N (N - 1) / 2 symmetric pairs with the plain D_split weights (constant-bank operands); the powers of
            // two ranocha_pair_normal leaves out (4 for the density flux, 8 for the others) are applied once in
            // step 3 -- exact, so the sums are the same bits as with pre-scaled weights
#pragma unroll
            for (int m = 0; m < N; ++m)
#pragma unroll
                for (int v = 0; v < 5; ++v) acc[m][v] = 0.0;
#pragma unroll
            for (int a = 0; a < N; ++a)
#pragma unroll
                for (int b = a + 1; b < N; ++b) {
                    double g[5];
                    ranocha_pair_normal(q[a], q[b], 0, igm1, g);
#pragma unroll
                    for (int v = 0; v < 5; ++v) {
                        acc[a][v] = fma(P.dsplit_c[a + N * b], g[v], acc[a][v]);
                        acc[b][v] = fma(P.dsplit_c[b + N * a], g[v], acc[b][v]);
                    }
                }
            if (d == 0) {
#pragma unroll
                for (int m = 0; m < N; ++m)
#pragma unroll
                    for (int v = 0; v < 5; ++v) sd[(pbase + m * pstep) * 5 + v] = acc[m][v];
            } else if (d == 1) {
#pragma unroll
                for (int m = 0; m < N; ++m)
#pragma unroll
                    for (int v = 0; v < 5; ++v) sd[(pbase + m * pstep) * 5 + v] += acc[m][v];
            }
        }
        if (d < 2) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
        }
    }

    // 3. finish the z line (i, j) = (a0, a1), nodes n = l + NL k
    double val[N][5];
    if (active) {
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int v = 0; v < 5; ++v)
                val[k][v] = (sd[(pbase + k * pstep) * 5 + v] + acc[k][v]) * (v == 0 ? 0.25 : 0.125);
    }
    fence_proxy_async();
    __syncthreads();  // the du tiles are dead: u_tmp takes their place
    if (need_ut && tid == 0) {
        mbar_expect_tx(bar_t, nvalid * bu1);
        for (int qq = 0; qq < nvalid; ++qq) tma_load(smem_u32(s_ut + qq * DU), P.u_tmp + (e0 + qq) * CONS, bu1, bar_t);
    }
    if (WITH_SURFACE) {
        while (!mbar_try_wait(bar_s, 0)) {
        }
    }
    if (active) {
        const int i = a0, j = a1;
        if constexpr (WITH_SURFACE) {
            // calc_surface_integral!: StructuredMesh "-" on the negative faces, P4estMesh "+" on all six
            const double wneg = P.p4est ? P.inv_weight0 : -P.inv_weight0;
            const double *ssf = s_sfv + le * PRIM;
            if (i == 0 || i == N - 1) {
                const double *sf = ssf + ((i == 0 ? 0 : 1) * NF + j) * 5;
                const double w = i == 0 ? wneg : P.inv_weight0;
#pragma unroll
                for (int k = 0; k < N; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[k][v] = fma(sf[N * 5 * k + v], w, val[k][v]);
            }
            if (j == 0 || j == N - 1) {
                const double *sf = ssf + ((j == 0 ? 2 : 3) * NF + i) * 5;
                const double w = j == 0 ? wneg : P.inv_weight0;
#pragma unroll
                for (int k = 0; k < N; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[k][v] = fma(sf[N * 5 * k + v], w, val[k][v]);
            }
            {
                const double *sf = ssf + (4 * NF + l) * 5;
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    val[0][v] = fma(sf[v], wneg, val[0][v]);
                    val[N - 1][v] = fma(sf[NF * 5 + v], P.inv_weight0, val[N - 1][v]);
                }
            }
            // apply_jacobian! with the nodal inverse Jacobian (dgsem_structured/dg_3d.jl:937-956)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double factor = -P.inverse_jacobian[e * NN + l + NL * k];
#pragma unroll
                for (int v = 0; v < 5; ++v) val[k][v] *= factor;
            }
        }
    }
    if (need_ut) {
        while (!mbar_try_wait(bar_t, 0)) {
        }
    }
    double *const sut = s_ut + le * DU;
    const bool rk2n = !GEN || P.mode == 1;  // (modes 2 and 3, the 3S* and SSP stages, always run resident)
    if (active) {
        if (rk2n && need_ut) {
#pragma unroll
            for (int k = 0; k < N; ++k)
#pragma unroll
                for (int v = 0; v < 5; ++v) val[k][v] -= sut[(l + NL * k) * 5 + v] * P.rk_a;
        }
        if (!rk || rk2n) {
#pragma unroll
            for (int k = 0; k < N; ++k)
#pragma unroll
                for (int v = 0; v < 5; ++v) sut[(l + NL * k) * 5 + v] = val[k][v];  // du (mode 0) or the new u_tmp
        }
    }
    if (rk) {
        if (!resident) {
            __syncthreads();  // every thread is done with the face tiles: b dt u_tmp takes their place
            if (active) {
                double *const sinc = s_sfv + le * PRIM;
#pragma unroll
                for (int k = 0; k < N; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) sinc[(l + NL * k) * 5 + v] = __dmul_rn(val[k][v], P.rk_b_dt);
            }
        } else if (active) {
            double *const suo = s_u + le * CONS;
#pragma unroll
            for (int k = 0; k < N; ++k)
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    double *out_u = suo + (l + NL * k) * 5 + v;
                    if (rk2n) {
                        *out_u = __dadd_rn(*out_u, __dmul_rn(val[k][v], P.rk_b_dt));
                    } else {  // 3S* / SSP stage (KParams::mode 2, 3): u_tmp2 straight from global memory
                        double *out_t = sut + (l + NL * k) * 5 + v;
                        double xn;
                        *out_u = rk_stage_3s_ssp(P, val[k][v], need_ut ? *out_t : 0.0, *out_u,
                                                 P.mode == 2 ? P.u_tmp2[e * CONS + (l + NL * k) * 5 + v] : 0.0, xn);
                        *out_t = xn;
                    }
                }
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        for (int qq = 0; qq < nvalid; ++qq) {
            if (!rk) {
                tma_store(P.du + (e0 + qq) * CONS, smem_u32(s_ut + qq * DU), bu1);
            } else {
                if (!GEN || P.rk_write_tmp) tma_store(P.u_tmp + (e0 + qq) * CONS, smem_u32(s_ut + qq * DU), bu1);
                if (resident)
                    tma_store(P.u_out + (e0 + qq) * CONS, smem_u32(s_u + qq * CONS), bu1);
                else
                    tma_reduce_add_f64(P.u_out + (e0 + qq) * CONS, smem_u32(s_sfv + qq * PRIM), bu1);
            }
        }
        tma_store_commit_and_wait_read();
    }
}

template <int N, int EPB>
cudaError_t preload_tuned_euler3d_curved_pn() {
    cudaError_t e = preload_kernel(k_element_euler3d_ranocha_curved_pn<N, EPB, true>);
    if (e != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_ranocha_curved_pn<N, EPB, true, true>)) != cudaSuccess) return e;
    return preload_kernel(k_element_euler3d_ranocha_curved_pn<N, EPB, false>);
}

template <int N, int EPB>
cudaError_t launch_element_euler3d_ranocha_curved_pn(const KParams &P, bool with_surface, cudaStream_t s) {
    using C = CurvedNCfg<N, EPB>;
    static PerDeviceFlag configured;
    if (!configured.test_and_set()) {
        cudaError_t err = cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_pn<N, EPB, true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_RESIDENT);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_pn<N, EPB, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_RESIDENT);
        if (err != cudaSuccess) return err;
        cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_pn<N, EPB, true>,
                             cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_pn<N, EPB, true, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_RESIDENT);
        if (err != cudaSuccess) return err;
        cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_pn<N, EPB, true, true>,
                             cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_element_euler3d_ranocha_curved_pn<N, EPB, false>,
                             cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    const unsigned blocks = (unsigned)((P.nelements + C::EPB - 1) / C::EPB);
    KParams Q = P;
    Q.want_cfl = 0;  // (k_max_dt_curved reduces the CFL speeds of curved meshes)
    const bool resident = tuned_u_resident(Q, with_surface);
    const size_t smem = resident ? C::SMEM_RESIDENT : C::SMEM_STREAM;
    if (Q.mode > 1)  // 3S* / SSP stage (always with the surface terms)
        k_element_euler3d_ranocha_curved_pn<N, EPB, true, true><<<blocks, C::THREADS, smem, s>>>(Q);
    else if (with_surface)
        k_element_euler3d_ranocha_curved_pn<N, EPB, true><<<blocks, C::THREADS, smem, s>>>(Q);
    else
        k_element_euler3d_ranocha_curved_pn<N, EPB, false><<<blocks, C::THREADS, smem, s>>>(Q);
    return cudaSuccess;
}

}  // namespace tb
