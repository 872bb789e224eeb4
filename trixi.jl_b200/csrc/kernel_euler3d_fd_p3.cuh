// Tuned element kernel: 3D compressible Euler, polydeg 3 (4^3 nodes), flux-differencing volume integral
// with flux_ranocha, fused with surface integral, Jacobian, source terms and the 2N Runge-Kutta stage.
// This is the headline configuration (BASELINE.json: 3D Euler EC p=3).
//
// Work decomposition, second generation.  The first generation (one warp per element, two
// threads per line, kernel_euler3d_fd_p3_v7.cuh) was issue-bound: ncu counted 1136 thread instructions per DOF
// of which 41% were FP64 -- the rest was shared-memory traffic for operands that two threads of a line both
// needed, read-modify-write of the du tile by both of them, run-time line roles, and libm logarithms.
//  * One warp = one CTA = TWO elements, 16 threads per element.  In every direction pass a thread owns one
//    whole line of 4 nodes: it loads the 4 node records once (28 LDS for 6 two-point fluxes instead of 28 for 3),
//    evaluates all 6 symmetric pairs with compile-time roles (D_split entries are constant-bank operands of the
//    DFMAs, no weight registers, no selects), and accumulates the 4 x 5 results in registers.  Every two-point
//    flux is evaluated exactly once (288 per element, like the reference's symmetric loop dg_3d.jl:177-211);
//    nobody writes into another thread's nodes, so the du tile sees one store (x), one read-modify-write (y)
//    and one read (z) per node instead of six accesses.
//  * 8 CTAs = 16 elements resident per SM (shared-memory-limited, as before); each thread carries six
//    independent flux evaluations per pass, which is where the latency hiding comes from now.
//  * Tile I/O is TMA: `cp.async.bulk` loads of the two elements' contiguous u / surface_flux_values / u_tmp
//    records on mbarriers (u is awaited at once, the surface fluxes only before the epilogue), results leave
//    through `cp.async.bulk` stores.
//  * flux_ranocha is evaluated in the hoisted form of the reference's own SIMD kernel
//    (dg_3d_compressible_euler.jl:289-309,360-385): primitive variables, log(rho) and log(rho) - log(p) once
//    per node.  The logarithm is an fdlibm-style kernel for positive normal arguments (libm's takes twice the
//    instructions for its special cases); the factors 1/2 of the arithmetic means are folded into the
//    D_split weights (powers of two: bit-identical results).
//  * The prim and du tiles are AoS records at the swizzled node position pos(i,j,k) = 16k + 4(j^k) + (i^k):
//    in every direction pass the 16 lines of an element hit 16 distinct 8-byte banks (record strides 5 and 7
//    are odd); the TMA-filled buffers keep the global (natural) order, conflict-free for the per-node passes.
#pragma once
#include <cstdint>

#include "ranocha_common.cuh"

namespace tb {

struct TunedCfg {
    static constexpr int EPB = 2, THREADS = 32;  // one warp, two elements, one thread per line
    static constexpr int CONS = 320, PRIM = 64 * kNP, SFV = 480;  // doubles per element
    // Shared memory is what limits the resident warps, so the tiles take turns in one region per CTA (doubles):
    //   flux passes:  [prim tiles 0..896 | pad | du tiles 1004..1644]  (u arrives in the du tiles' storage and is
    //                                           consumed by the primitive-variable pass before the x pass)
    //   epilogue:     [surface_flux_values 0..1004 | u_tmp 1004..1644] (the faces are fetched while the z pass
    //                                           computes -- it has read the prim tile by then -- and u_tmp once
    //                                           the du tile has been read; b dt u_tmp finally goes to 0..640)
    // 6.4 KB per element: 16 CTAs = 32 elements per SM.  The updated u leaves as a bulk reduce-add of b dt u_tmp onto
    // u in L2 (cp.reduce.async.bulk .add.f64), so u is not needed in the epilogue at all.  Launches that do need it
    // there (source terms, the CFL reduction of the last stage, out-of-place updates) keep a resident u tile:
    // 8.9 KB per element, 12 CTAs per SM.
    // Face tiles: element h at h * 504; with single-copy face fluxes every face arrives by its own 640-byte bulk
    // copy anyway and sits at q * 84, so that in the surface integral the lanes of the two faces of a direction
    // (and of the two elements) hit different banks (offsets 0, 4, 8, 12 against a lane stride of 5:
    // conflict-free); with two-copy face fluxes an element's six faces arrive as one block (stride 80: the two
    // faces of a direction share banks).
    static constexpr int OFF_DU = 1004, REGION = OFF_DU + EPB * CONS, SFV_H = 504, SFV_PAD = 84;
    static_assert(SFV_H + 5 * SFV_PAD + 80 <= OFF_DU, "face tiles must stay clear of the du tile");
    static constexpr size_t SMEM_STREAM = sizeof(double) * REGION + 64;  // + 3 mbarriers + 6 neighbour ids
    static constexpr size_t SMEM_RESIDENT = SMEM_STREAM + sizeof(double) * EPB * CONS;
    static constexpr int MIN_BLOCKS = 16;
    static constexpr int blocks_per_sm(bool resident) { return resident ? 12 : 16; }
};

// GEN: the instantiation that also knows the 3S* and SSP stage updates (KParams::mode 2, 3); write-du and 2N launches
// use GEN = false, whose code is exactly the 2N kernel (the extra epilogue code costs 2% when it is merely present)
// LEAN: the instantiation for the launches that make up four of the five stages of a 2N step on a TreeMesh: 2N stage,
// streamed u (reduce-add), single-copy face fluxes, no source terms, no L2 hints.  The run-time branches for everything
// else fold away (2920 -> fewer SASS instructions; the kernel is sensitive to its code size).
template <bool WITH_SURFACE, bool GEN = false, bool LEAN = false>
__global__ void __launch_bounds__(TunedCfg::THREADS, TunedCfg::MIN_BLOCKS)
    k_element_euler3d_ranocha_p3(const KParams P) {
    using C = TunedCfg;
    constexpr int CONS = C::CONS, PRIM = C::PRIM, SFV = C::SFV, EPB = C::EPB;
    extern __shared__ __align__(128) double smem[];
    static_assert(!LEAN || (WITH_SURFACE && !GEN), "the lean instantiation is a 2N stage with surface terms");
    const bool have_src = LEAN ? false : WITH_SURFACE && P.source_terms != TRIXI_B200_SRC_NONE;
    const bool resident = LEAN ? false : tuned_u_resident(P, WITH_SURFACE);
    const bool sfv_single = LEAN ? true : P.sfv_single != 0;
    double *s_prim = smem;                 // [2][64][7] swizzled
    double *s_du = smem + C::OFF_DU;       // [2][64][5] swizzled; before the x pass: u, natural order (not resident)
    double *s_sfv = smem;                  // epilogue: six [16][5] faces per element at stride fs, element h at h * 504
    double *s_inc = smem;                  // then b dt u_tmp [2][64][5] natural (not resident)
    double *s_ut = smem + C::OFF_DU;       // epilogue: [2][64][5] natural: u_tmp in, u_tmp (or du) out
    const uint32_t bar_u = smem_u32(smem + C::REGION), bar_s = bar_u + 8, bar_t = bar_u + 16;
    int *s_nb = reinterpret_cast<int *>(smem + C::REGION + 4);  // [2][3] left neighbours (single-copy face fluxes)
    double *s_u = smem + C::REGION + 8;    // resident only: [2][64][5] natural: u in, updated u out

    const int lane = threadIdx.x;
    const int t = lane & 15;
    const long long e0 = P.elem_begin + (long long)EPB * blockIdx.x;
    const int nvalid = (int)min((long long)EPB, P.elem_end - e0);
    // on an odd tail the second half-warp mirrors the first (same tiles, same values; nothing extra is stored)
    const int eh = (lane >> 4) < nvalid ? (lane >> 4) : 0;
    const long long e = e0 + eh;
    const double gamma = P.eq.p[0], igm1 = P.eq.p[1];
    const bool rk = LEAN ? true : P.mode != 0;
    const bool need_ut = rk && P.rk_read_tmp;
    const uint32_t bu = nvalid * CONS * sizeof(double), bs = nvalid * SFV * sizeof(double);

    // L2 priorities: u is touched again by this CTA's reduce-add ~10 us later (evict_last); everything else streams
    const bool hints = LEAN ? false : P.l2_hints != 0;
    const uint64_t pol_keep = hints ? l2_policy_evict_last() : 0ull, pol_stream = hints ? l2_policy_evict_first() : 0ull;
    // 0. TMA load of the two contiguous u records; the records the epilogue will want are pulled into L2 meanwhile
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
        if (hints && !resident)
            tma_load_hint(smem_u32(s_uin), P.u + e0 * CONS, bu, bar_u, pol_keep);
        else
            tma_load(smem_u32(s_uin), P.u + e0 * CONS, bu, bar_u);
        // (single-copy face fluxes: half of the own block is never read and the rest is fetched a whole z pass before
        // it is needed, so nothing is prefetched)
        if (WITH_SURFACE && !sfv_single) tma_prefetch_l2(P.sfv + e0 * SFV, bs);
        if (need_ut) tma_prefetch_l2(P.u_tmp + e0 * CONS, bu);
        // warm L2 for the elements that will occupy this CTA slot next (blocks are scheduled in index order:
        // one wave further on), so their u tile sees L2 instead of HBM latency
        const long long en = e0 + P.prefetch_distance;
        if (P.prefetch_distance > 0 && en + EPB <= P.nelements)
            tma_prefetch_l2(P.u + en * CONS, EPB * CONS * sizeof(double));
    }
    if (WITH_SURFACE && sfv_single && lane < 3 * nvalid) s_nb[lane] = P.minus_nb[e0 * 3 + lane];
    const double *const su = s_uin + eh * CONS;
    double *const sp = s_prim + eh * PRIM, *const sd = s_du + eh * CONS;
    while (!mbar_try_wait(bar_u, 0)) {
    }

    // 1. cons2prim + logs for nodes t, t + 16, t + 32, t + 48 (the z line this thread also finishes in step 3);
    //    a half-warp reads 16 consecutive node records: conflict-free
#pragma unroll 2
    for (int r = 0; r < 4; ++r) {
        const double *c = su + (t + 16 * r) * 5;
        const double rho = c[0], m1 = c[1], m2 = c[2], m3 = c[3];
        const double inv_rho = fast_rcp(rho);
        // v = rho_v / rho with a residual correction (cons2prim, compressible_euler_3d.jl:1783-1793)
        double v1 = m1 * inv_rho, v2 = m2 * inv_rho, v3 = m3 * inv_rho;
        v1 = fma(fma(-rho, v1, m1), inv_rho, v1);
        v2 = fma(fma(-rho, v2, m2), inv_rho, v2);
        v3 = fma(fma(-rho, v3, m3), inv_rho, v3);
        const double pr = (gamma - 1) * (c[4] - 0.5 * (m1 * v1 + m2 * v2 + m3 * v3));
        const double lr = log_pos(rho);
        double *o = sp + (16 * r + (t ^ (5 * r))) * kNP;  // swz_pos(t + 16 r)
        o[0] = rho;
        o[1] = v1;
        o[2] = v2;
        o[3] = v3;
        o[4] = pr + pr;
        o[5] = lr;
        o[6] = lr - log_pos(pr);
    }
    __syncwarp();  // (also: every thread has read u before the x pass overwrites the du tile)

    // 2. direction passes x, y, z with ONE copy of the flux code; the direction only enters through
    // shared-memory offsets (velocity slots rotated while loading, momentum slots while storing).
    // The line of this thread: x: (j, k) = (t & 3, t >> 2), y: (i, k) = (t & 3, t >> 2), z: (i, j) likewise;
    // its node m sits at swizzled position B ^ (m * M) with M = 1, 4, 21.
    const int a0 = t & 3, a1 = t >> 2;
    const int B0 = 16 * a1 + 4 * (a0 ^ a1) + a1, B1 = 20 * a1 + (a0 ^ a1);
    double acc[4][5];
    int pos[4];
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        const int B = d == 0 ? B0 : (d == 1 ? B1 : t), M = d == 0 ? 1 : (d == 1 ? 4 : 21);
        const int on = 1 + d, ot1 = d == 2 ? 1 : 2 + d, ot2 = d == 0 ? 3 : d;  // 1 + (d + {0,1,2}) % 3
        double q[4][kNP];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            pos[m] = B ^ (m * M);
            const double *src = sp + pos[m] * kNP;
            q[m][0] = src[0];
            q[m][1] = src[on];
            q[m][2] = src[ot1];
            q[m][3] = src[ot2];
            q[m][4] = src[4];
            q[m][5] = src[5];
            q[m][6] = src[6];
        }
        if (WITH_SURFACE && d == 2) {
            // the prim tile has been read for the last time: the surface fluxes land in its storage while the z
            // pass computes
            fence_proxy_async();
            __syncwarp();
            constexpr uint32_t bs1 = SFV * sizeof(double), bf = 80 * sizeof(double);
            if (lane == 0) mbar_expect_tx(bar_s, bs);
            if (!sfv_single) {
                if (lane == 0) {
                    tma_load(smem_u32(s_sfv), P.sfv + e0 * SFV, bs1, bar_s);
                    if (nvalid == EPB) tma_load(smem_u32(s_sfv + C::SFV_H), P.sfv + (e0 + 1) * SFV, bs1, bar_s);
                }
            } else if (lane < 6 * nvalid) {
                // one 640-byte copy per face, addresses computed by twelve lanes in parallel: + faces from the
                // element's own block; - faces from the left neighbour's + face, or from the own block where the
                // face is a boundary, a mortar or shared with another rank
                const int q = lane >= 6 ? 1 : 0, f = lane - 6 * q, o = f >> 1;
                const int nb = (f & 1) ? -1 : s_nb[3 * q + o];
                const double *src = nb >= 0 ? P.sfv + (long long)nb * SFV + (2 * o + 1) * 80 : P.sfv + (e0 + q) * SFV + f * 80;
                const uint32_t dst = smem_u32(s_sfv + q * C::SFV_H + f * C::SFV_PAD);
                if (hints)
                    tma_load_hint(dst, src, bf, bar_s, pol_stream);
                else
                    tma_load(dst, src, bf, bar_s);
            }
        }
        // du[a] += D_split[a, b] f(a, b), du[b] += D_split[b, a] f(a, b) for the 6 pairs a < b of the line;
        // dsplit_h = D_split / 2 and dsplit_q = D_split / 4 undo the scaling of g
        double g[5];
#define TB_PAIR(a, b, FIRST_A, FIRST_B)                                                                        \
    ranocha_pair_rot(q[a], q[b], igm1, g);                                                                   \
    acc[a][0] = FIRST_A ? P.dsplit_h[a + 4 * b] * g[0] : fma(P.dsplit_h[a + 4 * b], g[0], acc[a][0]);          \
    acc[b][0] = FIRST_B ? P.dsplit_h[b + 4 * a] * g[0] : fma(P.dsplit_h[b + 4 * a], g[0], acc[b][0]);          \
    _Pragma("unroll") for (int v = 1; v < 5; ++v) {                                                            \
        acc[a][v] = FIRST_A ? P.dsplit_q[a + 4 * b] * g[v] : fma(P.dsplit_q[a + 4 * b], g[v], acc[a][v]);      \
        acc[b][v] = FIRST_B ? P.dsplit_q[b + 4 * a] * g[v] : fma(P.dsplit_q[b + 4 * a], g[v], acc[b][v]);      \
    }
        TB_PAIR(0, 1, true, true)
        TB_PAIR(2, 3, true, true)
        TB_PAIR(0, 2, false, false)
        TB_PAIR(1, 3, false, false)
        TB_PAIR(0, 3, false, false)
        TB_PAIR(1, 2, false, false)
#undef TB_PAIR
        // the x and y sums meet in the du tile; the z sums stay in registers for step 3
        if (d == 0) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                double *o = sd + pos[m] * 5;
                o[0] = acc[m][0];
                o[1] = acc[m][1];
                o[2] = acc[m][2];
                o[3] = acc[m][3];
                o[4] = acc[m][4];
            }
            __syncwarp();
        } else if (d == 1) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                double *o = sd + pos[m] * 5;  // accumulators (f_rho, f_y, f_z, f_x, f_E)
                o[0] += acc[m][0];
                o[2] += acc[m][1];
                o[3] += acc[m][2];
                o[1] += acc[m][3];
                o[4] += acc[m][4];
            }
            __syncwarp();
        }
    }

    // 3. finish the z line (i, j) = (a0, a1), nodes n = t + 16 k, in registers; the z-pass accumulators are
    // rotated: slots (1, 2, 3) hold the (v3, v1, v2) momentum components
    double val[4][5];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double *o = sd + pos[k] * 5;
        val[k][0] = o[0] + acc[k][0];
        val[k][1] = o[1] + acc[k][2];
        val[k][2] = o[2] + acc[k][3];
        val[k][3] = o[3] + acc[k][1];
        val[k][4] = o[4] + acc[k][4];
    }
    // the du tile is dead now: u_tmp takes its place
    fence_proxy_async();
    __syncwarp();
    if (need_ut && lane == 0) {
        mbar_expect_tx(bar_t, bu);
        if (hints)
            tma_load_hint(smem_u32(s_ut), P.u_tmp + e0 * CONS, bu, bar_t, pol_stream);
        else
            tma_load(smem_u32(s_ut), P.u_tmp + e0 * CONS, bu, bar_t);
    }
    if (WITH_SURFACE) {
        while (!mbar_try_wait(bar_s, 0)) {
        }
    }
    {
        const int i = a0, j = a1;
        if constexpr (WITH_SURFACE) {
            // calc_surface_integral! (dg_3d.jl:1337-1394): directions 1..6 = -x,+x,-y,+y,-z,+z
            const double *ssf = s_sfv + eh * C::SFV_H;
            const int fs = sfv_single ? C::SFV_PAD : 80;  // face stride in the tile
            if (i == 0 || i == 3) {
                const double *sf = ssf + (i == 0 ? 0 : fs) + j * 5;
                const double w = i == 0 ? -P.inv_weight0 : P.inv_weight0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[k][v] = fma(sf[20 * k + v], w, val[k][v]);
            }
            if (j == 0 || j == 3) {
                const double *sf = ssf + (j == 0 ? 2 * fs : 3 * fs) + i * 5;
                const double w = j == 0 ? -P.inv_weight0 : P.inv_weight0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) val[k][v] = fma(sf[20 * k + v], w, val[k][v]);
            }
            {
                const double *sf = ssf + 4 * fs + t * 5;
#pragma unroll
                for (int v = 0; v < 5; ++v) {
                    val[0][v] = fma(sf[v], -P.inv_weight0, val[0][v]);
                    val[3][v] = fma(sf[fs + v], P.inv_weight0, val[3][v]);
                }
            }
            // apply_jacobian! (dg_3d.jl:1396-1414)
            const double factor = -P.inverse_jacobian[e];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int v = 0; v < 5; ++v) val[k][v] *= factor;
            // calc_sources! (dg_3d.jl:1417-1437)
            if (have_src) {
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
                        // (k is a run-time index here: select instead of indexing the register array)
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
            // 2N stage (methods_2N.jl:152-158): u_tmp = du - u_tmp * a; u += u_tmp * (b * dt).  The product is
            // rounded before it is added, here and in the bulk reduce-add, so both forms give the same bits.
            const bool rk2n = !GEN || P.mode == 1;  // (modes 2 and 3, the 3S* and SSP stages, always run resident)
            if (need_ut && rk2n) {  // (warp-uniform)
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
                __syncwarp();  // every thread is done with surface_flux_values: b dt u_tmp takes its place
                double *const sinc = s_inc + eh * CONS;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < 5; ++v) sinc[(t + 16 * k) * 5 + v] = __dmul_rn(val[k][v], P.rk_b_dt);
            } else {
                double *const suo = s_u + eh * CONS;
                unsigned long long cfl0 = 0ull, cfl1 = 0ull, cfl2 = 0ull;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double *out_u = suo + (t + 16 * k) * 5;
                    double un[5];
                    if (rk2n) {
#pragma unroll
                        for (int v = 0; v < 5; ++v) {
                            un[v] = __dadd_rn(out_u[v], __dmul_rn(val[k][v], P.rk_b_dt));
                            out_u[v] = un[v];
                        }
                    } else {
                        // 3S* / SSP stage (KParams::mode 2, 3): u_tmp2 comes straight from global memory (40-byte
                        // node records, consecutive lanes = consecutive records)
                        const double *u2 = P.u_tmp2 + e * CONS + (t + 16 * k) * 5;
                        double *out_t = sut + (t + 16 * k) * 5;
#pragma unroll
                        for (int v = 0; v < 5; ++v) {
                            double xn;
                            un[v] = rk_stage_3s_ssp(P, val[k][v], need_ut ? out_t[v] : 0.0, out_u[v],
                                                    P.mode == 2 ? u2[v] : 0.0, xn);
                            out_t[v] = xn;
                            out_u[v] = un[v];
                        }
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
                        cfl0 = max(cfl0, cfl_encode(fabs(v1) + c));
                        cfl1 = max(cfl1, cfl_encode(fabs(v2) + c));
                        cfl2 = max(cfl2, cfl_encode(fabs(v3) + c));
                    }
                }
                if (P.want_cfl) {
#pragma unroll
                    for (int off = 8; off > 0; off >>= 1) {  // per element = per half-warp
                        cfl0 = max(cfl0, __shfl_xor_sync(0xffffffffu, cfl0, off));
                        cfl1 = max(cfl1, __shfl_xor_sync(0xffffffffu, cfl1, off));
                        cfl2 = max(cfl2, __shfl_xor_sync(0xffffffffu, cfl2, off));
                    }
                    if (t == 0 && (lane >> 4) < nvalid) {
                        double sum = 0.0;
                        sum += __longlong_as_double((long long)cfl0);
                        sum += __longlong_as_double((long long)cfl1);
                        sum += __longlong_as_double((long long)cfl2);
                        atomicMax(P.cfl_key + ((blockIdx.x * 2 + (lane >> 4)) & (kCflSlots - 1)),
                                  cfl_encode(P.inverse_jacobian[e] * sum));
                    }
                }
            }
        }
    }
    // 4. results leave through the async proxy
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (!rk) {
            tma_store(P.du + e0 * CONS, smem_u32(s_ut), bu);
        } else if (hints && !resident) {
            tma_store_hint(P.u_tmp + e0 * CONS, smem_u32(s_ut), bu, pol_stream);
            tma_reduce_add_f64_hint(P.u_out + e0 * CONS, smem_u32(s_inc), bu, pol_stream);
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

cudaError_t preload_tuned_euler3d() {
    cudaError_t e = preload_kernel(k_element_euler3d_ranocha_p3<true>);
    if (e != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_ranocha_p3<true, true>)) != cudaSuccess) return e;
    if ((e = preload_kernel(k_element_euler3d_ranocha_p3<true, false, true>)) != cudaSuccess) return e;
    return preload_kernel(k_element_euler3d_ranocha_p3<false>);
}

cudaError_t launch_element_euler3d_ranocha_p3(const KParams &P, bool with_surface, cudaStream_t s) {
    using C = TunedCfg;
    static PerDeviceFlag configured;
    if (!configured.test_and_set()) {
        cudaError_t err = cudaFuncSetAttribute(k_element_euler3d_ranocha_p3<true>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout,
                                               cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_p3<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_p3<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_element_euler3d_ranocha_p3<true, false, true>,
                                   cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
    }
    const unsigned blocks = (unsigned)((P.elem_end - P.elem_begin + C::EPB - 1) / C::EPB);
    const bool resident = tuned_u_resident(P, with_surface);  // (the kernel evaluates the same condition)
    const size_t smem = resident ? C::SMEM_RESIDENT : C::SMEM_STREAM;
    KParams Q = P;
    if (Q.prefetch_distance < 0) Q.prefetch_distance = C::EPB * C::blocks_per_sm(resident) * Q.sm_count;
    if (Q.mode > 1)  // 3S* / SSP stage (always with the surface terms)
        k_element_euler3d_ranocha_p3<true, true><<<blocks, C::THREADS, smem, s>>>(Q);
    else if (with_surface && Q.mode == 1 && !resident && Q.sfv_single && !Q.l2_hints)
        k_element_euler3d_ranocha_p3<true, false, true><<<blocks, C::THREADS, smem, s>>>(Q);
    else if (with_surface)
        k_element_euler3d_ranocha_p3<true><<<blocks, C::THREADS, smem, s>>>(Q);
    else
        k_element_euler3d_ranocha_p3<false><<<blocks, C::THREADS, smem, s>>>(Q);
    return cudaSuccess;
}

}  // namespace tb
