// Host-side launchers for the templated kernels and the per-equation dispatch tables.
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace tb {

struct Launchers {
    void (*interface_flux)(const KParams &, cudaStream_t);
    void (*boundary_flux)(const KParams &, cudaStream_t);
    void (*mortar_flux)(const KParams &, cudaStream_t);
    void (*error_norms)(const KParams &, const NormParams &, cudaStream_t);
    void (*integrate)(const KParams &, int quantity, double *sums, cudaStream_t);
    // with_surface = false: volume terms only (stage-level parity entry point)
    cudaError_t (*element)(const KParams &, bool with_surface, cudaStream_t);
    // IndicatorHennemannGassner blending factors of P.u into P.alpha (VolumeIntegralShockCapturingHG):
    // stage 1 = per-element values (alpha_raw, alpha), stage 2 = smoothing over interfaces, mortars and MPI faces
    void (*indicator)(const KParams &, int stage, cudaStream_t);
    void (*max_dt)(const KParams &, cudaStream_t);
    // true when the RK stage kernel `element` selects for P honours P.want_cfl (fused max_dt)
    bool (*fuses_cfl)(const KParams &);
    // true when the element kernel selected for P can fetch its - faces from the left neighbours (P.minus_nb)
    bool (*single_face_flux)(const KParams &);
    void (*sfv_fill_right)(const KParams &, cudaStream_t);
    void (*mpi_pack)(const KParams &, cudaStream_t);
    void (*mpi_interface_flux)(const KParams &, cudaStream_t);
    // Force-load every kernel of this table (CUDA loads kernels lazily at first launch, and that load can
    // need a context synchronisation: fatal while a halo wait kernel of another handle is spinning)
    cudaError_t (*preload)();
    int ndims, nvars, nnodes;
};

// which compile-time fast surface flux (see surface_numflux) the tuned path may use for P
template <class EQ>
int fast_surface_flux_mode(const KParams &P) {
    if constexpr (HasFastRanocha<EQ>::value) {
        if (P.kernel_path == 1) return 0;
        if (P.surface_flux == TRIXI_B200_FLUX_RANOCHA || P.surface_flux == TRIXI_B200_FLUX_RANOCHA_TURBO) return 1;
        if (P.surface_flux == TRIXI_B200_FLUX_LLF || P.surface_flux == TRIXI_B200_FLUX_LLF_NAIVE) return 2;
    }
    return 0;
}

template <class EQ, int N>
void launch_interface_flux(const KParams &P, cudaStream_t s) {
    constexpr int NF = ipow(N, EQ::NDIMS - 1);
    const long long total = P.ninterfaces * NF;
    if (total == 0) return;
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if (P.p4est) {
        k_interface_flux_p4est<EQ, N><<<blocks, threads, 0, s>>>(P);
    } else if (P.curved) {
        if constexpr (32 % NF == 0 && !EQ::kHasNoncons) {
            if (P.kernel_path == 0) {
                const long long per_block = 8 * (32 / NF);
                k_interface_flux_staged<EQ, N, 0, true>
                    <<<(unsigned)((P.ninterfaces + per_block - 1) / per_block), 256, 0, s>>>(P);
                return;
            }
        }
        k_interface_flux_curved<EQ, N><<<blocks, threads, 0, s>>>(P);
    } else {
        const int fast = fast_surface_flux_mode<EQ>(P);
        if constexpr (32 % NF == 0) {
            // a warp owns 32 / NF whole interfaces, 8 warps per block
            const long long per_block = 8 * (32 / NF);
            const unsigned sblocks = (unsigned)((P.ninterfaces + per_block - 1) / per_block);
            if (fast == 1)
                k_interface_flux_staged<EQ, N, 1><<<sblocks, 256, 0, s>>>(P);
            else if (fast == 2)
                k_interface_flux_staged<EQ, N, 2><<<sblocks, 256, 0, s>>>(P);
            else
                k_interface_flux_staged<EQ, N><<<sblocks, 256, 0, s>>>(P);
        } else {
            if (fast == 1)
                k_interface_flux<EQ, N, 1><<<blocks, threads, 0, s>>>(P);
            else if (fast == 2)
                k_interface_flux<EQ, N, 2><<<blocks, threads, 0, s>>>(P);
            else
                k_interface_flux<EQ, N><<<blocks, threads, 0, s>>>(P);
        }
    }
}

template <class EQ, int N>
void launch_boundary_flux(const KParams &P, cudaStream_t s) {
    constexpr int NF = ipow(N, EQ::NDIMS - 1);
    const long long total = P.nboundaries * NF;
    if (total == 0) return;
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if (P.p4est)
        k_boundary_flux_p4est<EQ, N><<<blocks, threads, 0, s>>>(P);
    else if (P.curved)
        k_boundary_flux_curved<EQ, N><<<blocks, threads, 0, s>>>(P);
    else
        k_boundary_flux<EQ, N><<<blocks, threads, 0, s>>>(P);
}

template <class EQ, int N>
void launch_mortar_flux(const KParams &P, cudaStream_t s) {
    if (P.nmortars == 0) return;
    constexpr int threads = (1 << (EQ::NDIMS - 1)) * ipow(N, EQ::NDIMS - 1);
    if (P.p4est)
        k_mortar_flux_p4est<EQ, N><<<(unsigned)P.nmortars, threads, 0, s>>>(P);
    else
        k_mortar_flux<EQ, N><<<(unsigned)P.nmortars, threads, 0, s>>>(P);
}

template <class EQ, int N>
void launch_mpi_pack(const KParams &P, cudaStream_t s) {
    constexpr int NF = ipow(N, EQ::NDIMS - 1);
    const long long total = P.nmpi * NF;
    if (total == 0) return;
    if (P.p4est) {
        k_mpi_pack_p4est<EQ, N><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
        return;
    }
    if constexpr (32 % NF == 0 && (NF * EQ::NVARS) % 2 == 0) {
        if (P.kernel_path != 1) {
            const long long per_block = 8 * (32 / NF);
            k_mpi_pack_staged<EQ, N><<<(unsigned)((P.nmpi + per_block - 1) / per_block), 256, 0, s>>>(P);
            return;
        }
    }
    k_mpi_pack<EQ, N><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
}

template <class EQ, int N>
void launch_mpi_conforming_flux(const KParams &P, cudaStream_t s) {
    constexpr int NF = ipow(N, EQ::NDIMS - 1);
    const long long total = P.nmpi * NF;
    if (total == 0) return;
    if (P.p4est)
        k_mpi_interface_flux_p4est<EQ, N><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
    else {
        const int fast = fast_surface_flux_mode<EQ>(P);
        if constexpr (32 % NF == 0) {
            if (P.kernel_path != 1) {
                const long long per_block = 8 * (32 / NF);
                const unsigned sblocks = (unsigned)((P.nmpi + per_block - 1) / per_block);
                if (fast == 1)
                    k_mpi_interface_flux_staged<EQ, N, 1><<<sblocks, 256, 0, s>>>(P);
                else if (fast == 2)
                    k_mpi_interface_flux_staged<EQ, N, 2><<<sblocks, 256, 0, s>>>(P);
                else
                    k_mpi_interface_flux_staged<EQ, N><<<sblocks, 256, 0, s>>>(P);
                return;
            }
        }
        if (fast == 1)
            k_mpi_interface_flux<EQ, N, 1><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
        else if (fast == 2)
            k_mpi_interface_flux<EQ, N, 2><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
        else
            k_mpi_interface_flux<EQ, N><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
    }
}

// shared conforming faces, then the mortars that straddle ranks (both read the same receive buffer)
template <class EQ, int N>
void launch_mpi_interface_flux(const KParams &P, cudaStream_t s) {
    launch_mpi_conforming_flux<EQ, N>(P, s);
    if (P.nmpimortars > 0) {
        constexpr int threads = (1 << (EQ::NDIMS - 1)) * ipow(N, EQ::NDIMS - 1);
        k_mpi_mortar_flux<EQ, N><<<(unsigned)P.nmpimortars, threads, 0, s>>>(P);
    }
}

// cudaFuncSetAttribute is per device: remember what was configured where
struct PerDeviceFlag {
    bool done[64] = {};
    bool test_and_set() {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        const bool was = done[dev];
        done[dev] = true;
        return was;
    }
};

template <class EQ, int N>
void launch_error_norms(const KParams &P, const NormParams &Q, cudaStream_t s) {
    if (P.nelements == 0) return;
    constexpr size_t smem = sizeof(double) * (EQ::NVARS + EQ::NDIMS + 1) * ipow(N, EQ::NDIMS);
    if constexpr (smem > 48 * 1024) {
        static PerDeviceFlag configured;
        if (!configured.test_and_set())
            cudaFuncSetAttribute(k_error_norms<EQ, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    k_error_norms<EQ, N><<<(unsigned)P.nelements, 128, smem, s>>>(P, Q);
}

template <class EQ, int N>
void launch_integrate(const KParams &P, int quantity, double *sums, cudaStream_t s) {
    const long long total = P.nelements * ipow(N, EQ::NDIMS);
    if (total == 0) return;
    k_integrate<EQ, N><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P, quantity, sums);
}

// tuned_euler3d.cu
cudaError_t launch_element_euler3d_ranocha_p3(const KParams &P, bool with_surface, cudaStream_t s);
cudaError_t launch_element_euler3d_ranocha_p3_v7(const KParams &P, bool with_surface, cudaStream_t s);
cudaError_t launch_element_euler3d_ranocha_curved_p3(const KParams &P, bool with_surface, cudaStream_t s);
cudaError_t launch_element_euler3d_ranocha_curved_p5(const KParams &P, bool with_surface, cudaStream_t s);  // tuned_euler3d_p5.cu
cudaError_t preload_tuned_euler3d_curved_p5();
cudaError_t launch_element_euler3d_weak_p3(const KParams &P, bool with_surface, cudaStream_t s);
cudaError_t launch_element_linesweep_euler3d(const KParams &P, bool with_surface, cudaStream_t s);
cudaError_t launch_element_linesweep_sc_euler3d(const KParams &P, bool with_surface, cudaStream_t s);
cudaError_t launch_element_linesweep_mhd3d(const KParams &P, bool with_surface, cudaStream_t s);
cudaError_t preload_linesweep();

template <class EQ, int N, int VOLINT, bool WS>
cudaError_t launch_element_variant(const KParams &P, cudaStream_t s) {
    using C = ElemCfg<EQ, N>;
    const size_t smem = sizeof(double) * ((size_t)C::EPB * C::NN * C::US + N * N +
                                          (VOLINT == TRIXI_B200_VOLINT_WEAK_FORM
                                               ? (size_t)C::ND * C::EPB * C::NN * C::US
                                               : 0));
    const unsigned blocks = (unsigned)(((P.curved ? P.nelements : P.elem_end - P.elem_begin) + C::EPB - 1) / C::EPB);
    if (P.curved) {
        auto kern = k_element_curved<EQ, N, VOLINT, WS>;
        if (smem > 48 * 1024) {
            static PerDeviceFlag configured;
            if (!configured.test_and_set()) {
                cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (err != cudaSuccess) return err;
            }
        }
        kern<<<blocks, C::THREADS, smem, s>>>(P);
        return cudaSuccess;
    }
    auto kern = k_element<EQ, N, VOLINT, WS>;
    if (smem > 48 * 1024) {
        static PerDeviceFlag configured;
        if (!configured.test_and_set()) {
            cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return err;
        }
    }
    kern<<<blocks, C::THREADS, smem, s>>>(P);
    return cudaSuccess;
}

template <class EQ, int N>
bool uses_tuned_element(const KParams &P) {
    if constexpr (std::is_same_v<EQ, Euler<3>> && N == 4) {
        if (P.kernel_path == 1) return false;
        if (P.volume_integral == TRIXI_B200_VOLINT_WEAK_FORM) return true;  // TreeMesh and curved meshes
        if (P.curved)  // flux differencing with flux_ranocha along averaged contravariant vectors
            return P.volume_integral == TRIXI_B200_VOLINT_FLUX_DIFFERENCING &&
                   (P.volume_flux == TRIXI_B200_FLUX_RANOCHA || P.volume_flux == TRIXI_B200_FLUX_RANOCHA_TURBO);
        // (the line-sweep kernels inline Euler::numflux_core: flux_hllc / flux_hlle as volume or subcell fluxes take
        // the generic kernels)
        if (P.volume_integral == TRIXI_B200_VOLINT_FLUX_DIFFERENCING) return Euler<3>::in_core_switch(P.volume_flux);
        return P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG && Euler<3>::in_core_switch(P.volume_flux) &&
               Euler<3>::in_core_switch(P.volume_flux_fv);  // blended line sweep
    }
    if constexpr (std::is_same_v<EQ, Mhd3D> && N == 4) {
        return P.kernel_path != 1 && !P.curved && P.volume_integral == TRIXI_B200_VOLINT_FLUX_DIFFERENCING;
    }
    if constexpr (std::is_same_v<EQ, Euler<3>> && N == 6) {  // curved flux differencing with flux_ranocha at polydeg 5
        return P.kernel_path != 1 && P.curved && P.volume_integral == TRIXI_B200_VOLINT_FLUX_DIFFERENCING &&
               (P.volume_flux == TRIXI_B200_FLUX_RANOCHA || P.volume_flux == TRIXI_B200_FLUX_RANOCHA_TURBO) &&
               P.source_terms == TRIXI_B200_SRC_NONE;
    }
    return false;
}

// true when the RK stage kernel selected for P also reduces the CFL wave speeds (TreeMesh tuned kernels)
template <class EQ, int N>
bool fuses_cfl(const KParams &P) {
    // (on curved meshes only the weak-form kernel reduces the CFL too; the curved flux-differencing kernel leaves it
    // to k_max_dt_curved)
    return uses_tuned_element<EQ, N>(P) && !(P.curved && P.volume_integral != TRIXI_B200_VOLINT_WEAK_FORM);
}

// the TreeMesh headline kernel is the one element kernel that reads its - faces through P.minus_nb
template <class EQ, int N>
bool single_face_flux(const KParams &P) {
    if constexpr (std::is_same_v<EQ, Euler<3>> && N == 4) {
        return P.kernel_path == 0 && !P.curved && P.minus_nb != nullptr &&
               P.volume_integral == TRIXI_B200_VOLINT_FLUX_DIFFERENCING &&
               (P.volume_flux == TRIXI_B200_FLUX_RANOCHA || P.volume_flux == TRIXI_B200_FLUX_RANOCHA_TURBO);
    }
    return false;
}

template <class EQ, int N>
void launch_sfv_fill_right(const KParams &P, cudaStream_t s) {
    const long long total = P.ninterfaces * ipow(N, EQ::NDIMS - 1) * EQ::NVARS;
    if (total == 0 || P.curved) return;
    k_sfv_fill_right<EQ, N><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
}

template <class EQ, int N>
cudaError_t launch_element(const KParams &P, bool with_surface, cudaStream_t s) {
    if (P.nelements == 0) return cudaSuccess;
    if constexpr (std::is_same_v<EQ, Euler<3>> && N == 4) {
        if (uses_tuned_element<EQ, N>(P)) {
            if (P.volume_integral == TRIXI_B200_VOLINT_WEAK_FORM) return launch_element_euler3d_weak_p3(P, with_surface, s);
            if (P.curved) return launch_element_euler3d_ranocha_curved_p3(P, with_surface, s);
            if (P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG)
                return launch_element_linesweep_sc_euler3d(P, with_surface, s);
            if (P.volume_flux == TRIXI_B200_FLUX_RANOCHA || P.volume_flux == TRIXI_B200_FLUX_RANOCHA_TURBO)
                return P.kernel_path == 2 ? launch_element_euler3d_ranocha_p3_v7(P, with_surface, s)
                                          : launch_element_euler3d_ranocha_p3(P, with_surface, s);
            return launch_element_linesweep_euler3d(P, with_surface, s);
        }
    }
    if constexpr (std::is_same_v<EQ, Mhd3D> && N == 4) {
        if (uses_tuned_element<EQ, N>(P)) return launch_element_linesweep_mhd3d(P, with_surface, s);
    }
    if constexpr (std::is_same_v<EQ, Euler<3>> && N == 6) {
        if (uses_tuned_element<EQ, N>(P)) return launch_element_euler3d_ranocha_curved_p5(P, with_surface, s);
    }
    if (P.volume_integral == TRIXI_B200_VOLINT_WEAK_FORM) {
        return with_surface ? launch_element_variant<EQ, N, TRIXI_B200_VOLINT_WEAK_FORM, true>(P, s)
                            : launch_element_variant<EQ, N, TRIXI_B200_VOLINT_WEAK_FORM, false>(P, s);
    }
    if (P.volume_integral == TRIXI_B200_VOLINT_PURE_LGL_FV)
        return with_surface ? launch_element_variant<EQ, N, TRIXI_B200_VOLINT_PURE_LGL_FV, true>(P, s)
                            : launch_element_variant<EQ, N, TRIXI_B200_VOLINT_PURE_LGL_FV, false>(P, s);
    if constexpr (HasFastRanocha<EQ>::value || EQ::kHasNoncons) {  // compressible Euler, ideal GLM-MHD
        if (P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG)
            return with_surface ? launch_element_variant<EQ, N, TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG, true>(P, s)
                                : launch_element_variant<EQ, N, TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG, false>(P, s);
    }
    return with_surface ? launch_element_variant<EQ, N, TRIXI_B200_VOLINT_FLUX_DIFFERENCING, true>(P, s)
                        : launch_element_variant<EQ, N, TRIXI_B200_VOLINT_FLUX_DIFFERENCING, false>(P, s);
}

// (indicator_hg::IndicatorHennemannGassner)(u, mesh, equations, dg, cache) (dgsem/indicators.jl:114-148)
template <class EQ, int N>
void launch_indicator(const KParams &P, int stage, cudaStream_t s) {
    if constexpr (HasFastRanocha<EQ>::value || EQ::kHasNoncons) {  // compressible Euler, ideal GLM-MHD
        using C = ElemCfg<EQ, N>;
        if (P.nelements == 0) return;
        if (stage == 1) {
            // magic parameters (indicators.jl:126-130)
            const double threshold = 0.5 * pow(10.0, -1.8 * pow((double)N, 0.25));
            const double parameter_s = log((1 - 0.0001) / 0.0001);
            if constexpr (EQ::NDIMS == 3 && N == 4) {
                if (P.kernel_path != 1) {  // warp per element
                    k_indicator_hg_3d_p3<EQ><<<(unsigned)((P.nelements + 7) / 8), 256, 0, s>>>(P, threshold, parameter_s);
                    return;
                }
            }
            k_indicator_hg<EQ, N><<<(unsigned)((P.nelements + C::EPB - 1) / C::EPB), C::THREADS, 0, s>>>(P, threshold, parameter_s);
        } else {
            const long long faces = P.ninterfaces + P.nmortars + P.nmpi;
            if (P.ind_smooth && faces > 0) k_indicator_smooth<EQ, N><<<(unsigned)((faces + 255) / 256), 256, 0, s>>>(P);
        }
    }
}

template <class EQ, int N>
void launch_max_dt(const KParams &P, cudaStream_t s) {
    using C = ElemCfg<EQ, N>;
    if (P.nelements == 0) return;
    const unsigned blocks = (unsigned)((P.nelements + C::EPB - 1) / C::EPB);
    if (P.curved)
        k_max_dt_curved<EQ, N><<<blocks, C::THREADS, 0, s>>>(P);
    else
        k_max_dt<EQ, N><<<blocks, C::THREADS, 0, s>>>(P);
}

template <class K>
cudaError_t preload_kernel(K kern) {
    cudaFuncAttributes attr;
    return cudaFuncGetAttributes(&attr, kern);
}

cudaError_t preload_tuned_euler3d();       // tuned_euler3d.cu
cudaError_t preload_tuned_euler3d_v7();
cudaError_t preload_tuned_euler3d_curved();
cudaError_t preload_tuned_euler3d_weak();  // tuned_euler3d.cu

template <class EQ, int N>
cudaError_t preload_all() {
    cudaError_t e;
#define TB_PRELOAD(k)                      \
    if ((e = preload_kernel(k)) != cudaSuccess) return e
    TB_PRELOAD((k_interface_flux<EQ, N>));
    TB_PRELOAD((k_interface_flux<EQ, N, 1>));
    TB_PRELOAD((k_interface_flux<EQ, N, 2>));
    if constexpr (32 % ipow(N, EQ::NDIMS - 1) == 0) {
        TB_PRELOAD((k_interface_flux_staged<EQ, N>));
        TB_PRELOAD((k_interface_flux_staged<EQ, N, 1>));
        TB_PRELOAD((k_interface_flux_staged<EQ, N, 2>));
        if constexpr (!EQ::kHasNoncons) TB_PRELOAD((k_interface_flux_staged<EQ, N, 0, true>));
        if constexpr ((ipow(N, EQ::NDIMS - 1) * EQ::NVARS) % 2 == 0) TB_PRELOAD((k_mpi_pack_staged<EQ, N>));
        TB_PRELOAD((k_mpi_interface_flux_staged<EQ, N>));
        TB_PRELOAD((k_mpi_interface_flux_staged<EQ, N, 1>));
        TB_PRELOAD((k_mpi_interface_flux_staged<EQ, N, 2>));
    }
    TB_PRELOAD((k_mpi_interface_flux<EQ, N, 1>));
    TB_PRELOAD((k_mpi_interface_flux<EQ, N, 2>));
    TB_PRELOAD((k_boundary_flux<EQ, N>));
    TB_PRELOAD((k_sfv_fill_right<EQ, N>));
    TB_PRELOAD((k_mortar_flux<EQ, N>));
    TB_PRELOAD((k_integrate<EQ, N>));
    TB_PRELOAD((k_mortar_flux_p4est<EQ, N>));
    TB_PRELOAD((k_mpi_mortar_flux<EQ, N>));
    TB_PRELOAD((k_error_norms<EQ, N>));
    TB_PRELOAD((k_mpi_pack<EQ, N>));
    TB_PRELOAD((k_mpi_interface_flux<EQ, N>));
    TB_PRELOAD((k_max_dt<EQ, N>));
    TB_PRELOAD((k_max_dt_curved<EQ, N>));
    TB_PRELOAD((k_interface_flux_curved<EQ, N>));
    TB_PRELOAD((k_interface_flux_p4est<EQ, N>));
    TB_PRELOAD((k_boundary_flux_p4est<EQ, N>));
    TB_PRELOAD((k_mpi_pack_p4est<EQ, N>));
    TB_PRELOAD((k_mpi_interface_flux_p4est<EQ, N>));
    TB_PRELOAD((k_boundary_flux_curved<EQ, N>));
    TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_WEAK_FORM, true>));
    TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_WEAK_FORM, false>));
    TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_FLUX_DIFFERENCING, true>));
    TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_FLUX_DIFFERENCING, false>));
    TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_WEAK_FORM, true>));
    TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_WEAK_FORM, false>));
    TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_FLUX_DIFFERENCING, true>));
    TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_FLUX_DIFFERENCING, false>));
    TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_PURE_LGL_FV, true>));
    TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_PURE_LGL_FV, false>));
    TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_PURE_LGL_FV, true>));
    TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_PURE_LGL_FV, false>));
    if constexpr (HasFastRanocha<EQ>::value || EQ::kHasNoncons) {
        TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG, true>));
        TB_PRELOAD((k_element<EQ, N, TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG, false>));
        TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG, true>));
        TB_PRELOAD((k_element_curved<EQ, N, TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG, false>));
        TB_PRELOAD((k_indicator_hg<EQ, N>));
        if constexpr (EQ::NDIMS == 3 && N == 4) TB_PRELOAD((k_indicator_hg_3d_p3<EQ>));
        TB_PRELOAD((k_indicator_smooth<EQ, N>));
    }
#undef TB_PRELOAD
    if constexpr (std::is_same_v<EQ, Euler<3>> && N == 4) {
        if ((e = preload_tuned_euler3d()) != cudaSuccess) return e;
        if ((e = preload_tuned_euler3d_v7()) != cudaSuccess) return e;
        if ((e = preload_tuned_euler3d_curved()) != cudaSuccess) return e;
        if ((e = preload_tuned_euler3d_weak()) != cudaSuccess) return e;
        return preload_linesweep();
    }
    if constexpr (std::is_same_v<EQ, Mhd3D> && N == 4) return preload_linesweep();
    if constexpr (std::is_same_v<EQ, Euler<3>> && N == 6) return preload_tuned_euler3d_curved_p5();
    return cudaSuccess;
}

template <class EQ, int N>
const Launchers *make_launchers() {
    static const Launchers L = {&launch_interface_flux<EQ, N>,
                                &launch_boundary_flux<EQ, N>,
                                &launch_mortar_flux<EQ, N>,
                                &launch_error_norms<EQ, N>,
                                &launch_integrate<EQ, N>,
                                &launch_element<EQ, N>,
                                &launch_indicator<EQ, N>,
                                &launch_max_dt<EQ, N>,
                                &fuses_cfl<EQ, N>,
                                &single_face_flux<EQ, N>,
                                &launch_sfv_fill_right<EQ, N>,
                                &launch_mpi_pack<EQ, N>,
                                &launch_mpi_interface_flux<EQ, N>,
                                &preload_all<EQ, N>,
                                EQ::NDIMS,
                                EQ::NVARS,
                                N};
    return &L;
}

// the launcher table for nnodes = n if n is one of NS..., else nullptr: a translation unit instantiates the kernels of the
// listed node counts only (the per-equation tables are split over several units so that they compile in parallel)
template <class EQ, int... NS>
const Launchers *launchers_among(int n) {
    const Launchers *r = nullptr;
    ((n == NS ? (void)(r = make_launchers<EQ, NS>()) : (void)0), ...);
    return r;
}

template <class EQ>
const Launchers *launchers_for_nnodes(int n) {
    switch (n) {
    case 2: return make_launchers<EQ, 2>();
    case 3: return make_launchers<EQ, 3>();
    case 4: return make_launchers<EQ, 4>();
    case 5: return make_launchers<EQ, 5>();
    case 6: return make_launchers<EQ, 6>();
    case 7: return make_launchers<EQ, 7>();
    case 8: return make_launchers<EQ, 8>();
    default: return nullptr;
    }
}

// one translation unit per equation (parallel nvcc)
const Launchers *get_launchers_advection2d(int nnodes);
const Launchers *get_launchers_advection3d(int nnodes);
const Launchers *get_launchers_euler2d(int nnodes);
const Launchers *get_launchers_euler3d(int nnodes);
const Launchers *get_launchers_mhd3d(int nnodes);
// second tables of the compressible Euler equations: every registered flux in the run-time switches (EulerAllFluxes)
const Launchers *get_launchers_euler2d_all(int nnodes);
const Launchers *get_launchers_euler3d_all(int nnodes);
const Launchers *get_launchers_euler2d_all_hi(int nnodes);
const Launchers *get_launchers_euler3d_all_n5(int nnodes);
const Launchers *get_launchers_euler3d_all_n6(int nnodes);
const Launchers *get_launchers_euler3d_all_n7(int nnodes);
const Launchers *get_launchers_euler3d_all_n8(int nnodes);
// (parts of the tables above, one translation unit each)
const Launchers *get_launchers_euler2d_hi(int nnodes);
const Launchers *get_launchers_euler3d_n5(int nnodes);
const Launchers *get_launchers_euler3d_n6(int nnodes);
const Launchers *get_launchers_euler3d_n7(int nnodes);
const Launchers *get_launchers_euler3d_n8(int nnodes);
const Launchers *get_launchers_mhd3d_n5(int nnodes);
const Launchers *get_launchers_mhd3d_n6(int nnodes);
const Launchers *get_launchers_mhd3d_n7(int nnodes);
const Launchers *get_launchers_mhd3d_n8(int nnodes);

}  // namespace tb
