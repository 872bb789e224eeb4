// Translation unit of the tuned 3D kernels at polydeg 3: the headline flux_ranocha kernel, the weak-form kernel
// and the line-sweep flux-differencing kernel for the other two-point fluxes and GLM-MHD.
#include "kernel_euler3d_fd_p3.cuh"
#include "kernel_euler3d_fd_p3_v7.cuh"
#include "kernel_euler3d_fd_curved_p3.cuh"
#include "kernel_euler3d_weak_p3.cuh"
#include "kernel_fd3d_p3.cuh"

namespace tb {
cudaError_t launch_element_linesweep_euler3d(const KParams &P, bool with_surface, cudaStream_t s) {
    return launch_element_fd3d_p3<Euler<3>>(P, with_surface, s);
}
cudaError_t launch_element_linesweep_sc_euler3d(const KParams &P, bool with_surface, cudaStream_t s) {
    return launch_element_fd3d_p3<Euler<3>, true>(P, with_surface, s);
}
cudaError_t launch_element_linesweep_mhd3d(const KParams &P, bool with_surface, cudaStream_t s) {
    return launch_element_fd3d_p3<Mhd3D>(P, with_surface, s);
}
cudaError_t preload_linesweep() {
    cudaError_t e = preload_fd3d_p3<Euler<3>>();
    if (e != cudaSuccess) return e;
    if ((e = preload_fd3d_p3<Euler<3>, true>()) != cudaSuccess) return e;
    return preload_fd3d_p3<Mhd3D>();
}
}  // namespace tb
