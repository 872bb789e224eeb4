// Translation unit of the tuned curved flux-differencing kernel at polydeg 5 (the reference's GPU benchmark).
#include "kernel_euler3d_fd_curved_pn.cuh"

namespace tb {
cudaError_t launch_element_euler3d_ranocha_curved_p5(const KParams &P, bool with_surface, cudaStream_t s) {
    // (TRIXI_B200_OPT_KERNEL_PATH = 2 selects the 4-elements-per-block variant for A/B measurements)
    if (P.kernel_path == 2) return launch_element_euler3d_ranocha_curved_pn<6, 4>(P, with_surface, s);
    return launch_element_euler3d_ranocha_curved_pn<6, 3>(P, with_surface, s);
}
cudaError_t preload_tuned_euler3d_curved_p5() {
    cudaError_t e = preload_tuned_euler3d_curved_pn<6, 3>();
    if (e != cudaSuccess) return e;
    return preload_tuned_euler3d_curved_pn<6, 4>();
}
}  // namespace tb
