// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_euler2d_all(int nnodes) {
    return nnodes >= 6 ? get_launchers_euler2d_all_hi(nnodes) : launchers_among<EulerAllFluxes<2>, 2, 3, 4, 5>(nnodes);
}
}  // namespace tb
