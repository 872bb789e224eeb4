// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_euler2d_all_hi(int nnodes) { return launchers_among<EulerAllFluxes<2>, 6, 7, 8>(nnodes); }
}  // namespace tb
