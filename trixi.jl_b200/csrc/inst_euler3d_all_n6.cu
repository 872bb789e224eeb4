// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_euler3d_all_n6(int nnodes) { return launchers_among<EulerAllFluxes<3>, 6>(nnodes); }
}  // namespace tb
