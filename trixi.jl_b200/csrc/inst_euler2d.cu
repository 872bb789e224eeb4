// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_euler2d(int nnodes) { return launchers_for_nnodes<Euler<2>>(nnodes); }
}  // namespace tb
