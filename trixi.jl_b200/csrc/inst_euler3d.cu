// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_euler3d(int nnodes) { return launchers_for_nnodes<Euler<3>>(nnodes); }
}  // namespace tb
