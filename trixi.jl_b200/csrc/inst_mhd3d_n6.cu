// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_mhd3d_n6(int nnodes) { return launchers_among<Mhd3D, 6>(nnodes); }
}  // namespace tb
