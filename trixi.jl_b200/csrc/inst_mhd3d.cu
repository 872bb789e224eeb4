// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_mhd3d(int nnodes) {
    switch (nnodes) {
    case 5: return get_launchers_mhd3d_n5(nnodes);
    case 6: return get_launchers_mhd3d_n6(nnodes);
    case 7: return get_launchers_mhd3d_n7(nnodes);
    case 8: return get_launchers_mhd3d_n8(nnodes);
    default: return launchers_among<Mhd3D, 2, 3, 4>(nnodes);
    }
}
}  // namespace tb
