// Explicit instantiation of the generic kernels for one equation system (parallel compilation unit).
#include "launch.cuh"
namespace tb {
const Launchers *get_launchers_euler3d_all(int nnodes) {
    switch (nnodes) {
    case 5: return get_launchers_euler3d_all_n5(nnodes);
    case 6: return get_launchers_euler3d_all_n6(nnodes);
    case 7: return get_launchers_euler3d_all_n7(nnodes);
    case 8: return get_launchers_euler3d_all_n8(nnodes);
    default: return launchers_among<EulerAllFluxes<3>, 2, 3, 4>(nnodes);
    }
}
}  // namespace tb
