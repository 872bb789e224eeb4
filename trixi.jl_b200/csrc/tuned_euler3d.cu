// Translation unit of the tuned 3D Euler kernels (headline configuration).
#include "kernel_euler3d_fd_p3.cuh"
