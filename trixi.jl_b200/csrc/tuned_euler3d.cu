// Translation unit of the tuned 3D Euler kernels (headline configuration and the weak-form configuration).
#include "kernel_euler3d_fd_p3.cuh"
#include "kernel_euler3d_weak_p3.cuh"
