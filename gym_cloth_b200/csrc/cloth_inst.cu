// cloth_inst.cu - the step kernel for one scalar type (CLOTH_T) and one compile-time grid width (CLOTH_INSTANTIATE_WC:
// 25, 64, or 0 = any width at run time).  build.py compiles this file six times, in parallel.
#include <cstring>
#include "cloth_kernels.cuh"
