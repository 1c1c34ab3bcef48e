// cbmd_lib.cu — single translation unit of libcbmd_cuda.so (kernels launch each
// other's helpers, so the pieces are compiled together instead of with -rdc).
#include "cbmd_system.cu"
#include "cbmd_integrate.cu"
#include "cbmd_binning.cu"
#include "cbmd_neighbor.cu"
#include "cbmd_force.cu"
#include "cbmd_comm.cu"
#include "cbmd_steps.cu"
