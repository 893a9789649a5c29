// tc_stub.cu -- placeholders for tcgen05 kernels that are not written yet: they decline every shape
// so the dispatcher uses the CUDA-core kernels.  Each one is replaced by a real kernel file.
#include "common.cuh"
#include "kernels.h"

