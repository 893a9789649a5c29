// tc_stub.cu -- placeholders for tcgen05 kernels that are not written yet: they decline every shape
// so the dispatcher uses the CUDA-core kernels.  Each one is replaced by a real kernel file.
#include "common.cuh"
#include "kernels.h"

#ifndef OMR_HAVE_TC_ATTN
int omr_attn_fwd_tc(const void*, long long, long long, const void*, long long, long long, const void*, long long,
                    long long, void*, long long, long long, float*, const float*, int, int, int, int, int, float, int,
                    int, const int*, const int*, int, cudaStream_t) {
  return OMR_TC_NOT_ELIGIBLE;
}
#endif
