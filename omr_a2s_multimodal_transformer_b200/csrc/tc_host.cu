// tc_host.cu -- host-side helper of the tensor-core kernels: CUtensorMap encoding through the driver entry
// point resolved at run time (no link-time libcuda dependency: the library must load on a GPU-less host for
// the ABI tests).
#include <cudaTypedefs.h>

#include "tc_common.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int omr_make_tensor_map(CUtensorMap* out, int elem_bytes, const void* base, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned int* box, const unsigned int* elem_stride,
                        int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    omr_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return OMR_ERR_CUDA;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_stride ? elem_stride[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMapDataType dtp = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dtp, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    omr_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] stride0 %llu", (int)r,
                  rank, dims[0], rank > 1 ? dims[1] : 0ull, rank > 2 ? dims[2] : 0ull, rank > 3 ? dims[3] : 0ull, box[0],
                  rank > 1 ? box[1] : 0u, rank > 2 ? box[2] : 0u, rank > 3 ? box[3] : 0u, rank > 1 ? strides_bytes[0] : 0ull);
    return OMR_ERR_CUDA;
  }
  return OMR_OK;
}
