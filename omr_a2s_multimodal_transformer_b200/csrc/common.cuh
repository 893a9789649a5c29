// common.cuh -- shared helpers for libomr_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/omr_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing -----------------------------------------------------------------------
void omr_set_error(const char* fmt, ...);
void omr_count_launch(int n = 1);

#define OMR_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      omr_set_error(__VA_ARGS__);   \
      return OMR_ERR_INVALID;       \
    }                               \
  } while (0)

#define OMR_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      omr_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return OMR_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// after a <<<>>> launch
#define OMR_LAUNCHED()                                                                   \
  do {                                                                                   \
    omr_count_launch();                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      omr_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return OMR_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// ---- programmatic dependent launch ------------------------------------------------------------
// Every kernel opens with omr_pdl_enter(): it lets the NEXT kernel of the stream start scheduling its CTAs while this
// grid drains (launch_dependents), then blocks until the PREVIOUS grid has completed and its writes are visible
// (wait) -- the data dependency is exactly that of ordinary stream order; only launch latency and the CTA ramp overlap.
// Both instructions are no-ops for a grid launched without the attribute (OMR_PDL=0, or a predecessor that is not a
// kernel).
__device__ __forceinline__ void omr_pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}
bool omr_pdl_enabled();  // api.cu: OMR_PDL=1 (default off)

struct OmrLaunch {
  cudaLaunchConfig_t lc;
  cudaLaunchAttribute at[1];
  OmrLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    lc = cudaLaunchConfig_t{};
    lc.gridDim = grid;
    lc.blockDim = block;
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = omr_pdl_enabled() ? 1 : 0;
  }
  template <typename... KArgs, typename... Args>
  void operator()(void (*kernel)(KArgs...), Args&&... args) {
    (void)cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);  // the error is picked up by OMR_LAUNCHED()
  }
};

#define OMR_DISPATCH_DT(dt, T, ...)                 \
  do {                                              \
    if ((dt) == OMR_F32) {                          \
      typedef float T;                              \
      __VA_ARGS__;                                  \
    } else if ((dt) == OMR_BF16) {                  \
      typedef bf16 T;                               \
      __VA_ARGS__;                                  \
    } else {                                        \
      omr_set_error("unsupported dtype code %d", (int)(dt)); \
      return OMR_ERR_INVALID;                       \
    }                                               \
  } while (0)

static inline cudaStream_t as_stream(omr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// ---- scalar conversion ----------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
// value after a round trip through the storage type (keeps fwd/bwd statistics self-consistent)
template <typename T>
__device__ __forceinline__ float round_to(float v) { return to_f(from_f<T>(v)); }

// ---- 4-wide vector access (16B for f32, 8B for bf16); pointer must be suitably aligned ---------
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// ---- warp / block reductions ------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in every thread; sm must hold >= 33 floats
__device__ __forceinline__ float block_sum(float v, float* sm) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? sm[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sm) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? sm[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}
