// conv_first.cu -- weight gradient of the FIRST convolution of each encoder (C_in = 1: the grey-scale score image /
// the spectrogram).  With one input channel the "GEMM" has K = 9 and is purely HBM-bound (it streams dY once:
// 16 channels x 4-5 M pixels), so it runs on the CUDA cores:
//   dW[co, 0, kh, kw] (+)= sum_p dY[p, co] * X[p shifted by (kh-1, kw-1)]
// thread = (pixel, group of 8 output channels): 72 fp32 accumulators, 16-byte dY load, nine 2/4-byte X loads that
// hit L1; partial sums are reduced over the warp with shuffles, over the block in shared memory and added to the
// gradient with 9*Co atomics per block.
#include "common.cuh"
#include "kernels.h"

namespace {

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}

template <typename T>
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                                                          int N, int H, int W, int Co, int Ho, int Wo, int sh, int sw) {
  omr_pdl_enter();
  extern __shared__ float red[];  // [8 warps][groups * 72]
  const int groups = Co / 8;
  const long long npix = (long long)N * Ho * Wo;
  const long long total = npix * groups;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  // consecutive threads: consecutive channel groups of the same pixel, then the next pixel (coalesced dY reads);
  // the stride keeps every thread on ONE channel group for its whole life
  // (32-bit index arithmetic: the launcher declines tensors of 2^31 elements or more; 64-bit divisions here used to
  // cost more than the 72 FMAs of a pixel)
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;  // multiple of groups (256 % groups == 0)
  const int grp = (int)(tid % (unsigned)groups);
  for (unsigned i = tid; i < (unsigned)total; i += stride) {
    const unsigned p = i / (unsigned)groups;
    const int ow = (int)(p % (unsigned)Wo);
    const unsigned r = p / (unsigned)Wo;
    const int oh = (int)(r % (unsigned)Ho), n = (int)(r / (unsigned)Ho);
    float g[8];
    load8<T>(dy + (long long)p * Co + grp * 8, g);
    const T* xn = x + (long long)n * H * W;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * sh + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * sw + kw - 1;
        const float xv = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? to_f(xn[(long long)ih * W + iw]) : 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[kh * 3 + kw][c] = fmaf(xv, g[c], acc[kh * 3 + kw][c]);
      }
    }
  }
  // lanes with the same (lane % groups) hold the same channel group: butterfly over the other lane bits
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v = acc[t][c];
      for (int o = 16; o >= groups; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[t][c] = v;
    }
  if (lane < groups) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < 8; ++c) red[(wid * groups + lane) * 72 + t * 8 + c] = acc[t][c];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < groups * 72; idx += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w * groups * 72 + idx];
    const int gq = idx / 72, rem = idx - gq * 72, t = rem / 8, c = rem - t * 8;
    atomicAdd(dw + (long long)(gq * 8 + c) * 9 + t, s);
  }
}

// ---- forward of the first convolution: y[p, 0..15] = relu(bias + sum_taps x[p + tap] * w[co, tap]) ----------------------
// One thread per output pixel: nine 2/4-byte input loads that hit L1, 144 FMAs against weights broadcast from shared
// memory, one 32-byte (bf16) store -- a warp writes 1 KB contiguous.  Pure output-write bound: 2 * 16 * pixels bytes.
template <typename T>
__global__ void __launch_bounds__(256, 3) conv1_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                                        T* __restrict__ y, int N, int H, int W, int relu) {
  omr_pdl_enter();
  __shared__ __align__(16) float sw[9][16];
  __shared__ float sb[16];
  if (threadIdx.x < 144) sw[threadIdx.x % 9][threadIdx.x / 9] = to_f(w[threadIdx.x]);  // w: [16][3][3][1]
  if (threadIdx.x < 16) sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const unsigned total = (unsigned)N * (unsigned)H * (unsigned)W;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int ow = (int)(p % (unsigned)W);
    const unsigned r = p / (unsigned)W;
    const int oh = (int)(r % (unsigned)H);
    const T* xc = x + p;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = sb[c];
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {  // not unrolled: keeps the weight rows of one kernel row (48 floats) live, not all 144
      const int ih = oh + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow + kw - 1;
        const float xv = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? to_f(xc[(kh - 1) * W + (kw - 1)]) : 0.f;
        const float4* wr = reinterpret_cast<const float4*>(sw[kh * 3 + kw]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
          acc[4 * q] = fmaf(xv, wv.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
        }
      }
    }
    if (relu) {
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = fmaxf(acc[c], 0.f);
    }
    T* yp = y + (long long)p * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v4[4] = {acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]};
      store4(yp + 4 * q, v4);
    }
  }
}

}  // namespace

// first-layer forward: Ci == 1, Co == 16, stride 1 (the reference's encoder, encoder.py:132-137 with in_channels = 1)
int omr_conv3x3_fwd_c1(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Co, int sh,
                       int sw, int relu, cudaStream_t st) {
  if (Co != 16 || sh != 1 || sw != 1 || (reinterpret_cast<uintptr_t>(y) & 15) != 0) return OMR_TC_NOT_ELIGIBLE;
  const long long total = (long long)N * H * W;
  if (total <= 0 || total >= (1LL << 31) - (1LL << 22)) return OMR_TC_NOT_ELIGIBLE;
  long long blocks = cdiv(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)blocks, 256, 0, st)(conv1_fwd_kernel<T>, (const T*)x, (const T*)w, bias, (T*)y, N, H, W, relu)));
  OMR_LAUNCHED();
  return OMR_OK;
}

// Ci == 1 only; groups = Co/8 must divide 32.  Returns OMR_TC_NOT_ELIGIBLE for other shapes.
int omr_conv3x3_wgrad_c1(int dt, const void* x, const void* dy, float* dw, int N, int H, int W, int Co, int sh, int sw,
                         int accumulate, cudaStream_t st) {
  if (Co % 8 != 0) return OMR_TC_NOT_ELIGIBLE;
  const int groups = Co / 8;
  if (groups > 32 || (32 % groups) != 0) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(dy) & 15) != 0) return OMR_TC_NOT_ELIGIBLE;
  const int Ho = (H + sh - 1) / sh, Wo = (W + sw - 1) / sw;
  if (!accumulate) OMR_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Co * 9, st));
  const long long total = (long long)N * Ho * Wo * groups;
  if (total >= (1LL << 31) - (1LL << 22)) return OMR_TC_NOT_ELIGIBLE;
  long long blocks = cdiv(total, 256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const size_t smem = sizeof(float) * 8 * groups * 72;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)blocks, 256, smem, st)(conv1_wgrad_kernel<T>, (const T*)x, (const T*)dy, dw, N, H, W, Co, Ho,
                                                                                    Wo, sh, sw)));
  OMR_LAUNCHED();
  return OMR_OK;
}
