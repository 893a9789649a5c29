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

// 8 consecutive elements as raw 16-byte words (bf16: one, fp32: two), unpacked at use
template <typename T> struct RawOf;
template <> struct RawOf<bf16> {
  typedef uint4 type;
  static __device__ __forceinline__ type load(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const type& t, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
};
template <> struct RawOf<float> {
  struct type { float4 a, b; };
  static __device__ __forceinline__ type load(const float* p) {
    type t;
    t.a = *reinterpret_cast<const float4*>(p);
    t.b = *reinterpret_cast<const float4*>(p + 4);
    return t;
  }
  static __device__ __forceinline__ void unpack(const type& t, float (&v)[8]) {
    v[0] = t.a.x; v[1] = t.a.y; v[2] = t.a.z; v[3] = t.a.w; v[4] = t.b.x; v[5] = t.b.y; v[6] = t.b.z; v[7] = t.b.w;
  }
};

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

// Strip walk (round 2, stride 1): a thread owns a strip of STRIP consecutive output pixels of one row and 8 output channels.
// Walking along the row it keeps the 3 x 3 input window in registers (three new 2-byte loads per pixel instead of nine, no
// division per pixel) -- what is left per pixel is the 16-byte dY load and the 72 FMAs, i.e. the kernel sits at its fp32
// FMA floor instead of 7x above the HBM floor (180 us for 161 MB of dY at 195 x 808 x 32 before).
template <typename T, int STRIP>
__global__ void __launch_bounds__(256, 2) conv1_wgrad_strip_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                   float* __restrict__ dw, int N, int H, int W, int Co) {
  omr_pdl_enter();
  extern __shared__ float red[];  // [8 warps][groups * 72]
  const int groups = Co / 8;
  const int strips_w = (W + STRIP - 1) / STRIP;
  const unsigned total = (unsigned)N * (unsigned)H * (unsigned)strips_w * (unsigned)groups;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;  // multiple of groups (256 % groups == 0)
  const int grp = (int)(tid % (unsigned)groups);
  for (unsigned i = tid; i < total; i += stride) {
    unsigned q = i / (unsigned)groups;  // (n, oh, strip)
    const int sidx = (int)(q % (unsigned)strips_w);
    q /= (unsigned)strips_w;
    const int oh = (int)(q % (unsigned)H), n = (int)(q / (unsigned)H);
    const int w0 = sidx * STRIP, w1 = min(W, w0 + STRIP);
    const T* xr1 = x + ((long long)n * H + oh) * W;  // row oh
    const bool up = oh > 0, dn = oh + 1 < H;
    const T* xr0 = xr1 - W;
    const T* xr2 = xr1 + W;
    const T* gp = dy + (((long long)n * H + oh) * W + w0) * Co + grp * 8;
    // window columns (ow - 1, ow, ow + 1) of rows (oh - 1, oh, oh + 1)
    float a0 = 0.f, a1, a2 = 0.f, b0 = 0.f, b1, b2 = 0.f;  // a: column ow - 1, b: column ow
    if (w0 > 0) {
      a1 = to_f(xr1[w0 - 1]);
      if (up) a0 = to_f(xr0[w0 - 1]);
      if (dn) a2 = to_f(xr2[w0 - 1]);
    } else {
      a1 = 0.f;
    }
    b1 = to_f(xr1[w0]);
    if (up) b0 = to_f(xr0[w0]);
    if (dn) b2 = to_f(xr2[w0]);
    // four pixels per iteration: their dY rows (raw 16-byte words) and input columns are all requested before the first
    // FMA, so that a thread keeps 64 bytes of the dY stream in flight instead of 16
    for (int ow = w0; ow < w1; ow += 4) {
      typename RawOf<T>::type gr[4];
      float cc[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cc[u][0] = cc[u][1] = cc[u][2] = 0.f;
        if (ow + u < w1) {
          gr[u] = RawOf<T>::load(gp + (long long)u * Co);
          if (ow + u + 1 < W) {
            cc[u][1] = to_f(xr1[ow + u + 1]);
            if (up) cc[u][0] = to_f(xr0[ow + u + 1]);
            if (dn) cc[u][2] = to_f(xr2[ow + u + 1]);
          }
        }
      }
      gp += 4 * Co;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ow + u < w1) {
          float g[8];
          RawOf<T>::unpack(gr[u], g);
          const float c0 = cc[u][0], c1 = cc[u][1], c2 = cc[u][2];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            acc[0][c] = fmaf(a0, g[c], acc[0][c]);
            acc[1][c] = fmaf(b0, g[c], acc[1][c]);
            acc[2][c] = fmaf(c0, g[c], acc[2][c]);
            acc[3][c] = fmaf(a1, g[c], acc[3][c]);
            acc[4][c] = fmaf(b1, g[c], acc[4][c]);
            acc[5][c] = fmaf(c1, g[c], acc[5][c]);
            acc[6][c] = fmaf(a2, g[c], acc[6][c]);
            acc[7][c] = fmaf(b2, g[c], acc[7][c]);
            acc[8][c] = fmaf(c2, g[c], acc[8][c]);
          }
          a0 = b0; a1 = b1; a2 = b2;
          b0 = c0; b1 = c1; b2 = c2;
        }
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

// Strip walk of the forward (round 2): thread = (row, strip of STRIP output pixels, 8 of the 16 output channels) with its 72
// weights and 8 biases in REGISTERS; walking along the row it shifts the 3 x 3 input window (three new loads per pixel) and
// writes one 16-byte (bf16) result per pixel.  No shared memory, no division per pixel: 72 FMAs per 16 bytes written.
template <typename T> struct Store8;
template <> struct Store8<bf16> {
  static __device__ __forceinline__ void st(bf16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
template <> struct Store8<float> {
  static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <typename T, int STRIP>
__global__ void __launch_bounds__(256, 2) conv1_fwd_strip_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                 const float* __restrict__ bias, T* __restrict__ y, int N, int H, int W,
                                                                 int relu) {
  omr_pdl_enter();
  const int strips_w = (W + STRIP - 1) / STRIP;
  const unsigned total = (unsigned)N * (unsigned)H * (unsigned)strips_w * 2u;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;  // even
  const int grp = (int)(tid & 1u);
  float wt[9][8], bs[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    bs[c] = bias ? bias[grp * 8 + c] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t][c] = to_f(w[(grp * 8 + c) * 9 + t]);  // w: [16][3][3][1]
  }
  for (unsigned i = tid; i < total; i += stride) {
    unsigned q = i >> 1;  // (n, oh, strip)
    const int sidx = (int)(q % (unsigned)strips_w);
    q /= (unsigned)strips_w;
    const int oh = (int)(q % (unsigned)H), n = (int)(q / (unsigned)H);
    const int w0 = sidx * STRIP, w1 = min(W, w0 + STRIP);
    const T* xr1 = x + ((long long)n * H + oh) * W;
    const bool up = oh > 0, dn = oh + 1 < H;
    const T* xr0 = xr1 - W;
    const T* xr2 = xr1 + W;
    T* yp = y + (((long long)n * H + oh) * W + w0) * 16 + grp * 8;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1, b2 = 0.f;
    if (w0 > 0) {
      a1 = to_f(xr1[w0 - 1]);
      if (up) a0 = to_f(xr0[w0 - 1]);
      if (dn) a2 = to_f(xr2[w0 - 1]);
    }
    b1 = to_f(xr1[w0]);
    if (up) b0 = to_f(xr0[w0]);
    if (dn) b2 = to_f(xr2[w0]);
    for (int ow = w0; ow < w1; ++ow) {
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (ow + 1 < W) {
        c1 = to_f(xr1[ow + 1]);
        if (up) c0 = to_f(xr0[ow + 1]);
        if (dn) c2 = to_f(xr2[ow + 1]);
      }
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = bs[c];
        v = fmaf(a0, wt[0][c], v); v = fmaf(b0, wt[1][c], v); v = fmaf(c0, wt[2][c], v);
        v = fmaf(a1, wt[3][c], v); v = fmaf(b1, wt[4][c], v); v = fmaf(c1, wt[5][c], v);
        v = fmaf(a2, wt[6][c], v); v = fmaf(b2, wt[7][c], v); v = fmaf(c2, wt[8][c], v);
        acc[c] = relu ? fmaxf(v, 0.f) : v;
      }
      Store8<T>::st(yp, acc);
      yp += 16;
      a0 = b0; a1 = b1; a2 = b2;
      b0 = c0; b1 = c1; b2 = c2;
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
  if (W >= 64) {  // strip walk along the rows, weights in registers
    constexpr int STRIP = 32;
    long long sblocks = cdiv((long long)N * H * ((W + STRIP - 1) / STRIP) * 2, 256);
    if (sblocks > 148 * 8) sblocks = 148 * 8;
    OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)sblocks, 256, 0, st)(conv1_fwd_strip_kernel<T, STRIP>, (const T*)x, (const T*)w, bias, (T*)y, N,
                                                                                      H, W, relu)));
    OMR_LAUNCHED();
    return OMR_OK;
  }
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
  const size_t smem_s = sizeof(float) * 8 * groups * 72;
  if (sh == 1 && sw == 1 && W >= 64) {  // strip walk along the rows
    constexpr int STRIP = 32;
    const long long threads = (long long)N * H * ((W + STRIP - 1) / STRIP) * groups;
    long long blocks = cdiv(threads, 256);
    if (blocks > 148 * 2) blocks = 148 * 2;  // resident blocks only: every block ends with 9 * Co atomics onto the same addresses
    if (blocks < 1) blocks = 1;
    OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)blocks, 256, smem_s, st)(conv1_wgrad_strip_kernel<T, STRIP>, (const T*)x, (const T*)dy, dw, N,
                                                                                          H, W, Co)));
    OMR_LAUNCHED();
    return OMR_OK;
  }
  long long blocks = cdiv(total, 256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const size_t smem = sizeof(float) * 8 * groups * 72;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)blocks, 256, smem, st)(conv1_wgrad_kernel<T>, (const T*)x, (const T*)dy, dw, N, H, W, Co, Ho,
                                                                                    Wo, sh, sw)));
  OMR_LAUNCHED();
  return OMR_OK;
}
