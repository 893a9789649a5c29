// dwconv.cu -- depthwise 3x3 (stride 1, pad 1) on NHWC: forward, data gradient, weight gradient.
// HBM-bound: one thread per 4 channels, coalesced along C, taps held in registers.
#include "common.cuh"

namespace {

// FLIP == 0: y = dwconv(x, w) + bias ; FLIP == 1: dx = dwconv(dy, flipped w) (data gradient)
template <typename T, int FLIP>
__global__ void dwconv3x3_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                 T* __restrict__ y, int N, int H, int W, int C) {
  const int c4n = C / 4;
  long long total = (long long)N * H * W * c4n;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c4 = (int)(i % c4n);
    long long p = i / c4n;
    int ww = (int)(p % W);
    long long r = p / W;
    int hh = (int)(r % H);
    int n = (int)(r / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (!FLIP && bias) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = bias[c4 * 4 + k];
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int hs = hh + kh - 1;
      if (hs < 0 || hs >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int ws = ww + kw - 1;
        if (ws < 0 || ws >= W) continue;
        int tap = FLIP ? (2 - kh) * 3 + (2 - kw) : kh * 3 + kw;
        float xv[4], wv[4];
        load4(x + (((long long)n * H + hs) * W + ws) * C + c4 * 4, xv);
        load4(w + (long long)tap * C + c4 * 4, wv);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(xv[k], wv[k], acc[k]);
      }
    }
    store4(y + p * C + c4 * 4, acc);
  }
}

// dw[c][tap] += sum_pix dy[pix][c] * x[pix @ tap][c] ; db[c] += sum dy.  One thread per channel,
// blockDim.x = C (<= 1024), each block reduces a contiguous chunk of pixels.
template <typename T>
__global__ void dwconv3x3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                                       float* __restrict__ db, int N, int H, int W, int C, int pix_per_block) {
  const int c = threadIdx.x;
  long long P = (long long)N * H * W;
  long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > P) p1 = P;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  float accb = 0.f;
  for (long long p = p0; p < p1; ++p) {
    int ww = (int)(p % W);
    long long r = p / W;
    int hh = (int)(r % H);
    int n = (int)(r / H);
    float g = to_f(dy[p * C + c]);
    accb += g;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int hs = hh + kh - 1;
      if (hs < 0 || hs >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int ws = ww + kw - 1;
        if (ws < 0 || ws >= W) continue;
        acc[kh * 3 + kw] = fmaf(g, to_f(x[(((long long)n * H + hs) * W + ws) * C + c]), acc[kh * 3 + kw]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) atomicAdd(dw + (long long)c * 9 + t, acc[t]);
  if (db) atomicAdd(db + c, accb);
}

// Wide variant: 256 threads = (C/4 channel quads) x (256/(C/4) pixel lanes); every thread owns 4 channels (one 8/16-byte
// load per tap), walks its pixel lane of the block's pixel range, and the lanes are combined in shared memory before
// ONE set of atomics per block -- many loads in flight per thread group instead of a serial per-channel loop.
template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_wide_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                   float* __restrict__ dw, float* __restrict__ db, int N, int H,
                                                                   int W, int C, int pix_per_block) {
  extern __shared__ float red[];  // [lanes][C/4][40]
  const int quads = C / 4, lanes = 256 / quads;
  const int cq = threadIdx.x % quads, lane = threadIdx.x / quads;
  const long long P = (long long)N * H * W;
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > P) p1 = P;
  float acc[9][4], accb[4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) accb[k] = 0.f;
  for (long long p = p0 + lane; p < p1; p += lanes) {
    const int ww = (int)(p % W);
    const long long r = p / W;
    const int hh = (int)(r % H), n = (int)(r / H);
    float g[4];
    load4(dy + p * C + cq * 4, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) accb[k] += g[k];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hs = hh + kh - 1;
      if (hs < 0 || hs >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ws = ww + kw - 1;
        if (ws < 0 || ws >= W) continue;
        float xv[4];
        load4(x + (((long long)n * H + hs) * W + ws) * C + cq * 4, xv);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[kh * 3 + kw][k] = fmaf(g[k], xv[k], acc[kh * 3 + kw][k]);
      }
    }
  }
  float* mine = red + ((long long)lane * quads + cq) * 40;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) mine[t * 4 + k] = acc[t][k];
#pragma unroll
  for (int k = 0; k < 4; ++k) mine[36 + k] = accb[k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < quads * 40; idx += 256) {
    float sum = 0.f;
    for (int l = 0; l < lanes; ++l) sum += red[(long long)l * quads * 40 + idx];
    const int q = idx / 40, e = idx - q * 40;
    if (e < 36) atomicAdd(dw + (long long)(q * 4 + (e & 3)) * 9 + (e >> 2), sum);
    else if (db) atomicAdd(db + q * 4 + (e - 36), sum);
  }
}

int grid_cap(long long n) {
  long long b = cdiv(n, 256);
  if (b < 1) b = 1;
  if (b > 148LL * 32) b = 148LL * 32;
  return (int)b;
}

}  // namespace

extern "C" int omr_dwconv3x3_fwd(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W,
                                 int C, omr_stream_t stream) {
  OMR_REQUIRE(C % 4 == 0, "omr_dwconv3x3_fwd: C must be a multiple of 4 (got %d)", C);
  long long total = (long long)N * H * W * (C / 4);
  if (total <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (dwconv3x3_kernel<T, 0><<<grid_cap(total), 256, 0, as_stream(stream)>>>(
                             (const T*)x, (const T*)w, bias, (T*)y, N, H, W, C)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_dwconv3x3_dgrad(int dt, const void* dy, const void* w, void* dx, int N, int H, int W, int C,
                                   omr_stream_t stream) {
  OMR_REQUIRE(C % 4 == 0, "omr_dwconv3x3_dgrad: C must be a multiple of 4 (got %d)", C);
  long long total = (long long)N * H * W * (C / 4);
  if (total <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (dwconv3x3_kernel<T, 1><<<grid_cap(total), 256, 0, as_stream(stream)>>>(
                             (const T*)dy, (const T*)w, nullptr, (T*)dx, N, H, W, C)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_dwconv3x3_wgrad(int dt, const void* x, const void* dy, float* dw, float* db, int N, int H, int W,
                                   int C, int accumulate, omr_stream_t stream) {
  OMR_REQUIRE(C >= 1 && C <= 1024, "omr_dwconv3x3_wgrad: C out of range (%d)", C);
  cudaStream_t st = as_stream(stream);
  if (!accumulate) {
    OMR_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 9, st));
    if (db) OMR_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)C, st));
  }
  long long P = (long long)N * H * W;
  if (P <= 0) return OMR_OK;
  if (C % 4 == 0 && C >= 16 && C <= 1024 && 256 % (C / 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    int per = (int)cdiv(P, 148LL * 4);
    if (per < 64) per = 64;
    const int blocks = (int)cdiv(P, per);
    const size_t smem = sizeof(float) * 256 * 40;
    OMR_DISPATCH_DT(dt, T, (dwconv3x3_wgrad_wide_kernel<T><<<blocks, 256, smem, st>>>((const T*)x, (const T*)dy, dw, db, N, H, W, C,
                                                                                     per)));
    OMR_LAUNCHED();
    return OMR_OK;
  }
  int per = (int)cdiv(P, 148LL * 8);
  if (per < 16) per = 16;
  int blocks = (int)cdiv(P, per);
  OMR_DISPATCH_DT(dt, T, (dwconv3x3_wgrad_kernel<T><<<blocks, C, 0, st>>>((const T*)x, (const T*)dy, dw, db, N, H, W,
                                                                          C, per)));
  OMR_LAUNCHED();
  return OMR_OK;
}
