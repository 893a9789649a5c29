// dwconv.cu -- depthwise 3x3 (stride 1, pad 1) on NHWC: forward, data gradient, weight gradient.
// HBM-bound: one thread per 4 channels, coalesced along C, taps held in registers.
#include "common.cuh"

namespace {

// FLIP == 0: y = dwconv(x, w) + bias ; FLIP == 1: dx = dwconv(dy, flipped w) (data gradient)
template <typename T, int FLIP>
__global__ void dwconv3x3_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                 T* __restrict__ y, int N, int H, int W, int C) {
  omr_pdl_enter();
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
  omr_pdl_enter();
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
  omr_pdl_enter();
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

// ---- column-strip kernels -------------------------------------------------------------------------------------------
// The depthwise layers run on the H/16 x W/8 feature maps (8 x 128, 13 x 101): few rows, wide.  A thread owns one
// (sample, column, group of 16 bytes of channels) and walks DOWN the column: every input row is loaded once (three
// 16-byte loads: left, centre, right) and scattered into the three output rows it feeds, whose accumulators rotate
// through registers.  3 loads per output instead of 9, weights packed in registers, no 64-bit index arithmetic.
template <typename T> struct R16;
template <> struct R16<bf16> {
  static constexpr int N = 8;
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { r = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
template <> struct R16<float> {
  static constexpr int N = 4;
  float4 r;
  __device__ __forceinline__ void load(const float* p) { r = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void zero() { r = make_float4(0, 0, 0, 0); }
  __device__ __forceinline__ void get(float (&v)[4]) const { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// FLIP == 0: y = dwconv(x, w) + bias ; FLIP == 1: dx = dwconv(dy, flipped w) (data gradient).  w: [3][3][C]
template <typename T, int FLIP>
__global__ void __launch_bounds__(256) dwconv3x3_col_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                            const float* __restrict__ bias, T* __restrict__ y, int N, int H,
                                                            int W, int C) {
  omr_pdl_enter();
  constexpr int VEC = R16<T>::N;
  const int cg = C / VEC;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= N * W * cg) return;
  const int c = (idx % cg) * VEC, ww = (idx / cg) % W, n = idx / (cg * W);
  R16<T> wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wt[t].load(w + (long long)(FLIP ? 8 - t : t) * C + c);
  float b[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) b[k] = (!FLIP && bias) ? bias[c + k] : 0.f;
  const long long rowpitch = (long long)W * C;
  const T* xp = x + (long long)n * H * rowpitch + (long long)ww * C + c;
  T* yp = y + (long long)n * H * rowpitch + (long long)ww * C + c;
  const bool hasl = ww > 0, hasr = ww + 1 < W;
  // acc[j]: output row (hin - 1 + j) while input row hin is being scattered (kh = 2 - j)
  float acc[3][VEC];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[j][k] = b[k];
  R16<T> cur[3], nxt[3];
  auto fetch = [&](R16<T>(&r)[3], int h) {
    const T* p = xp + (long long)h * rowpitch;
    if (hasl) r[0].load(p - C); else r[0].zero();
    r[1].load(p);
    if (hasr) r[2].load(p + C); else r[2].zero();
  };
  fetch(cur, 0);
  for (int hin = 0; hin < H; ++hin) {
    if (hin + 1 < H) fetch(nxt, hin + 1);
    float xv[3][VEC];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) cur[kw].get(xv[kw]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {  // output row hin - 1 + j reads this input row with kh = 2 - j
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float wv[VEC];
        wt[(2 - j) * 3 + kw].get(wv);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[j][k] = fmaf(xv[kw][k], wv[k], acc[j][k]);
      }
    }
    if (hin >= 1) R16<T>::store(yp + (long long)(hin - 1) * rowpitch, acc[0]);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { acc[0][k] = acc[1][k]; acc[1][k] = acc[2][k]; acc[2][k] = b[k]; }
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) cur[kw] = nxt[kw];
  }
  R16<T>::store(yp + (long long)(H - 1) * rowpitch, acc[0]);
}

// dw[c][tap] += sum_pix dy[pix][c] * x[pix @ tap][c] ; db[c] += sum dy.  Same column walk with 8-byte / 16-byte (4
// channel) groups: a three-row window of x in registers, 36 + 4 accumulators per thread, the threads of a block that
// share a channel group are combined in shared memory before ONE set of atomics per block.
template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_col_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                  float* __restrict__ dw, float* __restrict__ db, int N, int H,
                                                                  int W, int C) {
  omr_pdl_enter();
  extern __shared__ float red[];  // [256][41]
  const int cg = C / 4;
  float acc[9][4], accb[4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) accb[k] = 0.f;
  // grid-stride over the (sample, column, channel quad) groups: the stride is a multiple of the quads per row, so a thread
  // keeps its channel quad and its accumulators across groups -- a few hundred resident blocks end with ONE set of
  // atomics each (round 2: one block per 256 groups meant 512 blocks x 1280 atomics onto the same 1280 addresses, and
  // the same-address serialisation in L2 dominated the 37 us this kernel took for 17 MB of input)
  const int total = N * W * cg;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < total; idx += gridDim.x * 256) {
    const int c = (idx % cg) * 4, ww = (idx / cg) % W, n = idx / (cg * W);
    const long long rowpitch = (long long)W * C;
    const T* xp = x + (long long)n * H * rowpitch + (long long)ww * C + c;
    const T* gp = dy + (long long)n * H * rowpitch + (long long)ww * C + c;
    const bool hasl = ww > 0, hasr = ww + 1 < W;
    float xr[3][3][4];  // rows h-1, h, h+1 x (left, centre, right)
    auto fetch = [&](float (&r)[3][4], int h) {
      if (h < 0 || h >= H) {
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int k = 0; k < 4; ++k) r[q][k] = 0.f;
        return;
      }
      const T* p = xp + (long long)h * rowpitch;
      if (hasl) load4(p - C, r[0]); else { r[0][0] = r[0][1] = r[0][2] = r[0][3] = 0.f; }
      load4(p, r[1]);
      if (hasr) load4(p + C, r[2]); else { r[2][0] = r[2][1] = r[2][2] = r[2][3] = 0.f; }
    };
    // the loads of row h + 2 (x) and h + 1 (dy) are issued before the FMAs of row h: the walk is latency bound (13 rows,
    // four 8-byte loads each), so two rows' worth of loads are kept in flight per thread
    fetch(xr[0], -1);
    fetch(xr[1], 0);
    fetch(xr[2], 1);
    float g[4];
    load4(gp, g);
    for (int h = 0; h < H; ++h) {
      float xn[3][4], gn[4];
      fetch(xn, h + 2);
      if (h + 1 < H) load4(gp + (long long)(h + 1) * rowpitch, gn);
      else { gn[0] = gn[1] = gn[2] = gn[3] = 0.f; }
#pragma unroll
      for (int k = 0; k < 4; ++k) accb[k] += g[k];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[kh * 3 + kw][k] = fmaf(g[k], xr[kh][kw][k], acc[kh * 3 + kw][k]);
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) { xr[0][q][k] = xr[1][q][k]; xr[1][q][k] = xr[2][q][k]; xr[2][q][k] = xn[q][k]; }
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] = gn[k];
    }
  }
  // thread t's channel quad is (blockIdx.x * 256 + t) % cg: combine the threads of the block that share it
  float* mine = red + threadIdx.x * 41;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) mine[t * 4 + k] = acc[t][k];
#pragma unroll
  for (int k = 0; k < 4; ++k) mine[36 + k] = accb[k];
  __syncthreads();
  const int q0 = (int)(((long long)blockIdx.x * 256) % cg);  // channel quad of thread 0
  const int nq = cg < 256 ? cg : 256;                        // distinct quads in this block
  for (int i = threadIdx.x; i < nq * 40; i += 256) {
    const int t0 = i / 40, e = i - t0 * 40;  // t0: first thread holding this quad
    float sum = 0.f;
    for (int t = t0; t < 256; t += cg) sum += red[t * 41 + e];
    const int q = (q0 + t0) % cg;
    if (e < 36) atomicAdd(dw + (long long)(q * 4 + (e & 3)) * 9 + (e >> 2), sum);
    else if (db) atomicAdd(db + q * 4 + (e - 36), sum);
  }
}

// column kernels: 16-byte channel groups, a thread per (sample, column, group)
bool col_ok(int dt, const void* a, const void* w, const void* b, int N, int H, int W, int C) {
  const int vec = dt == OMR_BF16 ? 8 : 4;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return C % vec == 0 && al(a) && al(w) && al(b) && H >= 1 && H <= 64 && (long long)N * W * (C / vec) < (1LL << 31) &&
         (long long)N * W * (C / vec) >= 148LL * 128;
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
  if (col_ok(dt, x, w, y, N, H, W, C)) {
    const int vec = dt == OMR_BF16 ? 8 : 4;
    const int blocks = (int)cdiv((long long)N * W * (C / vec), 256);
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(blocks, 256, 0, as_stream(stream))(dwconv3x3_col_kernel<T, 0>, (const T*)x, (const T*)w, bias, (T*)y,
                                                                                            N, H, W, C)));
    OMR_LAUNCHED();
    return OMR_OK;
  }
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_cap(total), 256, 0, as_stream(stream))(dwconv3x3_kernel<T, 0>, 
                             (const T*)x, (const T*)w, bias, (T*)y, N, H, W, C)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_dwconv3x3_dgrad(int dt, const void* dy, const void* w, void* dx, int N, int H, int W, int C,
                                   omr_stream_t stream) {
  OMR_REQUIRE(C % 4 == 0, "omr_dwconv3x3_dgrad: C must be a multiple of 4 (got %d)", C);
  long long total = (long long)N * H * W * (C / 4);
  if (total <= 0) return OMR_OK;
  if (col_ok(dt, dy, w, dx, N, H, W, C)) {
    const int vec = dt == OMR_BF16 ? 8 : 4;
    const int blocks = (int)cdiv((long long)N * W * (C / vec), 256);
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(blocks, 256, 0, as_stream(stream))(dwconv3x3_col_kernel<T, 1>, (const T*)dy, (const T*)w, nullptr,
                                                                                            (T*)dx, N, H, W, C)));
    OMR_LAUNCHED();
    return OMR_OK;
  }
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_cap(total), 256, 0, as_stream(stream))(dwconv3x3_kernel<T, 1>, 
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
  // few rows, many columns (the encoder's depthwise layers): column walk, enough threads to fill the machine
  if (C % 4 == 0 && (256 % (C / 4) == 0 || (C / 4) % 256 == 0) && H <= 64 && (long long)N * W * (C / 4) >= 148LL * 256 &&
      (long long)N * W * (C / 4) < (1LL << 31) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    int blocks = (int)cdiv((long long)N * W * (C / 4), 256);
    if (blocks > 2 * 148 && C / 4 <= 256) blocks = 2 * 148;  // grid-stride kernel: two resident blocks per SM (256 % (C/4) == 0 keeps the quads fixed)
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(blocks, 256, sizeof(float) * 256 * 41, st)(dwconv3x3_wgrad_col_kernel<T>, (const T*)x, (const T*)dy, dw,
                                                                                                       db, N, H, W, C)));
    OMR_LAUNCHED();
    return OMR_OK;
  }
  if (C % 4 == 0 && C >= 16 && C <= 1024 && 256 % (C / 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    int per = (int)cdiv(P, 148LL * 4);
    if (per < 64) per = 64;
    const int blocks = (int)cdiv(P, per);
    const size_t smem = sizeof(float) * 256 * 40;
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(blocks, 256, smem, st)(dwconv3x3_wgrad_wide_kernel<T>, (const T*)x, (const T*)dy, dw, db, N, H, W, C,
                                                                                     per)));
    OMR_LAUNCHED();
    return OMR_OK;
  }
  int per = (int)cdiv(P, 148LL * 8);
  if (per < 16) per = 16;
  int blocks = (int)cdiv(P, per);
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(blocks, C, 0, st)(dwconv3x3_wgrad_kernel<T>, (const T*)x, (const T*)dy, dw, db, N, H, W,
                                                                          C, per)));
  OMR_LAUNCHED();
  return OMR_OK;
}
