// norm.cu -- InstanceNorm2d (NHWC, no affine) and residual-add + LayerNorm, forward and backward.
// All HBM-bound: coalesced along the channel axis, fp32 statistics, warp-shuffle reductions.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// InstanceNorm: per-(n,c) statistics over HW.  Phase 1 accumulates two sums per (n,c) with one
// atomicAdd pair per block; phase 2 turns them into (mean, rstd) / applies them.
// Both phases see a sample as a flat array of HW*C elements walked with 16-byte accesses; the grid
// stride is a multiple of C, so a thread keeps ONE group of 16/sizeof(T) channels for its whole life
// (statistics in registers, no index arithmetic in the loop) and several loads are in flight per thread.
// ---------------------------------------------------------------------------------------------
template <typename T> struct V16;
template <> struct V16<bf16> {
  static constexpr int N = 8;
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw load_raw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
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
template <> struct V16<float> {
  static constexpr int N = 4;
  typedef float4 Raw;
  static __device__ __forceinline__ Raw load_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

constexpr int IN_UNROLL = 4;

// MODE 0: (sum x, sum x^2).  MODE 1: (sum dy, sum dy * xhat) with xhat from saved stats.  MODE 2: (sum dy, sum dy * x)
// -- the "raw" backward sums that the convolution data-gradient epilogue also produces (conv_tc.cu MODE 3).
// grid (blocks per sample, N); requires (256 * VEC) % C == 0.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) in_partial_kernel(const T* __restrict__ a, const T* __restrict__ xin,
                                                         const float* __restrict__ stats, double* __restrict__ out,
                                                         long long per_sample, int C) {
  omr_pdl_enter();
  constexpr int VEC = V16<T>::N;
  __shared__ float sm0[256 * VEC], sm1[256 * VEC];
  const int n = blockIdx.y;
  const long long step = (long long)gridDim.x * 256 * VEC;
  long long e = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC;
  const int c0 = (int)(e % C);
  const T* ap = a + (long long)n * per_sample;
  const T* xp = MODE >= 1 ? xin + (long long)n * per_sample : nullptr;
  float mean[VEC], rstd[VEC], s0[VEC], s1[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    s0[k] = 0.f; s1[k] = 0.f;
    if (MODE == 1) {
      mean[k] = stats[((long long)n * C + c0 + k) * 2];
      rstd[k] = stats[((long long)n * C + c0 + k) * 2 + 1];
    }
  }
  for (; e < per_sample; e += step * IN_UNROLL) {
    float v[IN_UNROLL][VEC], x[IN_UNROLL][VEC];
#pragma unroll
    for (int u = 0; u < IN_UNROLL; ++u) {
      const long long eu = e + u * step;
      if (eu < per_sample) {
        V16<T>::load(ap + eu, v[u]);
        if (MODE >= 1) V16<T>::load(xp + eu, x[u]);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) { v[u][k] = 0.f; if (MODE == 1) x[u][k] = mean[k]; if (MODE == 2) x[u][k] = 0.f; }
      }
    }
#pragma unroll
    for (int u = 0; u < IN_UNROLL; ++u)
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        s0[k] += v[u][k];
        if (MODE == 0) s1[k] = fmaf(v[u][k], v[u][k], s1[k]);
        else if (MODE == 1) s1[k] = fmaf(v[u][k], (x[u][k] - mean[k]) * rstd[k], s1[k]);
        else s1[k] = fmaf(v[u][k], x[u][k], s1[k]);
      }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) { sm0[threadIdx.x * VEC + k] = s0[k]; sm1[threadIdx.x * VEC + k] = s1[k]; }
  __syncthreads();
  // thread t's channel group starts at (t * VEC) % C: channel c lives in the slots c, c + C, c + 2C, ... of the block's
  // 256 * VEC values.  Block partials are combined in double: the variance E[x^2] - E[x]^2 and the backward's mean
  // subtractions are cancellation-prone, and fp32 parity (1e-4 on gradients) needs them clean.
  // (blockIdx.x * 256 * VEC is a multiple of C, so the block offset does not shift the mapping.)
  for (int c = threadIdx.x; c < C; c += 256) {
    double d0 = 0.0, d1 = 0.0;
    for (int i = c; i < 256 * VEC; i += C) { d0 += (double)sm0[i]; d1 += (double)sm1[i]; }
    atomicAdd(out + ((long long)n * C + c) * 2, d0);
    atomicAdd(out + ((long long)n * C + c) * 2 + 1, d1);
  }
}

__global__ void in_finalize_kernel(const double* __restrict__ sums, float* __restrict__ stats, long long NC,
                                   double inv_hw, double eps) {
  omr_pdl_enter();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NC) {
    double mean = sums[i * 2] * inv_hw;
    double var = sums[i * 2 + 1] * inv_hw - mean * mean;
    var = var > 0.0 ? var : 0.0;
    stats[i * 2] = (float)mean;
    stats[i * 2 + 1] = (float)(1.0 / sqrt(var + eps));
  }
}

// grid (blocks per sample, N); requires (256 * VEC) % C == 0
template <typename T, int U, int MB>
__global__ void __launch_bounds__(256, MB) in_apply_fwd_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                                               T* __restrict__ y, long long per_sample, int C) {
  omr_pdl_enter();
  constexpr int VEC = V16<T>::N;
  const int n = blockIdx.y;
  const long long step = (long long)gridDim.x * 256 * VEC;
  long long e = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC;
  const int c0 = (int)(e % C);
  const T* xp = x + (long long)n * per_sample;
  T* yp = y + (long long)n * per_sample;
  float mean[VEC], rstd[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    mean[k] = stats[((long long)n * C + c0 + k) * 2];
    rstd[k] = stats[((long long)n * C + c0 + k) * 2 + 1];
  }
  for (; e < per_sample; e += step * U) {
    float v[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * step < per_sample) V16<T>::load(xp + e + u * step, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * step < per_sample) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[u][k] = (v[u][k] - mean[k]) * rstd[k];
        V16<T>::store(yp + e + u * step, v[u]);
      }
  }
}

// dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)) = A * dy + B * x + D per channel: three coefficients in registers
// and the loads kept as raw 16-byte words until they are used (round 2: the first version held four statistics per
// channel and unpacked 2 x U x 8 floats up front -- 160 registers, ONE block of 8 warps per SM, 3.6 TB/s; this one fits
// three blocks per SM).
template <typename T, int U, int MB>
__global__ void __launch_bounds__(256, MB) in_apply_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                               const float* __restrict__ stats, const double* __restrict__ sums,
                                                               T* __restrict__ dx, long long per_sample, int C, double inv_hw,
                                                               int relu_mask, float mask_scale, int raw_sums, float* __restrict__ colsum) {
  omr_pdl_enter();
  constexpr int VEC = V16<T>::N;
  typedef typename V16<T>::Raw Raw;
  __shared__ float smc[256 * VEC];
  const int n = blockIdx.y;
  const long long step = (long long)gridDim.x * 256 * VEC;
  long long e = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC;
  const int c0 = (int)(e % C);
  const T* gp = dy + (long long)n * per_sample;
  const T* xp = x + (long long)n * per_sample;
  T* op = dx + (long long)n * per_sample;
  float ca[VEC], cb[VEC], cd[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const long long i = (long long)n * C + c0 + k;
    const float mean = stats[i * 2], rstd = stats[i * 2 + 1];
    const float m1 = (float)(sums[i * 2] * inv_hw);
    // raw_sums: sums = (sum dy, sum dy * x) -> sum dy * xhat = rstd * (sum dy * x - mean * sum dy), combined in double
    const float m2 = raw_sums ? (float)((double)rstd * (sums[i * 2 + 1] - (double)mean * sums[i * 2]) * inv_hw)
                              : (float)(sums[i * 2 + 1] * inv_hw);
    ca[k] = rstd;
    cb[k] = -rstd * rstd * m2;
    cd[k] = -rstd * m1 - cb[k] * mean;
  }
  float cs[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) cs[k] = 0.f;
  for (; e < per_sample; e += step * U) {
    Raw gr[U], xr[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * step < per_sample) {
        gr[u] = V16<T>::load_raw(gp + e + u * step);
        xr[u] = V16<T>::load_raw(xp + e + u * step);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * step < per_sample) {
        float g[VEC], v[VEC];
        V16<T>::unpack(gr[u], g);
        V16<T>::unpack(xr[u], v);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float xin = v[k];
          float r = fmaf(ca[k], g[k], fmaf(cb[k], xin, cd[k]));
          if (relu_mask) r = xin > 0.f ? r * mask_scale : 0.f;
          v[k] = r;
          if (colsum) cs[k] += round_to<T>(r);
        }
        V16<T>::store(op + e + u * step, v);
      }
  }
  if (colsum) {  // column sums of the stored dx = bias gradient of the convolution in front of the ReLU / norm
#pragma unroll
    for (int k = 0; k < VEC; ++k) smc[threadIdx.x * VEC + k] = cs[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float d0 = 0.f;
      for (int i = c; i < 256 * VEC; i += C) d0 += smc[i];
      atomicAdd(colsum + c, d0);
    }
  }
}

int pick_ct(int C) {
  if (C >= 256) return (C % 256 == 0) ? 256 : -1;
  return (256 % C == 0) ? C : -1;
}
int grid_cap(long long n) {
  long long b = cdiv(n, 256);
  if (b < 1) b = 1;
  if (b > 148LL * 32) b = 148LL * 32;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, VPL = D / 32 values per lane.
// ---------------------------------------------------------------------------------------------
template <typename T, int VPL>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, T* __restrict__ s_out,
                                                         T* __restrict__ y, float* __restrict__ stats, long long rows,
                                                         float eps) {
  omr_pdl_enter();
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    int d = k * 32 + lane;
    float a = to_f(x[row * D + d]);
    if (res) a += to_f(res[row * D + d]);
    // statistics and output come from the fp32 sum (one bf16 rounding less per block on the residual stream);
    // the stored copy of s is rounded, which perturbs the backward's xhat by 2^-9 relative at most
    v[k] = a;
    sum += a;
  }
  float mean = warp_sum(sum) * (1.f / D);
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    float d = v[k] - mean;
    var = fmaf(d, d, var);
  }
  float rstd = rsqrtf(warp_sum(var) * (1.f / D) + eps);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    int d = k * 32 + lane;
    if (s_out) s_out[row * D + d] = from_f<T>(v[k]);
    y[row * D + d] = from_f<T>((v[k] - mean) * rstd * gamma[d] + beta[d]);
  }
  if (stats && lane == 0) {
    stats[row * 2] = mean;
    stats[row * 2 + 1] = rstd;
  }
}

template <typename T, int VPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ s,
                                                     const float* __restrict__ stats, const float* __restrict__ gamma,
                                                     T* __restrict__ ds, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, long long rows) {
  omr_pdl_enter();
  constexpr int D = VPL * 32;
  __shared__ float sg[8][D + 1];
  __shared__ float sb[8][D + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float gsum[VPL], bsum[VPL], gam[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    gsum[k] = 0.f; bsum[k] = 0.f; gam[k] = gamma[k * 32 + lane];
  }
  for (long long row = (long long)blockIdx.x * nw + wid; row < rows; row += (long long)gridDim.x * nw) {
    float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float xh[VPL], g[VPL];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int d = k * 32 + lane;
      float gy = to_f(dy[row * D + d]);
      xh[k] = (to_f(s[row * D + d]) - mean) * rstd;
      gsum[k] = fmaf(gy, xh[k], gsum[k]);
      bsum[k] += gy;
      g[k] = gy * gam[k];
      m1 += g[k];
      m2 = fmaf(g[k], xh[k], m2);
    }
    m1 = warp_sum(m1) * (1.f / D);
    m2 = warp_sum(m2) * (1.f / D);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int d = k * 32 + lane;
      ds[row * D + d] = from_f<T>(rstd * (g[k] - m1 - xh[k] * m2));
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    sg[wid][k * 32 + lane] = gsum[k];
    sb[wid][k * 32 + lane] = bsum[k];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < nw; ++w) { a += sg[w][d]; b += sb[w][d]; }
    atomicAdd(dgamma + d, a);
    atomicAdd(dbeta + d, b);
  }
}

}  // namespace

// Unroll / occupancy of the apply kernels, measured on B200 at 32 x 195 x 808 x 32 bf16 (323 MB; scripts/bench_stream.py): forward
// unroll 4 (78 registers, three blocks per SM) 111 us = 5.8 TB/s; backward unroll 2 with >= 4 blocks per SM 161 us = 6.0 TB/s
// (unroll 4: 177 us; the round-1 kernel with 160 registers and one block per SM: 263 us).
constexpr int IN_FWD_U = 4, IN_FWD_MB = 1, IN_BWD_U = 2, IN_BWD_MB = 4;

// blocks per sample for the flat InstanceNorm kernels: ~8 CTAs per SM over the batch, at least IN_UNROLL loads per thread
static int in_blocks(long long per_sample, int vec, int N) {
  long long want = cdiv(148LL * 8, N);
  long long most = cdiv(per_sample, 256LL * vec * IN_UNROLL);
  if (want > most) want = most;
  return (int)(want < 1 ? 1 : want);
}

// (sum a, sum a * b-ish) per (n, c) into `out` (zeroed here): mode 0 / 2 of in_partial_kernel, for the callers in dispatch.cu
int omr_in_partial_sums(int dt, int mode, const void* a, const void* xin, double* out, int N, int HW, int C, cudaStream_t st) {
  const int vec = dt == OMR_BF16 ? 8 : 4;
  OMR_REQUIRE(C >= vec && C % vec == 0 && (256 * vec) % C == 0, "instnorm sums: unsupported channel count %d", C);
  OMR_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)N * C * 2, st));
  if ((long long)N * HW * C <= 0) return OMR_OK;
  const long long per = (long long)HW * C;
  dim3 grid((unsigned)in_blocks(per, vec, N), (unsigned)N);
  if (mode == 0) {
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, 256, 0, st)(in_partial_kernel<T, 0>, (const T*)a, nullptr, nullptr, out, per, C)));
  } else {
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, 256, 0, st)(in_partial_kernel<T, 2>, (const T*)a, (const T*)xin, nullptr, out, per, C)));
  }
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_instnorm_fwd(int dt, const void* x, void* y, float* stats, double* ws, int N, int HW, int C,
                                float eps, int sums_ready, omr_stream_t stream) {
  const int vec = dt == OMR_BF16 ? 8 : 4;
  OMR_REQUIRE(C >= vec && C % vec == 0 && (256 * vec) % C == 0, "omr_instnorm_fwd: unsupported channel count %d", C);
  OMR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
              "omr_instnorm_fwd: x and y must be 16-byte aligned");
  if ((long long)N * HW * C <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  const long long per = (long long)HW * C;
  dim3 grid((unsigned)in_blocks(per, vec, N), (unsigned)N);
  if (!sums_ready) {  // otherwise ws already holds (sum x, sum x^2): omr_conv3x3_fwd accumulated them in its epilogue
    OMR_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)N * C * 2, st));
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, 256, 0, st)(in_partial_kernel<T, 0>, (const T*)x, nullptr, nullptr, ws, per, C)));
    OMR_LAUNCHED();
  }
  OmrLaunch((int)cdiv((long long)N * C, 256), 256, 0, st)(in_finalize_kernel, ws, stats, (long long)N * C, 1.0 / HW,
                                                                                (double)eps);
  OMR_LAUNCHED();
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, 256, 0, st)(in_apply_fwd_kernel<T, IN_FWD_U, IN_FWD_MB>, (const T*)x, stats, (T*)y, per, C)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_instnorm_bwd(int dt, const void* dy, const void* x, const float* stats, void* dx, double* ws, int N,
                                int HW, int C, int relu_mask, float mask_scale, int sums_ready, float* colsum,
                                omr_stream_t stream) {
  const int vec = dt == OMR_BF16 ? 8 : 4;
  OMR_REQUIRE(C >= vec && C % vec == 0 && (256 * vec) % C == 0, "omr_instnorm_bwd: unsupported channel count %d", C);
  OMR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dx) & 15) == 0,
              "omr_instnorm_bwd: dy, x and dx must be 16-byte aligned");
  if ((long long)N * HW * C <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  const long long per = (long long)HW * C;
  dim3 grid((unsigned)in_blocks(per, vec, N), (unsigned)N);
  if (!sums_ready) {  // otherwise ws holds the RAW sums (sum dy, sum dy * x) from omr_conv3x3_dgrad's epilogue
    OMR_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)N * C * 2, st));
    OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, 256, 0, st)(in_partial_kernel<T, 1>, (const T*)dy, (const T*)x, stats, ws, per, C)));
    OMR_LAUNCHED();
  }
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, 256, 0, st)(in_apply_bwd_kernel<T, IN_BWD_U, IN_BWD_MB>, (const T*)dy, (const T*)x, stats, ws,
                                                                      (T*)dx, per, C, 1.0 / HW, relu_mask, mask_scale, sums_ready ? 1 : 0,
                                                                      colsum)));
  OMR_LAUNCHED();
  return OMR_OK;
}

#define LN_SWITCH(D, CALL)                                     \
  switch ((D) / 32) {                                          \
    case 1: { constexpr int VPL = 1; CALL; } break;            \
    case 2: { constexpr int VPL = 2; CALL; } break;            \
    case 4: { constexpr int VPL = 4; CALL; } break;            \
    case 8: { constexpr int VPL = 8; CALL; } break;            \
    case 16: { constexpr int VPL = 16; CALL; } break;          \
    default:                                                   \
      omr_set_error("layernorm: unsupported width %d (32,64,128,256,512)", (int)(D)); \
      return OMR_ERR_INVALID;                                  \
  }

extern "C" int omr_add_layernorm_fwd(int dt, const void* x, const void* res, const float* gamma, const float* beta,
                                     void* s_out, void* y, float* stats, long long rows, int D, float eps,
                                     omr_stream_t stream) {
  OMR_REQUIRE(D % 32 == 0, "omr_add_layernorm_fwd: D must be a multiple of 32");
  if (rows <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  int blocks = (int)cdiv(rows, 8);
  OMR_DISPATCH_DT(dt, T, LN_SWITCH(D, (OmrLaunch(blocks, 256, 0, st)(add_ln_fwd_kernel<T, VPL>, 
                                          (const T*)x, (const T*)res, gamma, beta, (T*)s_out, (T*)y, stats, rows, eps))));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_layernorm_bwd(int dt, const void* dy, const void* s, const float* stats, const float* gamma,
                                 void* ds, float* dgamma, float* dbeta, long long rows, int D, omr_stream_t stream) {
  OMR_REQUIRE(D % 32 == 0, "omr_layernorm_bwd: D must be a multiple of 32");
  if (rows <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  long long blocks = cdiv(rows, 8 * 4);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  OMR_DISPATCH_DT(dt, T, LN_SWITCH(D, (OmrLaunch((int)blocks, 256, 0, st)(ln_bwd_kernel<T, VPL>, 
                                          (const T*)dy, (const T*)s, stats, gamma, (T*)ds, dgamma, dbeta, rows))));
  OMR_LAUNCHED();
  return OMR_OK;
}
