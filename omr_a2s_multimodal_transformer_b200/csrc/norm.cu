// norm.cu -- InstanceNorm2d (NHWC, no affine) and residual-add + LayerNorm, forward and backward.
// All HBM-bound: coalesced along the channel axis, fp32 statistics, warp-shuffle reductions.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// InstanceNorm: per-(n,c) statistics over HW.  Phase 1 accumulates two sums per (n,c) with one
// atomicAdd pair per block; phase 2 turns them into (mean, rstd) / applies them.
// ---------------------------------------------------------------------------------------------
// MODE 0: (sum x, sum x^2).  MODE 1: (sum dy, sum dy * xhat) with xhat from saved stats.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) in_partial_kernel(const T* __restrict__ a, const T* __restrict__ xin,
                                                         const float* __restrict__ stats, double* __restrict__ out,
                                                         int HW, int C, int CT, int rows_per_block) {
  __shared__ float sm0[256], sm1[256];
  const int n = blockIdx.y;
  const int RS = 256 / CT;
  const int cl = threadIdx.x % CT, rl = threadIdx.x / CT;
  const int c = blockIdx.z * CT + cl;
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > HW) r1 = HW;
  float mean = 0.f, rstd = 0.f;
  if (MODE == 1) {
    mean = stats[((long long)n * C + c) * 2];
    rstd = stats[((long long)n * C + c) * 2 + 1];
  }
  float s0 = 0.f, s1 = 0.f;
  const long long base = (long long)n * HW;
  for (long long r = r0 + rl; r < r1; r += RS) {
    float v = to_f(a[(base + r) * C + c]);
    if (MODE == 0) {
      s0 += v; s1 = fmaf(v, v, s1);
    } else {
      float xh = (to_f(xin[(base + r) * C + c]) - mean) * rstd;
      s0 += v; s1 = fmaf(v, xh, s1);
    }
  }
  sm0[threadIdx.x] = s0; sm1[threadIdx.x] = s1;
  __syncthreads();
  if (rl == 0) {
    // block partials are combined in double: the variance E[x^2] - E[x]^2 and the backward's mean
    // subtractions are cancellation-prone, and fp32 parity (1e-4 on gradients) needs them clean
    double d0 = s0, d1 = s1;
    for (int k = 1; k < RS; ++k) { d0 += (double)sm0[k * CT + cl]; d1 += (double)sm1[k * CT + cl]; }
    atomicAdd(out + ((long long)n * C + c) * 2, d0);
    atomicAdd(out + ((long long)n * C + c) * 2 + 1, d1);
  }
}

__global__ void in_finalize_kernel(const double* __restrict__ sums, float* __restrict__ stats, long long NC,
                                   double inv_hw, double eps) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NC) {
    double mean = sums[i * 2] * inv_hw;
    double var = sums[i * 2 + 1] * inv_hw - mean * mean;
    var = var > 0.0 ? var : 0.0;
    stats[i * 2] = (float)mean;
    stats[i * 2 + 1] = (float)(1.0 / sqrt(var + eps));
  }
}

template <typename T>
__global__ void in_apply_fwd_kernel(const T* __restrict__ x, const float* __restrict__ stats, T* __restrict__ y, int N,
                                    int HW, int C) {
  const int c4n = C / 4;
  long long total = (long long)N * HW * c4n;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c4 = (int)(i % c4n);
    long long p = i / c4n;
    int n = (int)(p / HW);
    float v[4];
    load4(x + p * C + c4 * 4, v);
    const float* s = stats + ((long long)n * C + c4 * 4) * 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (v[k] - s[2 * k]) * s[2 * k + 1];
    store4(y + p * C + c4 * 4, v);
  }
}

template <typename T>
__global__ void in_apply_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ stats,
                                    const double* __restrict__ sums, T* __restrict__ dx, int N, int HW, int C,
                                    double inv_hw, int relu_mask, float mask_scale) {
  const int c4n = C / 4;
  long long total = (long long)N * HW * c4n;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c4 = (int)(i % c4n);
    long long p = i / c4n;
    int n = (int)(p / HW);
    float g[4], v[4];
    load4(dy + p * C + c4 * 4, g);
    load4(x + p * C + c4 * 4, v);
    const float* s = stats + ((long long)n * C + c4 * 4) * 2;
    const double* q = sums + ((long long)n * C + c4 * 4) * 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float rstd = s[2 * k + 1];
      float xh = (v[k] - s[2 * k]) * rstd;
      const float xin = v[k];
      v[k] = rstd * (g[k] - (float)(q[2 * k] * inv_hw) - xh * (float)(q[2 * k + 1] * inv_hw));
      if (relu_mask) v[k] = xin > 0.f ? v[k] * mask_scale : 0.f;
    }
    store4(dx + p * C + c4 * 4, v);
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

extern "C" int omr_instnorm_fwd(int dt, const void* x, void* y, float* stats, double* ws, int N, int HW, int C,
                                float eps, omr_stream_t stream) {
  int CT = pick_ct(C);
  OMR_REQUIRE(CT > 0 && C % 4 == 0, "omr_instnorm_fwd: unsupported channel count %d", C);
  if ((long long)N * HW * C <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  OMR_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)N * C * 2, st));
  int RS = 256 / CT;
  int rpb = (int)cdiv(HW, cdiv(148LL * 8, (long long)N * (C / CT)));
  if (rpb < RS * 8) rpb = RS * 8;
  dim3 grid((unsigned)cdiv(HW, rpb), (unsigned)N, (unsigned)(C / CT));
  OMR_DISPATCH_DT(dt, T, (in_partial_kernel<T, 0><<<grid, 256, 0, st>>>((const T*)x, nullptr, nullptr, ws, HW, C, CT,
                                                                        rpb)));
  OMR_LAUNCHED();
  in_finalize_kernel<<<(int)cdiv((long long)N * C, 256), 256, 0, st>>>(ws, stats, (long long)N * C, 1.0 / HW,
                                                                                (double)eps);
  OMR_LAUNCHED();
  long long total = (long long)N * HW * (C / 4);
  OMR_DISPATCH_DT(dt, T, (in_apply_fwd_kernel<T><<<grid_cap(total), 256, 0, st>>>((const T*)x, stats, (T*)y, N, HW, C)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_instnorm_bwd(int dt, const void* dy, const void* x, const float* stats, void* dx, double* ws, int N,
                                int HW, int C, int relu_mask, float mask_scale, omr_stream_t stream) {
  int CT = pick_ct(C);
  OMR_REQUIRE(CT > 0 && C % 4 == 0, "omr_instnorm_bwd: unsupported channel count %d", C);
  if ((long long)N * HW * C <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  OMR_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)N * C * 2, st));
  int RS = 256 / CT;
  int rpb = (int)cdiv(HW, cdiv(148LL * 8, (long long)N * (C / CT)));
  if (rpb < RS * 8) rpb = RS * 8;
  dim3 grid((unsigned)cdiv(HW, rpb), (unsigned)N, (unsigned)(C / CT));
  OMR_DISPATCH_DT(dt, T, (in_partial_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)dy, (const T*)x, stats, ws, HW, C, CT,
                                                                        rpb)));
  OMR_LAUNCHED();
  long long total = (long long)N * HW * (C / 4);
  OMR_DISPATCH_DT(dt, T, (in_apply_bwd_kernel<T><<<grid_cap(total), 256, 0, st>>>((const T*)dy, (const T*)x, stats, ws,
                                                                                 (T*)dx, N, HW, C, 1.0 / HW, relu_mask, mask_scale)));
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
  OMR_DISPATCH_DT(dt, T, LN_SWITCH(D, (add_ln_fwd_kernel<T, VPL><<<blocks, 256, 0, st>>>(
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
  OMR_DISPATCH_DT(dt, T, LN_SWITCH(D, (ln_bwd_kernel<T, VPL><<<(int)blocks, 256, 0, st>>>(
                                          (const T*)dy, (const T*)s, stats, gamma, (T*)ds, dgamma, dbeta, rows))));
  OMR_LAUNCHED();
  return OMR_OK;
}
