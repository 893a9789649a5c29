// conv_simt.cu -- CUDA-core implicit-GEMM 3x3 convolution on NHWC (forward, data gradient, weight
// gradient), exact fp32 accumulation.  fp32-mode path and validation reference for the tcgen05
// implicit-GEMM kernels.
//
// Implicit GEMM view: M = destination pixels, N = destination channels,
// K = 9 * source channels, A gathered on the fly from the NHWC source with zero padding.
#include "simt_tile.cuh"

namespace {

struct ConvArgs {
  const void* src; const void* w; const float* bias; void* dst;
  int N, Hs, Ws, Cs;  // source tensor [N,Hs,Ws,Cs]
  int Hd, Wd, Cd;     // destination tensor [N,Hd,Wd,Cd]
  int sh, sw, relu;
};

// MODE 0: forward (src = x, dst = y).  MODE 1: data gradient (src = dy, dst = dx, w = [Ci,3,3,Co]).
template <typename T, int TX, int MODE>
__global__ void __launch_bounds__(256) conv3x3_igemm_kernel(ConvArgs a) {
  omr_pdl_enter();
  constexpr int BN_ = 4 * TX, BM_ = 4 * (256 / TX);
  constexpr int LDA = BM_ + 4, LDB = BN_ + 4;
  constexpr int A_PER = BM_ / 16, B_PER = (BN_ * 16 + 255) / 256;
  __shared__ __align__(16) float As[SIMT_BK * LDA];
  __shared__ __align__(16) float Bs[SIMT_BK * LDB];
  __shared__ int pix_n[BM_], pix_h[BM_], pix_w[BM_];

  const int tid = threadIdx.x;
  const int ty = tid / TX, tx = tid % TX;
  const long long Mtot = (long long)a.N * a.Hd * a.Wd;
  const long long m_blk = (long long)blockIdx.x * BM_;
  const int n_blk = blockIdx.y * BN_;
  const int K = 9 * a.Cs;
  const T* src = (const T*)a.src;
  const T* w = (const T*)a.w;

  for (int i = tid; i < BM_; i += 256) {
    long long m = m_blk + i;
    if (m < Mtot) {
      int wd = (int)(m % a.Wd);
      long long r = m / a.Wd;
      pix_w[i] = wd; pix_h[i] = (int)(r % a.Hd); pix_n[i] = (int)(r / a.Hd);
    } else {
      pix_n[i] = -1; pix_h[i] = 0; pix_w[i] = 0;
    }
  }
  __syncthreads();

  float ra[A_PER], rb[B_PER];
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  const int kl = tid & 15;  // this thread's k lane inside every k-tile (constant)
  auto load_tiles = [&](int k0) {
    int kg = k0 + kl;
    bool kvalid = kg < K;
    int tap = kvalid ? kg / a.Cs : 0;
    int c = kg - tap * a.Cs;
    int kh = tap / 3, kw = tap - kh * 3;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int m = (tid >> 4) + i * 16;
      int n = pix_n[m];
      float v = 0.f;
      if (kvalid && n >= 0) {
        int hs, ws;
        bool ok;
        if (MODE == 0) {
          hs = pix_h[m] * a.sh + kh - 1;
          ws = pix_w[m] * a.sw + kw - 1;
          ok = hs >= 0 && hs < a.Hs && ws >= 0 && ws < a.Ws;
        } else {
          int th = pix_h[m] + 1 - kh, tw = pix_w[m] + 1 - kw;
          ok = th >= 0 && tw >= 0 && (th % a.sh) == 0 && (tw % a.sw) == 0;
          hs = th / a.sh; ws = tw / a.sw;
          ok = ok && hs < a.Hs && ws < a.Ws;
        }
        if (ok) v = to_f(src[(((long long)n * a.Hs + hs) * a.Ws + ws) * a.Cs + c]);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int idx = tid + i * 256;
      int n = idx >> 4;
      float v = 0.f;
      if (n < BN_ && kvalid && n_blk + n < a.Cd) v = to_f(w[(long long)(n_blk + n) * K + kg]);
      rb[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) As[kl * LDA + (tid >> 4) + i * 16] = ra[i];
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int n = (tid + i * 256) >> 4;
      if (n < BN_) Bs[kl * LDB + n] = rb[i];
    }
  };

  const int nk = (K + SIMT_BK - 1) / SIMT_BK;
  load_tiles(0);
  store_tiles();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tiles((kt + 1) * SIMT_BK);
    simt_mma_4x4<LDA, LDB, SIMT_BK>(As, Bs, ty * 4, tx * 4, acc);
    __syncthreads();
    if (kt + 1 < nk) {
      store_tiles();
      __syncthreads();
    }
  }

  T* dst = (T*)a.dst;
  const int gn = n_blk + tx * 4;
  if (gn < a.Cd) {
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE == 0 && a.bias) {
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = (gn + c < a.Cd) ? a.bias[gn + c] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      long long m = m_blk + ty * 4 + r;
      if (m >= Mtot) continue;
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[c] = acc[r][c] + bv[c];
        if (MODE == 0 && a.relu) v[c] = fmaxf(v[c], 0.f);
      }
      if (gn + 3 < a.Cd && (a.Cd & 3) == 0) {
        store4(dst + m * a.Cd + gn, v);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (gn + c < a.Cd) dst[m * a.Cd + gn + c] = from_f<T>(v[c]);
      }
    }
  }
}

template <typename T, int MODE>
int launch_conv(const ConvArgs& a, cudaStream_t st) {
  long long Mtot = (long long)a.N * a.Hd * a.Wd;
  if (Mtot <= 0) return OMR_OK;
  if (a.Cd <= 16) {
    dim3 grid((unsigned)cdiv(Mtot, 256), (unsigned)cdiv(a.Cd, 16));
    OmrLaunch(grid, 256, 0, st)(conv3x3_igemm_kernel<T, 4, MODE>, a);
  } else if (a.Cd <= 32) {
    dim3 grid((unsigned)cdiv(Mtot, 128), (unsigned)cdiv(a.Cd, 32));
    OmrLaunch(grid, 256, 0, st)(conv3x3_igemm_kernel<T, 8, MODE>, a);
  } else {
    dim3 grid((unsigned)cdiv(Mtot, 64), (unsigned)cdiv(a.Cd, 64));
    OmrLaunch(grid, 256, 0, st)(conv3x3_igemm_kernel<T, 16, MODE>, a);
  }
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- weight gradient: dw[co][ci][tap] += sum_pix dy[pix][co] * x[pix @ tap][ci] -----------------
struct WgradArgs {
  const void* x; const void* dy; float* dw;
  int N, H, W, Ci, Ho, Wo, Co, sh, sw;
  int ktiles_per_split;
};

template <typename T>
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(WgradArgs a) {
  omr_pdl_enter();
  constexpr int LDS = 68;
  __shared__ __align__(16) float As[SIMT_BK * LDS];  // [k=pix][m=co]
  __shared__ __align__(16) float Bs[SIMT_BK * LDS];  // [k=pix][n=(tap,ci)]
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m_blk = blockIdx.y * 64, n_blk = blockIdx.x * 64;
  const long long Ktot = (long long)a.N * a.Ho * a.Wo;
  const int Ntot = 9 * a.Ci;
  const T* x = (const T*)a.x;
  const T* dy = (const T*)a.dy;

  // this thread's fixed column (n) for B loads and fixed row (m) for A loads
  const int ln = tid & 63;
  const int gnB = n_blk + ln;
  const bool nvalid = gnB < Ntot;
  const int tap = nvalid ? gnB / a.Ci : 0;
  const int ci = gnB - tap * a.Ci;
  const int kh = tap / 3, kw = tap - kh * 3;
  const int gmA = m_blk + ln;
  const bool mvalid = gmA < a.Co;

  float ra[4], rb[4];
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  const long long kt_begin = (long long)blockIdx.z * a.ktiles_per_split;
  const long long nk_all = (Ktot + SIMT_BK - 1) / SIMT_BK;
  long long kt_end = kt_begin + a.ktiles_per_split;
  if (kt_end > nk_all) kt_end = nk_all;
  if (kt_begin >= kt_end) return;

  auto load_tiles = [&](long long k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long p = k0 + (tid >> 6) + i * 4;
      float va = 0.f, vb = 0.f;
      if (p < Ktot) {
        if (mvalid) va = to_f(dy[p * a.Co + gmA]);
        if (nvalid) {
          int wo = (int)(p % a.Wo);
          long long r = p / a.Wo;
          int ho = (int)(r % a.Ho);
          int n = (int)(r / a.Ho);
          int hs = ho * a.sh + kh - 1, ws = wo * a.sw + kw - 1;
          if (hs >= 0 && hs < a.H && ws >= 0 && ws < a.W) vb = to_f(x[(((long long)n * a.H + hs) * a.W + ws) * a.Ci + ci]);
        }
      }
      ra[i] = va; rb[i] = vb;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = (tid >> 6) + i * 4;
      As[k * LDS + ln] = ra[i];
      Bs[k * LDS + ln] = rb[i];
    }
  };

  load_tiles(kt_begin * SIMT_BK);
  store_tiles();
  __syncthreads();
  for (long long kt = kt_begin; kt < kt_end; ++kt) {
    if (kt + 1 < kt_end) load_tiles((kt + 1) * SIMT_BK);
    simt_mma_4x4<LDS, LDS, SIMT_BK>(As, Bs, ty * 4, tx * 4, acc);
    __syncthreads();
    if (kt + 1 < kt_end) {
      store_tiles();
      __syncthreads();
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int co = m_blk + ty * 4 + r;
    if (co >= a.Co) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int gn = n_blk + tx * 4 + c;
      if (gn >= Ntot) continue;
      int t = gn / a.Ci, cc = gn - t * a.Ci;
      atomicAdd(a.dw + ((long long)co * a.Ci + cc) * 9 + t, acc[r][c]);
    }
  }
}

}  // namespace

int omr_conv3x3_fwd_simt(int dt, const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Ci,
                         int Co, int sh, int sw, int relu, cudaStream_t st) {
  ConvArgs a{x, w, bias, y, N, H, W, Ci, (H + sh - 1) / sh, (W + sw - 1) / sw, Co, sh, sw, relu};
  OMR_DISPATCH_DT(dt, T, return (launch_conv<T, 0>(a, st)));
  return OMR_OK;
}

int omr_conv3x3_dgrad_simt(int dt, const void* dy, const void* wT, void* dx, int N, int H, int W, int Ci, int Co, int sh,
                           int sw, cudaStream_t st) {
  ConvArgs a{dy, wT, nullptr, dx, N, (H + sh - 1) / sh, (W + sw - 1) / sw, Co, H, W, Ci, sh, sw, 0};
  OMR_DISPATCH_DT(dt, T, return (launch_conv<T, 1>(a, st)));
  return OMR_OK;
}

int omr_conv3x3_wgrad_simt(int dt, const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Co, int sh,
                           int sw, int accumulate, cudaStream_t st) {
  int Ho = (H + sh - 1) / sh, Wo = (W + sw - 1) / sw;
  if (!accumulate) OMR_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Co * Ci * 9, st));
  long long Ktot = (long long)N * Ho * Wo;
  if (Ktot <= 0) return OMR_OK;
  long long nk = cdiv(Ktot, SIMT_BK);
  int tiles = (int)(cdiv(Co, 64) * cdiv(9 * Ci, 64));
  long long split = (148LL * 6) / tiles;
  if (split < 1) split = 1;
  if (split > nk) split = nk;
  long long per = cdiv(nk, split);
  split = cdiv(nk, per);
  WgradArgs a{x, dy, dw, N, H, W, Ci, Ho, Wo, Co, sh, sw, (int)per};
  dim3 grid((unsigned)cdiv(9 * Ci, 64), (unsigned)cdiv(Co, 64), (unsigned)split);
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, 256, 0, st)(conv3x3_wgrad_kernel<T>, a)));
  OMR_LAUNCHED();
  return OMR_OK;
}
