// gemm_simt.cu -- exact-fp32-accumulate CUDA-core (batched) GEMM with fused bias / ReLU / accumulate.
// Serves every fp32-mode projection and all shapes the tcgen05 GEMM does not take.
#include "simt_tile.cuh"

namespace {

constexpr int BM = 64, BN = 64, LDS = 68;

struct GemmArgs {
  const void* A; const void* B; void* C;
  long long lda, ldb, ldc, sA, sB, sC;
  int M, N, K;
  const float* bias; int bias_mode; int relu; int accumulate;
};

template <typename TI, typename TO, int TA, int TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  omr_pdl_enter();
  __shared__ __align__(16) float As[SIMT_BK * LDS];
  __shared__ __align__(16) float Bs[SIMT_BK * LDS];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m_blk = blockIdx.y * BM, n_blk = blockIdx.x * BN;
  const TI* A = (const TI*)g.A + (long long)blockIdx.z * g.sA;
  const TI* B = (const TI*)g.B + (long long)blockIdx.z * g.sB;
  TO* C = (TO*)g.C + (long long)blockIdx.z * g.sC;

  float ra[4], rb[4];
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int m, k;
      if (TA == 0) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      int gm = m_blk + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < g.K) v = to_f(TA == 0 ? A[(long long)gm * g.lda + gk] : A[(long long)gk * g.lda + gm]);
      ra[i] = v;
      int n, kb;
      if (TB == 0) { n = idx & 63; kb = idx >> 6; } else { n = idx >> 4; kb = idx & 15; }
      int gn = n_blk + n, gkb = k0 + kb;
      float w = 0.f;
      if (gn < g.N && gkb < g.K) w = to_f(TB == 0 ? B[(long long)gkb * g.ldb + gn] : B[(long long)gn * g.ldb + gkb]);
      rb[i] = w;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int m, k;
      if (TA == 0) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      As[k * LDS + m] = ra[i];
      int n, kb;
      if (TB == 0) { n = idx & 63; kb = idx >> 6; } else { n = idx >> 4; kb = idx & 15; }
      Bs[kb * LDS + n] = rb[i];
    }
  };

  const int nk = (g.K + SIMT_BK - 1) / SIMT_BK;
  if (nk > 0) {
    load_tiles(0);
    store_tiles();
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tiles((kt + 1) * SIMT_BK);
    simt_mma_4x4<LDS, LDS, SIMT_BK>(As, Bs, ty * 4, tx * 4, acc);
    __syncthreads();
    if (kt + 1 < nk) {
      store_tiles();
      __syncthreads();
    }
  }

#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int gm = m_blk + ty * 4 + r;
    if (gm >= g.M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int gn = n_blk + tx * 4 + c;
      if (gn >= g.N) continue;
      float v = acc[r][c];
      if (g.bias_mode == 1) v += g.bias[gn];
      else if (g.bias_mode == 2) v += g.bias[gm];
      if (g.relu) v = fmaxf(v, 0.f);
      long long o = (long long)gm * g.ldc + gn;
      if (g.accumulate) v += to_f(C[o]);
      C[o] = from_f<TO>(v);
    }
  }
}

// Skinny GEMM for the greedy-decode step (M = batch <= 32 rows per pass):  C[M,N] = act(A[M,K] W[N,K]^T + bias).
// The problem is a weight stream (N*K elements read once), so it is spread over N/8 blocks instead of the 1-3 tiles a
// 128-row tensor-core tile would give.  Per 256-wide K slab the block stages A (<= 32 x 256) and its 8 weight rows in
// shared memory with coalesced 16-byte loads; then lane = row m, warp = 2 output columns: one conflict-free 16-byte
// LDS of x, two broadcast LDS of w and 16 FMAs per 8 k -- no shuffles, no dependent global loads in the loop.
constexpr int SK_KS = 256;          // K slab
constexpr int SK_LDX = SK_KS + 8;   // padded row (elements) -> 16-byte units of consecutive rows land in distinct banks
template <typename TI, typename TO>
__global__ void __launch_bounds__(128) gemm_skinny_kernel(GemmArgs g) {
  omr_pdl_enter();
  __shared__ __align__(16) TI sx[32 * SK_LDX];
  __shared__ __align__(16) TI sw[8 * SK_LDX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = blockIdx.x * 8;
  const TI* A = (const TI*)g.A;
  const TI* B = (const TI*)g.B;
  TO* C = (TO*)g.C;
  constexpr int VEC = 16 / (int)sizeof(TI);  // elements per 16-byte unit
  for (int mb = 0; mb < g.M; mb += 32) {
    float acc[2] = {0.f, 0.f};
    for (int k0 = 0; k0 < g.K; k0 += SK_KS) {
      const int kn = g.K - k0 < SK_KS ? g.K - k0 : SK_KS;  // multiple of 8
      const int units = kn / VEC;
      __syncthreads();
      for (int idx = threadIdx.x; idx < 32 * units; idx += 128) {
        const int r = idx / units, u = idx - r * units;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (mb + r < g.M) v = *reinterpret_cast<const uint4*>(A + (long long)(mb + r) * g.lda + k0 + u * VEC);
        *reinterpret_cast<uint4*>(sx + r * SK_LDX + u * VEC) = v;
      }
      for (int idx = threadIdx.x; idx < 8 * units; idx += 128) {
        const int r = idx / units, u = idx - r * units;
        const int n = nb + r < g.N ? nb + r : g.N - 1;
        *reinterpret_cast<uint4*>(sw + r * SK_LDX + u * VEC) = *reinterpret_cast<const uint4*>(B + (long long)n * g.ldb + k0 + u * VEC);
      }
      __syncthreads();
      const TI* xr = sx + lane * SK_LDX;
      const TI* w0 = sw + (warp * 2) * SK_LDX;
      const TI* w1 = w0 + SK_LDX;
#pragma unroll 4
      for (int k = 0; k < kn; k += 8) {
        float x[8], a0[8], a1[8];
        load4(xr + k, *reinterpret_cast<float(*)[4]>(x));
        load4(xr + k + 4, *reinterpret_cast<float(*)[4]>(x + 4));
        load4(w0 + k, *reinterpret_cast<float(*)[4]>(a0));
        load4(w0 + k + 4, *reinterpret_cast<float(*)[4]>(a0 + 4));
        load4(w1 + k, *reinterpret_cast<float(*)[4]>(a1));
        load4(w1 + k + 4, *reinterpret_cast<float(*)[4]>(a1 + 4));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[0] = fmaf(x[e], a0[e], acc[0]);
          acc[1] = fmaf(x[e], a1[e], acc[1]);
        }
      }
    }
    const int m = mb + lane;
    if (m < g.M) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n = nb + warp * 2 + c;
        if (n >= g.N) continue;
        float v = acc[c];
        if (g.bias_mode == 1) v += g.bias[n];
        else if (g.bias_mode == 2) v += g.bias[m];
        if (g.relu) v = fmaxf(v, 0.f);
        const long long o = (long long)m * g.ldc + n;
        if (g.accumulate) v += to_f(C[o]);
        C[o] = from_f<TO>(v);
      }
    }
  }
}

template <typename TI, typename TO>
int launch_gemm(const GemmArgs& g, int transA, int transB, int batch, cudaStream_t st) {
  if (batch == 1 && g.M <= 64 && transA == 0 && transB == 1 && g.K % 8 == 0 && g.lda % 8 == 0 && g.ldb % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0) {
    OmrLaunch((unsigned)cdiv(g.N, 8), 128, 0, st)(gemm_skinny_kernel<TI, TO>, g);
    OMR_LAUNCHED();
    return OMR_OK;
  }
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), (unsigned)batch);
  if (transA == 0 && transB == 0) OmrLaunch(grid, 256, 0, st)(gemm_simt_kernel<TI, TO, 0, 0>, g);
  else if (transA == 0 && transB == 1) OmrLaunch(grid, 256, 0, st)(gemm_simt_kernel<TI, TO, 0, 1>, g);
  else if (transA == 1 && transB == 0) OmrLaunch(grid, 256, 0, st)(gemm_simt_kernel<TI, TO, 1, 0>, g);
  else OmrLaunch(grid, 256, 0, st)(gemm_simt_kernel<TI, TO, 1, 1>, g);
  OMR_LAUNCHED();
  return OMR_OK;
}

}  // namespace

int omr_gemm_simt(int in_dt, int out_dt, int transA, int transB, int M, int N, int K, const void* A, long long lda,
                  long long strideA, const void* B, long long ldb, long long strideB, void* C, long long ldc,
                  long long strideC, int batch, const float* bias, int bias_mode, int relu, int accumulate,
                  cudaStream_t st) {
  GemmArgs g{A, B, C, lda, ldb, ldc, strideA, strideB, strideC, M, N, K, bias, bias ? bias_mode : 0, relu, accumulate};
  if (in_dt == OMR_F32 && out_dt == OMR_F32) return launch_gemm<float, float>(g, transA, transB, batch, st);
  if (in_dt == OMR_BF16 && out_dt == OMR_BF16) return launch_gemm<bf16, bf16>(g, transA, transB, batch, st);
  if (in_dt == OMR_BF16 && out_dt == OMR_F32) return launch_gemm<bf16, float>(g, transA, transB, batch, st);
  if (in_dt == OMR_F32 && out_dt == OMR_BF16) return launch_gemm<float, bf16>(g, transA, transB, batch, st);
  omr_set_error("omr_gemm: unsupported dtypes in=%d out=%d", in_dt, out_dt);
  return OMR_ERR_INVALID;
}
