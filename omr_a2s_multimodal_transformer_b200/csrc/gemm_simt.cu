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

template <typename TI, typename TO>
int launch_gemm(const GemmArgs& g, int transA, int transB, int batch, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), (unsigned)batch);
  if (transA == 0 && transB == 0) gemm_simt_kernel<TI, TO, 0, 0><<<grid, 256, 0, st>>>(g);
  else if (transA == 0 && transB == 1) gemm_simt_kernel<TI, TO, 0, 1><<<grid, 256, 0, st>>>(g);
  else if (transA == 1 && transB == 0) gemm_simt_kernel<TI, TO, 1, 0><<<grid, 256, 0, st>>>(g);
  else gemm_simt_kernel<TI, TO, 1, 1><<<grid, 256, 0, st>>>(g);
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
