// gemm_tc.cu -- bf16 GEMM on the 5th-generation tensor cores: TMA (cp.async.bulk.tensor, 128B swizzle) ->
// shared-memory ring -> tcgen05.mma with the fp32 accumulator in TMEM -> tcgen05.ld epilogue with fused
// bias / ReLU / accumulate and bf16 or fp32 output.  Serves every projection of the decoder, the point-wise
// convolutions of the encoder and the vocabulary classifier, forward and backward:
//     forward   y  = x W^T        A = x  (K-major)   B = W  (K-major)
//     dgrad     dx = dy W         A = dy (K-major)   B = W  (MN-major: rows are the reduction index)
//     wgrad     dW += dy^T x      A = dy (MN-major)  B = x  (MN-major), split along the (long) reduction
//                                 dimension with fp32 atomic accumulation into the gradient
// One CTA computes one 128 x BN output tile (x one K split): warps 0-3 epilogue (TMEM lanes 32w..32w+31),
// warp 4 TMA producer, warp 5 MMA issuer + TMEM allocator.  The decoder's GEMMs have K = 256, i.e. they are
// HBM-bound (AI ~ 50 flop/B): BN covers the whole N when N <= 256 so that A is read from HBM exactly once,
// and 2-3 CTAs are resident per SM so that one tile's stores overlap the next tile's loads.
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB

struct GemmTcArgs {
  void* C;
  long long ldc;
  int M, N, K;
  int kb_per_split;  // k-blocks (of 64) per blockIdx.z
  const float* bias;
  int bias_mode, relu;
  int acc_mode;  // 0 overwrite, 1 read-modify-write, 2 atomicAdd (fp32 output only)
  int stages;    // smem ring depth: 2 for the short-K (K <= 512) decoder GEMMs so that 2 CTAs share an SM and one tile's
                 // epilogue overlaps the other's loads; 3-4 for the long reductions
};

template <int BN>
constexpr int num_stages() { return BN == 256 ? 4 : (BN == 128 ? 3 : 4); }
template <int BN>
constexpr int smem_bytes(int stages) { return stages * (A_TILE_BYTES + BN * BK * 2) + 1024 + 256; }
template <int BN>
constexpr int epilogue_bytes(int out_esz) { return 128 * BN * out_esz; }  // the staging rows alias the ring

template <int BN, int A_MN, int B_MN, typename TO>
__global__ void __launch_bounds__(320) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                     const __grid_constant__ CUtensorMap tmB, GemmTcArgs g) {
  omr_pdl_enter();
  const int STAGES = g.stages;
  constexpr int B_TILE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb_total = (g.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * g.kb_per_split;
  int kb_end = kb_begin + g.kb_per_split;
  if (kb_end > kb_total) kb_end = kb_total;
  const int nkb = kb_end - kb_begin;  // >= 1 by construction of the grid

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);

  if (warp == 8) {
    {  // TMA producer: warp-uniform loop (see tc_common.cuh), one elected lane issues
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_TILE_BYTES;
        const int k0 = (kb_begin + i) * BK;
        if (elect_one()) {
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        if (A_MN == 0) {
          tma_load_2d(a_dst, &tmA, &full_bar[s], k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_2d(a_dst + c * 8192, &tmA, &full_bar[s], m0 + c * 64, k0);
        }
        if (B_MN == 0) {
          tma_load_2d(b_dst, &tmB, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * 8192, &tmB, &full_bar[s], n0 + c * 64, k0);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    {  // MMA issuer: warp-uniform loop, tcgen05.mma / commit under elect_one()
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      const uint32_t smem0 = smem_u32(smem);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem0 + (uint32_t)s * STAGE_BYTES;
        const uint32_t b_addr = a_addr + A_TILE_BYTES;
        if (elect_one()) {
#pragma unroll
        for (int j = 0; j < BK / UMMA_K; ++j) {
          const uint64_t ad = A_MN == 0 ? make_smem_desc(a_addr + j * 32, 16, 1024, 128)
                                        : make_smem_desc(a_addr + j * 2048, 8192, 1024, 128);
          const uint64_t bd = B_MN == 0 ? make_smem_desc(b_addr + j * 32, 16, 1024, 128)
                                        : make_smem_desc(b_addr + j * 2048, 8192, 1024, 128);
          if (j > 0)
            umma_bf16_acc(tmem_base, ad, bd, idesc);
          else
            umma_bf16(tmem_base, ad, bd, idesc, i > 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
        if (i == nkb - 1) umma_commit(accum_bar);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: 8 warps; warp w owns TMEM lanes / tile rows 32(w&3) .. +31 and column half w>>2 (BN >= 64) ----
    // phase 1: thread = row: TMEM -> registers -> bias/ReLU -> output type -> this warp's staging rows in the
    // (now idle) pipeline buffers, 16-byte units XOR-swizzled by the row so that neither phase bank-conflicts;
    // phase 2: consecutive lanes write consecutive 16-byte units of a row -> fully coalesced global stores.
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    constexpr int HALVES = BN >= 64 ? 2 : 1;
    constexpr int HC = BN / HALVES;  // columns per warp
    const int quarter = warp & 3, half = warp >> 2;
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < g.M;
    const float brow = (g.bias_mode == 2 && row_ok) ? g.bias[row] : 0.f;
    const bool vec_ok = ((g.ldc * sizeof(TO)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
    constexpr int EPU = 16 / (int)sizeof(TO);      // elements per 16-byte unit
    constexpr int U = HC / EPU;                    // units per staged row (this warp's column half)
    constexpr int UMASK = U >= 8 ? 7 : (U - 1);
    uint4* stage = reinterpret_cast<uint4*>(smem) + (half * 4 + quarter) * 32 * U;
    const int nh0 = n0 + half * HC;  // first global column of this warp's half
    const bool staged = vec_ok && g.acc_mode != 2;
#pragma unroll 1
    for (int c0 = 0; c0 < HC; c0 += 32) {
      if (half >= HALVES || nh0 + c0 >= g.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * HC + c0), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + brow;
      const int ncol = nh0 + c0;
      const bool full = ncol + 32 <= g.N;
      if (g.bias_mode == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || ncol + j < g.N) v[j] += __ldg(g.bias + ncol + j);
      }
      if (g.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (g.acc_mode == 2) {
        if constexpr (sizeof(TO) == 4) {
          if (row_ok) {
            float* cf = reinterpret_cast<float*>(g.C) + (long long)row * g.ldc + ncol;
            if (full && vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cf + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                             "f"(v[j + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || ncol + j < g.N) atomicAdd(cf + j, v[j]);
            }
          }
        }
        continue;
      }
      if (staged) {
        uint4* srow = stage + lane * U;
        if constexpr (sizeof(TO) == 4) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int u = c0 / EPU + q;
            srow[u ^ (lane & UMASK)] = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                                                  __float_as_uint(v[4 * q + 3]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int u = c0 / EPU + q;
            srow[u ^ (lane & UMASK)] = make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                                  pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
          }
        }
      } else if (row_ok) {
        TO* crow = reinterpret_cast<TO*>(g.C) + (long long)row * g.ldc;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (ncol + j < g.N) {
            float o = v[j];
            if (g.acc_mode == 1) o += to_f(crow[ncol + j]);
            crow[ncol + j] = from_f<TO>(o);
          }
        }
      }
    }
    if (staged && half < HALVES) {
      __syncwarp();
      int ncols = g.N - nh0;
      if (ncols > HC) ncols = HC;
      if (ncols < 0) ncols = 0;
      const int nunits = (ncols + EPU - 1) / EPU;  // units that hold at least one valid column
#pragma unroll 1
      for (int idx = lane; idx < 32 * U; idx += 32) {
        const int rr = idx / U, u = idx - rr * U;
        const int grow = m0 + quarter * 32 + rr;
        if (grow >= g.M || u >= nunits) continue;
        uint4 val = stage[rr * U + (u ^ (rr & UMASK))];
        TO* dst = reinterpret_cast<TO*>(g.C) + (long long)grow * g.ldc + nh0 + u * EPU;
        if ((u + 1) * EPU <= ncols) {
          if (g.acc_mode == 1) {
            const uint4 old = *reinterpret_cast<const uint4*>(dst);
            if constexpr (sizeof(TO) == 4) {
              val.x = __float_as_uint(__uint_as_float(val.x) + __uint_as_float(old.x));
              val.y = __float_as_uint(__uint_as_float(val.y) + __uint_as_float(old.y));
              val.z = __float_as_uint(__uint_as_float(val.z) + __uint_as_float(old.z));
              val.w = __float_as_uint(__uint_as_float(val.w) + __uint_as_float(old.w));
            } else {
              const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
              __nv_bfloat162* nb = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                nb[e] = __floats2bfloat162_rn(__low2float(nb[e]) + __low2float(ob[e]), __high2float(nb[e]) + __high2float(ob[e]));
            }
          }
          *reinterpret_cast<uint4*>(dst) = val;
        } else {  // the unit straddles N: element-wise tail
          const TO* sv = reinterpret_cast<const TO*>(&val);
          for (int e = 0; e < EPU && u * EPU + e < ncols; ++e) {
            float o = to_f(sv[e]);
            if (g.acc_mode == 1) o += to_f(dst[e]);
            dst[e] = from_f<TO>(o);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int A_MN, int B_MN, typename TO>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcArgs& g, dim3 grid, cudaStream_t st) {
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, TO>;
  static bool configured = false;  // per template instantiation
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int smem = smem_bytes<BN>(g.stages);
  const int need = epilogue_bytes<BN>((int)sizeof(TO)) + 1024 + 256;
  if (smem < need) smem = need;
  OmrLaunch(grid, 320, smem, st)(kern, tmA, tmB, g);
  OMR_LAUNCHED();
  return OMR_OK;
}

template <int BN, typename TO>
int launch_major(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcArgs& g, dim3 grid,
                 cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, 0, 0, TO>(tmA, tmB, g, grid, st);
  if constexpr (BN >= 64) {
    if (!a_mn && b_mn) return launch<BN, 0, 1, TO>(tmA, tmB, g, grid, st);
    if (a_mn && b_mn) return launch<BN, 1, 1, TO>(tmA, tmB, g, grid, st);
    if (a_mn && !b_mn) return launch<BN, 1, 0, TO>(tmA, tmB, g, grid, st);
  }
  return OMR_TC_NOT_ELIGIBLE;
}

}  // namespace

int omr_gemm_tc(int out_dt, int transA, int transB, int M, int N, int K, const void* A, long long lda,
                long long strideA, const void* B, long long ldb, long long strideB, void* C, long long ldc,
                long long strideC, int batch, const float* bias, int bias_mode, int relu, int accumulate,
                cudaStream_t st) {
  (void)strideA; (void)strideB; (void)strideC;
  if (batch != 1 || K < 1 || M < 1 || N < 1) return OMR_TC_NOT_ELIGIBLE;
  const int a_mn = transA ? 1 : 0;  // A[k*lda + m]: M contiguous
  const int b_mn = transB ? 0 : 1;  // B[k*ldb + n]: N contiguous
  if ((lda * 2) % 16 || (ldb * 2) % 16) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return OMR_TC_NOT_ELIGIBLE;
  if (out_dt != OMR_F32 && out_dt != OMR_BF16) return OMR_TC_NOT_ELIGIBLE;
  // tiny problems stay on the CUDA-core kernels (launch/TMEM set-up would dominate), and so do the skinny forward
  // GEMMs of the decode step (M <= 64): a 128-row tile would leave 1-3 CTAs to stream the whole weight matrix
  if ((long long)M * N * K < (1LL << 18)) return OMR_TC_NOT_ELIGIBLE;
  if (M <= 64 && !a_mn && !b_mn && !accumulate) return OMR_TC_NOT_ELIGIBLE;
  int BN = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  static int bn_cap = -1, split_div = -1;
  if (bn_cap < 0) {
    const char* e = getenv("OMR_GEMM_BN_CAP");
    bn_cap = e ? atoi(e) : 128;
    const char* f = getenv("OMR_GEMM_SPLIT_DIV");
    split_div = f ? atoi(f) : 8;
    if (split_div < 1) split_div = 8;
  }
  // short-K problems with few tiles (the decoder's 256-wide projections): narrower tiles double the CTA count so that
  // two CTAs share an SM and one's loads overlap the other's epilogue
  if (bn_cap > 0 && BN > bn_cap && (long long)((M + BM - 1) / BM) * ((N + BN - 1) / BN) < 2 * 148) BN = bn_cap;
  if (b_mn && BN < 64) BN = 64;
  if (a_mn && BN < 64) BN = 64;

  CUtensorMap tmA, tmB;
  {
    unsigned long long dims[2], strides[1];
    unsigned int box[2];
    if (!a_mn) { dims[0] = K; dims[1] = M; box[0] = BK; box[1] = BM; }
    else       { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = BK; }
    strides[0] = (unsigned long long)lda * 2;
    int rc = omr_make_tensor_map(&tmA, 2, A, 2, dims, strides, box, nullptr, 128);
    if (rc) return rc;
    if (!b_mn) { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = (unsigned)BN; }
    else       { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = BK; }
    strides[0] = (unsigned long long)ldb * 2;
    rc = omr_make_tensor_map(&tmB, 2, B, 2, dims, strides, box, nullptr, 128);
    if (rc) return rc;
  }
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int kb_total = (K + BK - 1) / BK;
  int splits = 1;
  int acc_mode = accumulate ? 1 : 0;
  if (accumulate && out_dt == OMR_F32 && !relu) {
    // gradient accumulation: split the long reduction across CTAs until the machine is full
    long long tiles = (long long)tiles_m * tiles_n;
    long long want = (148 + tiles - 1) / tiles;
    long long cap = kb_total / split_div;
    if (want > cap) want = cap;
    if (want > 1) { splits = (int)want; acc_mode = 2; }
  }
  int kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per_split - 1) / kb_per_split;
  int stages = kb_per_split <= 8 ? 2 : (BN == 128 ? 3 : 4);
  if (stages > kb_per_split) stages = kb_per_split < 2 ? 2 : kb_per_split;
  GemmTcArgs g{C, ldc, M, N, K, kb_per_split, bias, bias ? bias_mode : 0, relu, acc_mode, stages};
  if (acc_mode == 2 && g.bias_mode != 0) return OMR_TC_NOT_ELIGIBLE;
  dim3 grid((unsigned)tiles_m, (unsigned)tiles_n, (unsigned)splits);
#define OMR_GEMM_BN(BNV)                                                                         \
  (out_dt == OMR_F32 ? launch_major<BNV, float>(a_mn, b_mn, tmA, tmB, g, grid, st)               \
                     : launch_major<BNV, bf16>(a_mn, b_mn, tmA, tmB, g, grid, st))
  switch (BN) {
    case 32: return OMR_GEMM_BN(32);
    case 64: return OMR_GEMM_BN(64);
    case 128: return OMR_GEMM_BN(128);
    default: return OMR_GEMM_BN(256);
  }
#undef OMR_GEMM_BN
}
