// wgrad_small.cu -- 3x3 weight gradient of the NARROW full-resolution layers (C_in, C_out in {16, 32}, stride 1).
//
//   dW[co, ci, kh, kw] (+)= sum over pixels p of  dY[p, co] * X[p + (kh-1, kw-1), ci]
//
// These layers hold 70 % of the weight-gradient time of the encoders but are HBM-bound by nature (a pixel row is 32 or
// 64 bytes; 23 GFLOP over 322 MB for 16->16 at 195 x 808 x 32).  On the tcgen05 path (wgrad_tc.cu) both operands are
// MN-major with 32/64-byte rows and every tcgen05.mma costs ~137 clk whatever its N (measured):
// the operand fetch, not the math, sets the pace, 9x above the HBM floor.  Here the same tiles -- a TH x TW patch of dY
// and its (TH+2) x (TW+2) halo of X, fetched by TMA with the swizzle mode of the row width -- are read with
// ldmatrix.trans, which delivers exactly the fragments of mma.sync.m16n8k16 for a pixel-major layout:
//     A (co x pixel)  <- ldmatrix.x4.trans over 16 pixel rows of dY  (a thread gets 2 consecutive pixels of one channel)
//     B (pixel x ci)  <- ldmatrix.x4.trans over the SAME 16 pixels of X shifted by the tap: any 8 consecutive pixel
//                        rows hit 8 distinct bank groups under the TMA swizzle, so the nine shifted reads of a tile
//                        are conflict-free and X is fetched from HBM once.
// A warp keeps one 16 x 16 (co x ci) block of all nine taps in registers (72 fp32) across its whole pixel range; wider
// layers split the blocks over the warps.  Partial sums are combined in shared memory and added to the fp32 gradient
// with one atomic per element and CTA.  Warp roles: 0-7 MMA, 8 TMA producer.
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int WS_TH = 8, WS_TW = 32;  // output pixels per tile: 8 rows x 32 columns
constexpr int WS_PITCH = WS_TW + 2;   // halo row pitch in pixels
constexpr int WS_WARPS = 8;           // consumer warps

struct WsArgs {
  float* dw;
  int Ci, Co;
  int tiles_w, tiles_h, num_tiles;
  int stages, a_bytes, stage_bytes;
};

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `c16` of pixel row `R` in a tile of RB-byte rows written by TMA with SWIZZLE_<RB>B:
// address bits [4, 4+log2(RB/16)) are XORed with bits [7, ...) (cute::Swizzle<1|2, 4, 3>)
template <int RB>
__device__ __forceinline__ uint32_t swz(uint32_t R, uint32_t c16) {
  const uint32_t o = R * RB + c16 * 16;
  return o ^ (((o >> 7) & (RB / 16 - 1)) << 4);
}

template <int RBA, int RBB>  // row bytes of dY (2 Co) and X (2 Ci)
__global__ void __launch_bounds__((WS_WARPS + 1) * 32) wgrad_small_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                                          const __grid_constant__ CUtensorMap tmX, WsArgs g) {
  omr_pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + g.stages * g.stage_bytes);
  uint64_t* empty_bar = full_bar + g.stages;
  const int warp = (int)tc::warp_idx_sync(), lane = threadIdx.x & 31;
  constexpr int MT = RBA / 32, NH = RBB / 32;  // 16-channel blocks of dY / X
  constexpr int NCOMBO = MT * NH;              // (co block, ci block) pairs, dealt to the warps
  constexpr int ROWS = NCOMBO;                 // tile rows per warp (8 warps cover NCOMBO blocks x 8 / NCOMBO row groups)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], WS_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int my_tiles = g.num_tiles > (int)blockIdx.x ? (g.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == WS_WARPS) {
    {  // TMA producer: warp-uniform loop, one elected lane issues (see tc_common.cuh)
      const uint32_t tx = (uint32_t)(WS_TH * WS_TW * RBA + (WS_TH + 2) * WS_PITCH * RBB);
      int s = 0;
      uint32_t ph = 1;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int tw = tile % g.tiles_w;
        const int th = (tile / g.tiles_w) % g.tiles_h;
        const int n = tile / (g.tiles_w * g.tiles_h);
        mbar_wait(&empty_bar[s], ph);
        if (tc::elect_one()) {
          mbar_expect_tx(&full_bar[s], tx);
          uint8_t* stage = smem + s * g.stage_bytes;
          tma_load_4d(stage, &tmDY, &full_bar[s], 0, tw * WS_TW, th * WS_TH, n);
          tma_load_4d(stage + g.a_bytes, &tmX, &full_bar[s], 0, tw * WS_TW - 1, th * WS_TH - 1, n);
        }
        __syncwarp();
        if (++s == g.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    const int combo = warp % NCOMBO, pg = warp / NCOMBO;
    const int mt = combo % MT, nh = combo / MT;
    float acc[9][2][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[t][j][e] = 0.f;
    const int mi = lane >> 3, row8 = lane & 7;
    // A: matrices (pixels 0-7 | 8-15) x (channels 0-7 | 8-15) in the order a0, a1, a2, a3
    const uint32_t a_col = (uint32_t)((mi >> 1) * 8 + row8), a_c16 = (uint32_t)(mt * 2 + (mi & 1));
    // B: b0, b1 of the first 8 input channels, then of the next 8
    const uint32_t b_col = (uint32_t)((mi & 1) * 8 + row8), b_c16 = (uint32_t)(nh * 2 + (mi >> 1));
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i % g.stages;
      mbar_wait(&full_bar[s], (i / g.stages) & 1);
      const uint32_t a_base = smem_u32(smem + s * g.stage_bytes), b_base = a_base + (uint32_t)g.a_bytes;
#pragma unroll 1
      for (int rr = 0; rr < ROWS; ++rr) {
        const uint32_t r = (uint32_t)(pg * ROWS + rr);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t c0 = (uint32_t)ks * 16;
          uint32_t a[4];
          ldsm_x4_trans(a_base + swz<RBA>(r * WS_TW + c0 + a_col, a_c16), a);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              uint32_t b[4];
              ldsm_x4_trans(b_base + swz<RBB>((r + kh) * WS_PITCH + c0 + b_col + kw, b_c16), b);
              mma_16816(acc[kh * 3 + kw][0], a, b[0], b[1]);
              mma_16816(acc[kh * 3 + kw][1], a, b[2], b[3]);
            }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    // ---- combine the warps of the CTA in shared memory (the pipeline buffers are free once every warp is here) ----
    asm volatile("bar.sync 1, %0;" ::"n"(WS_WARPS * 32) : "memory");
    float* red = reinterpret_cast<float*>(smem);
    const int nel = g.Co * g.Ci * 9;
    for (int k = threadIdx.x; k < nel; k += WS_WARPS * 32) red[k] = 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(WS_WARPS * 32) : "memory");
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = mt * 16 + gq + (e >> 1) * 8;
          const int ci = nh * 16 + j * 8 + 2 * tq + (e & 1);
          atomicAdd(&red[(co * g.Ci + ci) * 9 + t], acc[t][j][e]);
        }
    asm volatile("bar.sync 1, %0;" ::"n"(WS_WARPS * 32) : "memory");
    for (int k = threadIdx.x; k < nel; k += WS_WARPS * 32) atomicAdd(g.dw + k, red[k]);
  }
}

int ws_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int RBA, int RBB>
int launch_ws(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WsArgs& g, int smem_bytes, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(wgrad_small_kernel<RBA, RBB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  OmrLaunch(grid, (WS_WARPS + 1) * 32, smem_bytes, st)(wgrad_small_kernel<RBA, RBB>, tmDY, tmX, g);
  OMR_LAUNCHED();
  return OMR_OK;
}

}  // namespace

int omr_conv3x3_wgrad_small(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Co, int sh, int sw,
                            int accumulate, cudaStream_t st) {
  if (!((Ci == 16 || Ci == 32) && (Co == 16 || Co == 32)) || sh != 1 || sw != 1 || N < 1) return OMR_TC_NOT_ELIGIBLE;
  // 32 -> 32: since the tcgen05 issue loops became warp-uniform (round 2) the tcgen05 weight-gradient kernel is faster than this
  // legacy-MMA kernel, which sits at 64 % of the HMMA pipe (ncu): 250 vs 276 us at 128 x 1024 x 32, 304 vs 338 us at 195 x 808 x 32
  if (Ci == 32 && Co == 32) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15)) return OMR_TC_NOT_ELIGIBLE;
  WsArgs g{};
  g.dw = dw; g.Ci = Ci; g.Co = Co;
  g.tiles_w = (W + WS_TW - 1) / WS_TW;
  g.tiles_h = (H + WS_TH - 1) / WS_TH;
  g.num_tiles = N * g.tiles_h * g.tiles_w;
  const int rba = Co * 2, rbb = Ci * 2;
  g.a_bytes = (WS_TH * WS_TW * rba + 1023) / 1024 * 1024;
  g.stage_bytes = g.a_bytes + ((WS_TH + 2) * WS_PITCH * rbb + 1023) / 1024 * 1024;
  // two CTAs per SM hide the ldmatrix / mma.sync latencies better than a deeper pipeline: 3 stages where 4 would
  // leave room for only one CTA
  g.stages = 4;
  {
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("OMR_WGRAD_SMALL_STAGES");
      forced = e ? atoi(e) : 0;
    }
    if (forced >= 2 && forced <= 8) g.stages = forced;
    else if (2 * (4 * g.stage_bytes + 3 * 1024) > 220 * 1024 && 2 * (3 * g.stage_bytes + 3 * 1024) <= 220 * 1024) g.stages = 3;
    else if (2 * (3 * g.stage_bytes + 3 * 1024) > 220 * 1024 && 2 * (2 * g.stage_bytes + 3 * 1024) <= 220 * 1024) g.stages = 2;
  }
  int smem_bytes = g.stages * g.stage_bytes + 1024 + 256;
  const int red_bytes = Co * Ci * 9 * 4 + 1024 + 256;
  if (smem_bytes < red_bytes) smem_bytes = red_bytes;
  int per_sm = (220 * 1024) / (smem_bytes + 1024);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) return OMR_TC_NOT_ELIGIBLE;
  int grid = ws_sms() * per_sm;
  if (grid > g.num_tiles) grid = g.num_tiles;

  CUtensorMap tmDY, tmX;
  {
    unsigned long long dims[4] = {(unsigned long long)Co, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N};
    unsigned long long strides[3] = {(unsigned long long)Co * 2, (unsigned long long)W * Co * 2, (unsigned long long)H * W * Co * 2};
    unsigned int box[4] = {(unsigned)Co, (unsigned)WS_TW, (unsigned)WS_TH, 1u};
    int rc = omr_make_tensor_map(&tmDY, 2, dy, 4, dims, strides, box, nullptr, rba);
    if (rc) return rc;
    unsigned long long xd[4] = {(unsigned long long)Ci, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N};
    unsigned long long xs[3] = {(unsigned long long)Ci * 2, (unsigned long long)W * Ci * 2, (unsigned long long)H * W * Ci * 2};
    unsigned int xb[4] = {(unsigned)Ci, (unsigned)WS_PITCH, (unsigned)(WS_TH + 2), 1u};
    rc = omr_make_tensor_map(&tmX, 2, x, 4, xd, xs, xb, nullptr, rbb);
    if (rc) return rc;
  }
  if (!accumulate) OMR_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Co * Ci * 9, st));
  if (rba == 32 && rbb == 32) return launch_ws<32, 32>(tmDY, tmX, g, smem_bytes, grid, st);
  if (rba == 64 && rbb == 32) return launch_ws<64, 32>(tmDY, tmX, g, smem_bytes, grid, st);
  if (rba == 32 && rbb == 64) return launch_ws<32, 64>(tmDY, tmX, g, smem_bytes, grid, st);
  return launch_ws<64, 64>(tmDY, tmX, g, smem_bytes, grid, st);
}
