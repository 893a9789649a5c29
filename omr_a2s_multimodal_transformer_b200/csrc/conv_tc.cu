// conv_tc.cu -- 3x3 convolution (forward and data gradient) as an implicit GEMM on the tcgen05 tensor cores.
//
//   D[pixel, co] = sum over taps t, channels c of  X[n, oh*ish + dh_t, ow*isw + dw_t, c] * Wt[co, widx_t, c]
//
// * activations are NHWC bf16; the A operand of tap t is fetched by ONE 4-D TMA box per (tap, 64-channel chunk)
//   whose start coordinate carries the tap shift -- out-of-range coordinates (the zero padding) are filled with
//   zeros by the TMA unit and strided convolutions use the tensor map's element strides, so no thread ever
//   computes an im2col address;
// * one smem row = one pixel = min(C,64) channels (32/64/128 bytes) in the TMA swizzle of that width, which is
//   exactly the K-major UMMA operand layout; the weight slice [Cout x chunk] of the tap is the B operand;
// * persistent CTAs (one per SM) walk the output tiles (128 pixels = TH x TW patch of one image); the fp32
//   accumulator lives in TMEM and is double buffered, so the epilogue of tile i (bias + ReLU + bf16 NHWC store)
//   overlaps the TMA/MMA main loop of tile i+1;
// * the same kernel computes the data gradient: stride 1 -> taps mirrored, weights from the [Ci,3,3,Co] pack;
//   stride 2 -> one launch per output parity class (1, 2, 2 or 4 taps each), written with a strided epilogue.
// Warp roles: 0-3 epilogue (TMEM lanes 32w..32w+31), 4 TMA producer, 5 MMA issuer + TMEM allocator.
#define OMR_HAVE_TC_CONV 1
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int MAX_TAPS = 9;

struct ConvTcArgs {
  bf16* y;
  const float* bias;
  int relu;
  int N, Cin, Cout;
  int TH, TW, tiles_h, tiles_w, num_tiles;  // logical output grid tiling
  int GH, GW;                               // logical output grid extent (rows/cols of this launch's pixel grid)
  int ish, isw;                             // input coordinate = grid coordinate * is + d
  int ntaps;
  int dh[MAX_TAPS], dw[MAX_TAPS], widx[MAX_TAPS];
  int OH, OW, osh, osw, oph, opw;  // output tensor extent and the affine map grid -> output pixel
  const bf16* mask;                // optional: out = mask > 0 ? out * mask_scale : 0 (same layout as y)
  float mask_scale;
  int debug;  // OMR_CONV_DEBUG (diagnostics, results are WRONG when set): 1 = no MMAs, 2 = no TMA loads, 4 = no global stores
  long long* dbg;  // OMR_CONV_DEBUG & 8: per-tile clock64() stamps of CTA 0 (halo kernel), [role][tile][4]
};
#define DBG_STAMP(role, it, k)                                                     \
  do {                                                                             \
    if (g.dbg && blockIdx.x == 0 && (it) < 64) g.dbg[((role) * 64 + (it)) * 4 + (k)] = clock64(); \
  } while (0)

// RB: row bytes (= 2 * min(Cin, 64)); G: (tap, chunk) sub-tiles per pipeline stage
template <int RB, int G>
struct Cfg {
  static constexpr int A_SUB = 128 * RB;
  static constexpr int B_SUB = ((128 * RB) + 1023) / 1024 * 1024;  // room for Cout <= 128 rows, 1 KB aligned
  static constexpr int STAGE = G * (A_SUB + B_SUB);
  static constexpr int STAGES = (STAGE * 4 <= 160 * 1024) ? 4 : (STAGE * 3 <= 160 * 1024 ? 3 : 2);
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
};

template <int RB, int G>
__global__ void __launch_bounds__(192, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                         const __grid_constant__ CUtensorMap tmW, ConvTcArgs g) {
  omr_pdl_enter();
  using C = Cfg<RB, G>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int chunks = (g.Cin * 2 + RB - 1) / RB;  // 64-channel chunks per tap (1 or 2)
  const int nsub = g.ntaps * chunks;
  const int ngroups = (nsub + G - 1) / G;
  const uint32_t tmem_cols = g.Cout * 2 <= 32 ? 32 : (g.Cout * 2 <= 64 ? 64 : (g.Cout * 2 <= 128 ? 128 : 256));

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t a_box_bytes = (uint32_t)(g.TH * g.TW) * RB;
  const uint32_t b_box_bytes = (uint32_t)g.Cout * RB;

  if (warp == 4) {
    // ---- TMA producer: warp-uniform loop (see tc_common.cuh), one elected lane issues ----
    int s = 0;
    uint32_t ph = 1;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const int tw = tile % g.tiles_w;
      const int th = (tile / g.tiles_w) % g.tiles_h;
      const int n = tile / (g.tiles_w * g.tiles_h);
      const int oh0 = th * g.TH, ow0 = tw * g.TW;
      for (int grp = 0; grp < ngroups; ++grp) {
        mbar_wait(&empty_bar[s], ph);
        const int sub0 = grp * G;
        const int cnt = (nsub - sub0) < G ? (nsub - sub0) : G;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], (uint32_t)cnt * (a_box_bytes + b_box_bytes));
          uint8_t* stage = smem + s * C::STAGE;
          for (int q = 0; q < cnt; ++q) {
            const int sub = sub0 + q;
            const int tap = sub / chunks, ch = sub - tap * chunks;
            const int c0 = ch * (RB / 2);
            tma_load_4d(stage + q * C::A_SUB, &tmX, &full_bar[s], c0, ow0 * g.isw + g.dw[tap], oh0 * g.ish + g.dh[tap], n);
            tma_load_2d(stage + G * C::A_SUB + q * C::B_SUB, &tmW, &full_bar[s], g.widx[tap] * g.Cin + c0, 0);
          }
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    // ---- MMA issuer: warp-uniform loop, tcgen05.mma / commit under elect_one() ----
    const uint32_t idesc = make_idesc_bf16(128, g.Cout, 0, 0);
    const uint64_t d_hi = make_smem_desc(0, 16, 8 * RB, RB);
    const uint32_t smem_lo = smem_u32(smem) >> 4;
    uint32_t tcount = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait(&tempty_bar[a], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * (uint32_t)g.Cout;
      for (int grp = 0; grp < ngroups; ++grp) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const int sub0 = grp * G;
        const int cnt = (nsub - sub0) < G ? (nsub - sub0) : G;
        const uint32_t stage = smem_lo + (uint32_t)s * (C::STAGE >> 4);
        if (elect_one()) {
#pragma unroll
          for (int q = 0; q < G; ++q) {
            if (q < cnt) {
              const uint32_t a_addr = stage + q * (C::A_SUB >> 4);
              const uint32_t b_addr = stage + (G * C::A_SUB + q * C::B_SUB) / 16;
#pragma unroll
              for (int j = 0; j < RB / 32; ++j) {
                if (grp == 0 && q == 0 && j == 0)
                  umma_bf16_new(d_tmem, d_hi | (uint64_t)(a_addr + 2 * j), d_hi | (uint64_t)(b_addr + 2 * j), idesc);
                else
                  umma_bf16_acc(d_tmem, d_hi | (uint64_t)(a_addr + 2 * j), d_hi | (uint64_t)(b_addr + 2 * j), idesc);
              }
            }
          }
          umma_commit(&empty_bar[s]);
          if (grp == ngroups - 1) umma_commit(&tfull_bar[a]);
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ---- epilogue: thread = one output pixel; Cout channels from TMEM -> bias/ReLU -> bf16 NHWC ----
    uint32_t tcount = 0;
    const int r = warp * 32 + lane;  // tile row = pixel index inside the TH x TW patch
    const int pr = r / g.TW, pc = r - pr * g.TW;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
      const int tw = tile % g.tiles_w;
      const int th = (tile / g.tiles_w) % g.tiles_h;
      const int n = tile / (g.tiles_w * g.tiles_h);
      const int gh = th * g.TH + pr, gw = tw * g.TW + pc;
      const int oh = gh * g.osh + g.oph, ow = gw * g.osw + g.opw;
      const bool ok = pr < g.TH && gh < g.GH && gw < g.GW && oh < g.OH && ow < g.OW;
      bf16* dst = g.y + (((long long)n * g.OH + oh) * g.OW + ow) * g.Cout;
      mbar_wait(&tfull_bar[a], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + a * (uint32_t)g.Cout + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < g.Cout; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_addr + c0, v);
        tmem_ld_wait();
        if (ok) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            f[j] = __uint_as_float(v[j]);
            if (g.bias) f[j] += __ldg(g.bias + c0 + j);
            if (g.relu) f[j] = fmaxf(f[j], 0.f);
          }
          if (g.mask) {  // fused ReLU / dropout backward of the layer that produced this conv's forward input
            const uint4 m0 = reinterpret_cast<const uint4*>(g.mask + (dst - g.y) + c0)[0];
            const uint4 m1 = reinterpret_cast<const uint4*>(g.mask + (dst - g.y) + c0)[1];
            const __nv_bfloat162* mb0 = reinterpret_cast<const __nv_bfloat162*>(&m0);
            const __nv_bfloat162* mb1 = reinterpret_cast<const __nv_bfloat162*>(&m1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              f[2 * j] = __low2float(mb0[j]) > 0.f ? f[2 * j] * g.mask_scale : 0.f;
              f[2 * j + 1] = __high2float(mb0[j]) > 0.f ? f[2 * j + 1] * g.mask_scale : 0.f;
              f[8 + 2 * j] = __low2float(mb1[j]) > 0.f ? f[8 + 2 * j] * g.mask_scale : 0.f;
              f[8 + 2 * j + 1] = __high2float(mb1[j]) > 0.f ? f[8 + 2 * j + 1] * g.mask_scale : 0.f;
            }
          }
          uint4 o0, o1;
          o0.x = pack_bf16(f[0], f[1]); o0.y = pack_bf16(f[2], f[3]); o0.z = pack_bf16(f[4], f[5]); o0.w = pack_bf16(f[6], f[7]);
          o1.x = pack_bf16(f[8], f[9]); o1.y = pack_bf16(f[10], f[11]); o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
          uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
          d4[0] = o0;
          d4[1] = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Halo variant for stride-1 convolutions with C_in <= 64 (the high-resolution layers, which are bound by L2->SM
// traffic and TMA box rate, not by the tensor cores): per 128-pixel row tile ONE TMA box brings the 3 x (TW+2)
// input pixels, and the nine taps are nine UMMA descriptors whose start address is shifted by whole pixel rows
// ((dh+1)*(TW+2) + (dw+1) rows of RB bytes) inside that box -- the swizzle is a function of the shared-memory
// address, so a row-shifted window of a TMA-written tile is still a valid K-major operand.  All nine weight taps
// stay resident in shared memory for the life of the persistent CTA.  Input traffic per tile drops from 9 boxes
// to 1 (3.05x the output pixels instead of 9x) and the per-tile TMA operations from 18 to 1.
// ---------------------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(192, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap tmX,
                                                           const __grid_constant__ CUtensorMap tmW, ConvTcArgs g, int wsub,
                                                           int hsub, int stages) {
  omr_pdl_enter();
  // one tile = TH output rows x TW pixels of one image: ONE (TH+2) x (TW+2) TMA box, TH accumulators of [128 x Cout]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                // 9 taps x wsub
  uint8_t* sH = smem + 9 * wsub;     // stages x hsub
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + stages * hsub);
  uint64_t* w_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int TH = g.TH;
  const uint32_t acc_cols = (uint32_t)(TH * g.Cout);  // per accumulator buffer
  const uint32_t need = 2 * acc_cols;
  const uint32_t tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
  const int pitch = g.TW + 2;  // pixel rows per input image row inside the halo box

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);

  if (warp == 4) {
    // ---- TMA producer: the whole warp walks the tiles (uniform state), one elected lane issues ----
    if (elect_one()) {
      mbar_expect_tx(w_full, 9u * (uint32_t)g.Cout * RB);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * wsub, &tmW, w_full, t * g.Cin, 0);
    }
    uint32_t it = 0;
    int s = 0;
    uint32_t ph = 1;  // parity to wait for on empty_bar[s]
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const int tw = tile % g.tiles_w;
      const int th = (tile / g.tiles_w) % g.tiles_h;
      const int n = tile / (g.tiles_w * g.tiles_h);
      DBG_STAMP(0, it, 0);
      mbar_wait(&empty_bar[s], ph);
      DBG_STAMP(0, it, 1);
      if (elect_one()) {
        if (g.debug & 2) {
          mbar_arrive(&full_bar[s]);
        } else {
          mbar_expect_tx(&full_bar[s], (uint32_t)(TH + 2) * (uint32_t)pitch * RB);
          tma_load_4d(sH + s * hsub, &tmX, &full_bar[s], 0, tw * g.TW - 1, th * TH - 1, n);
        }
      }
      DBG_STAMP(0, it, 2);
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 5) {
    // ---- MMA issuer: warp-uniform loop, tcgen05.mma / commit under elect_one() ----
    const uint32_t idesc = make_idesc_bf16(128, g.Cout, 0, 0);
    const uint64_t d_hi = make_smem_desc(0, 16, 8 * RB, RB);
    // per-tap start offsets inside the halo (A) and the weight bank (B), in 16-byte units
    uint32_t ta[9], tb[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int tt = t < g.ntaps ? t : 0;
      ta[t] = (uint32_t)(((g.dh[tt] + 1) * pitch + (g.dw[tt] + 1)) * RB) >> 4;
      tb[t] = (uint32_t)(g.widx[tt] * wsub) >> 4;
    }
    const uint32_t row_step = (uint32_t)(pitch * RB) >> 4;
    const uint32_t cout = (uint32_t)g.Cout;
    const int ntaps = g.ntaps;
    const bool no_mma = (g.debug & 1) != 0;
    mbar_wait(w_full, 0);
    const uint32_t w_lo = smem_u32(sW) >> 4;
    const uint32_t h_base = smem_u32(sH) >> 4, h_step = (uint32_t)hsub >> 4;
    uint32_t it = 0;
    int s = 0;
    uint32_t ph = 0;  // parity to wait for on full_bar[s]
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t a = it & 1, aph = (it >> 1) & 1;
      DBG_STAMP(1, it, 0);
      mbar_wait(&tempty_bar[a], aph ^ 1);
      DBG_STAMP(1, it, 1);
      mbar_wait(&full_bar[s], ph);
      DBG_STAMP(1, it, 2);
      tc_fence_after();
      const uint32_t h_lo = h_base + (uint32_t)s * h_step;
      const uint32_t d0 = tmem_base + a * acc_cols;
      if (elect_one()) {
        if (!no_mma) {
          for (int r = 0; r < TH; ++r) {
            const uint32_t d_tmem = d0 + (uint32_t)r * cout;
            const uint32_t hr = h_lo + (uint32_t)r * row_step;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t < ntaps) {
#pragma unroll
                for (int j = 0; j < RB / 32; ++j) {
                  if (t > 0 || j > 0)
                    umma_bf16_acc(d_tmem, d_hi | (uint64_t)(hr + ta[t] + 2 * j), d_hi | (uint64_t)(w_lo + tb[t] + 2 * j), idesc);
                  else
                    umma_bf16_new(d_tmem, d_hi | (uint64_t)(hr + ta[t] + 2 * j), d_hi | (uint64_t)(w_lo + tb[t] + 2 * j), idesc);
                }
              }
            }
          }
        }
        umma_commit(&empty_bar[s]);
        umma_commit(&tfull_bar[a]);
      }
      __syncwarp();
      DBG_STAMP(1, it, 3);
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    uint32_t it = 0;
    const int px = warp * 32 + lane;  // pixel inside an output row of the tile
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t a = it & 1, aph = (it >> 1) & 1;
      const int tw = tile % g.tiles_w;
      const int th = (tile / g.tiles_w) % g.tiles_h;
      const int n = tile / (g.tiles_w * g.tiles_h);
      const int gw = tw * g.TW + px;
      if (threadIdx.x == 0) DBG_STAMP(2, it, 0);
      mbar_wait(&tfull_bar[a], aph);
      if (threadIdx.x == 0) DBG_STAMP(2, it, 1);
      tc_fence_after();
      for (int r = 0; r < TH; ++r) {
        const int gh = th * TH + r;
        const bool ok = px < g.TW && gw < g.GW && gh < g.GH;
        bf16* dst = g.y + (((long long)n * g.OH + gh) * g.OW + gw) * g.Cout;
        const uint32_t t_addr = tmem_base + a * acc_cols + (uint32_t)(r * g.Cout) + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < g.Cout; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + c0, v);
          tmem_ld_wait();
          if (ok) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              f[j] = __uint_as_float(v[j]);
              if (g.bias) f[j] += __ldg(g.bias + c0 + j);
              if (g.relu) f[j] = fmaxf(f[j], 0.f);
            }
            if (g.mask) {  // fused ReLU / dropout backward of the layer that produced this conv's forward input
              const uint4 m0 = reinterpret_cast<const uint4*>(g.mask + (dst - g.y) + c0)[0];
              const uint4 m1 = reinterpret_cast<const uint4*>(g.mask + (dst - g.y) + c0)[1];
              const __nv_bfloat162* mb0 = reinterpret_cast<const __nv_bfloat162*>(&m0);
              const __nv_bfloat162* mb1 = reinterpret_cast<const __nv_bfloat162*>(&m1);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                f[2 * j] = __low2float(mb0[j]) > 0.f ? f[2 * j] * g.mask_scale : 0.f;
                f[2 * j + 1] = __high2float(mb0[j]) > 0.f ? f[2 * j + 1] * g.mask_scale : 0.f;
                f[8 + 2 * j] = __low2float(mb1[j]) > 0.f ? f[8 + 2 * j] * g.mask_scale : 0.f;
                f[8 + 2 * j + 1] = __high2float(mb1[j]) > 0.f ? f[8 + 2 * j + 1] * g.mask_scale : 0.f;
              }
            }
            uint4 o0, o1;
            o0.x = pack_bf16(f[0], f[1]); o0.y = pack_bf16(f[2], f[3]); o0.z = pack_bf16(f[4], f[5]); o0.w = pack_bf16(f[6], f[7]);
            o1.x = pack_bf16(f[8], f[9]); o1.y = pack_bf16(f[10], f[11]); o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
            uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
            if (!(g.debug & 4)) {
              d4[0] = o0;
              d4[1] = o1;
            }
          }
        }
      }
      if (threadIdx.x == 0) DBG_STAMP(2, it, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
      if (threadIdx.x == 0) DBG_STAMP(2, it, 3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int g_conv_debug = -1;
int conv_debug() {
  if (g_conv_debug < 0) {
    const char* e = getenv("OMR_CONV_DEBUG");
    g_conv_debug = e ? atoi(e) : 0;
  }
  return g_conv_debug;
}

int g_halo_mode = -1;
bool halo_enabled() {
  if (g_halo_mode < 0) {
    const char* e = getenv("OMR_CONV_HALO");
    g_halo_mode = (e && e[0] == '0') ? 0 : 1;  // on unless explicitly disabled
  }
  return g_halo_mode == 1;
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int RB, int G>
int launch_cfg(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvTcArgs& a, cudaStream_t st) {
  auto kern = conv_tc_kernel<RB, G>;
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<RB, G>::SMEM));
    configured = true;
  }
  int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  OmrLaunch(grid, 192, Cfg<RB, G>::SMEM, st)(kern, tmX, tmW, a);
  OMR_LAUNCHED();
  return OMR_OK;
}

// One launch of the tap-GEMM.  x: [N, XH, XW, Cin] bf16; wpack: [Cout, 9*Cin] bf16 (tap-major, channels innermost).
int run_taps(const void* x, int N, int XH, int XW, int Cin, const void* wpack, int Cout, ConvTcArgs a, cudaStream_t st) {
  const int rb = (Cin >= 64 ? 64 : Cin) * 2;
  if (halo_enabled() && a.ish == 1 && a.isw == 1 && a.ntaps == 9 && Cin <= 64 && a.osh == 1 && a.osw == 1) {
    const int TW = a.GW >= 128 ? 128 : a.GW;
    const int pitch = TW + 2;
    const int wsub = (Cout * rb + 1023) / 1024 * 1024;
    // TH output rows per tile: fewer TMA rows per output pixel ((TH+2)/TH instead of 3) and fewer barrier round trips;
    // bounded by TMEM (2 buffers x TH x Cout columns <= 512) and by shared memory (>= 2 halo stages next to the weights)
    int TH = 4, hsub = 0, stages = 0;
    for (; TH >= 1; TH >>= 1) {
      if (2 * TH * Cout > 512 || (TH > 1 && a.GH < TH)) continue;
      int rows = (TH + 2) * pitch;
      if (rows < (TH + 1) * pitch + 2 + 128) rows = (TH + 1) * pitch + 2 + 128;
      hsub = (rows * rb + 1023) / 1024 * 1024;
      stages = (225 * 1024 - 1024 - 512 - 9 * wsub) / hsub;
      if (stages > 4) stages = 4;
      if (stages >= 2) break;
    }
    if (TH >= 1 && stages >= 2) {
      ConvTcArgs h = a;
      h.debug = conv_debug();
      static long long* dbg_dev = nullptr;
      if (h.debug & 8) {
        if (!dbg_dev) cudaMalloc(&dbg_dev, 3 * 64 * 4 * sizeof(long long));
        cudaMemsetAsync(dbg_dev, 0, 3 * 64 * 4 * sizeof(long long), st);
        h.dbg = dbg_dev;
      }
      h.TH = TH; h.TW = TW;
      h.tiles_w = (a.GW + TW - 1) / TW;
      h.tiles_h = (a.GH + TH - 1) / TH;
      h.num_tiles = a.N * h.tiles_h * h.tiles_w;
      h.Cin = Cin; h.Cout = Cout;
      CUtensorMap tmX, tmW;
      unsigned long long dims[4] = {(unsigned long long)Cin, (unsigned long long)XW, (unsigned long long)XH, (unsigned long long)N};
      unsigned long long strides[3] = {(unsigned long long)Cin * 2, (unsigned long long)XW * Cin * 2, (unsigned long long)XH * XW * Cin * 2};
      unsigned int box[4] = {(unsigned)Cin, (unsigned)pitch, (unsigned)(TH + 2), 1u};
      int rc = omr_make_tensor_map(&tmX, 2, x, 4, dims, strides, box, nullptr, rb);
      if (rc) return rc;
      unsigned long long wd[2] = {(unsigned long long)9 * Cin, (unsigned long long)Cout};
      unsigned long long ws[1] = {(unsigned long long)9 * Cin * 2};
      unsigned int wb[2] = {(unsigned)Cin, (unsigned)Cout};
      rc = omr_make_tensor_map(&tmW, 2, wpack, 2, wd, ws, wb, nullptr, rb);
      if (rc) return rc;
      const int smem_bytes = 9 * wsub + stages * hsub + 1024 + 512;
      const int grid = h.num_tiles < num_sms() ? h.num_tiles : num_sms();
      static bool cfgd[3] = {false, false, false};
      const int ci = rb == 32 ? 0 : (rb == 64 ? 1 : 2);
      if (!cfgd[ci]) {
        if (rb == 32) OMR_CUDA(cudaFuncSetAttribute(conv_halo_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (rb == 64) OMR_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (rb == 128) OMR_CUDA(cudaFuncSetAttribute(conv_halo_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cfgd[ci] = true;
      }
      if (rb == 32) OmrLaunch(grid, 192, smem_bytes, st)(conv_halo_kernel<32>, tmX, tmW, h, wsub, hsub, stages);
      else if (rb == 64) OmrLaunch(grid, 192, smem_bytes, st)(conv_halo_kernel<64>, tmX, tmW, h, wsub, hsub, stages);
      else OmrLaunch(grid, 192, smem_bytes, st)(conv_halo_kernel<128>, tmX, tmW, h, wsub, hsub, stages);
      OMR_LAUNCHED();
      if (h.dbg) {  // diagnostics only: synchronises
        static int dumps = 0;
        cudaStreamSynchronize(st);
        static long long hb[3 * 64 * 4];
        cudaMemcpy(hb, dbg_dev, sizeof(hb), cudaMemcpyDeviceToHost);
        if (dumps++ < 3) {
          const long long t0 = hb[(0 * 64 + 0) * 4 + 0];
          fprintf(stderr, "[conv_halo dbg] Cin %d Cout %d TH %d TW %d tiles %d grid %d stages %d (clk rel. to producer start)\n", Cin, Cout, TH, TW,
                  h.num_tiles, grid, stages);
          for (int it = 0; it < 24; ++it) {
            fprintf(stderr, " it%2d  TMA wait %6lld got %6lld issued %6lld | MMA wait_te %6lld got %6lld full %6lld commit %6lld | EPI wait %6lld got %6lld done %6lld arr %6lld\n",
                    it, hb[(0 * 64 + it) * 4 + 0] - t0, hb[(0 * 64 + it) * 4 + 1] - t0, hb[(0 * 64 + it) * 4 + 2] - t0,
                    hb[(1 * 64 + it) * 4 + 0] - t0, hb[(1 * 64 + it) * 4 + 1] - t0, hb[(1 * 64 + it) * 4 + 2] - t0, hb[(1 * 64 + it) * 4 + 3] - t0,
                    hb[(2 * 64 + it) * 4 + 0] - t0, hb[(2 * 64 + it) * 4 + 1] - t0, hb[(2 * 64 + it) * 4 + 2] - t0, hb[(2 * 64 + it) * 4 + 3] - t0);
          }
        }
      }
      return OMR_OK;
    }
  }
  // tile geometry: a TH x TW patch of the logical output grid, TH*TW <= 128
  int TW = a.GW >= 128 ? 128 : a.GW;
  int TH = 128 / TW;
  if (TH > a.GH) TH = a.GH;
  if (TH * a.ish > 256 || TW * a.isw > 256) return OMR_TC_NOT_ELIGIBLE;
  a.TH = TH; a.TW = TW;
  a.tiles_w = (a.GW + TW - 1) / TW;
  a.tiles_h = (a.GH + TH - 1) / TH;
  a.num_tiles = a.N * a.tiles_h * a.tiles_w;
  a.Cin = Cin; a.Cout = Cout;
  CUtensorMap tmX, tmW;
  {
    unsigned long long dims[4] = {(unsigned long long)Cin, (unsigned long long)XW, (unsigned long long)XH, (unsigned long long)N};
    unsigned long long strides[3] = {(unsigned long long)Cin * 2, (unsigned long long)XW * Cin * 2, (unsigned long long)XH * XW * Cin * 2};
    unsigned int box[4] = {(unsigned)(rb / 2), (unsigned)(TW * a.isw), (unsigned)(TH * a.ish), 1u};
    unsigned int es[4] = {1u, (unsigned)a.isw, (unsigned)a.ish, 1u};
    int rc = omr_make_tensor_map(&tmX, 2, x, 4, dims, strides, box, es, rb);
    if (rc) return rc;
    unsigned long long wd[2] = {(unsigned long long)9 * Cin, (unsigned long long)Cout};
    unsigned long long ws[1] = {(unsigned long long)9 * Cin * 2};
    unsigned int wb[2] = {(unsigned)(rb / 2), (unsigned)Cout};
    rc = omr_make_tensor_map(&tmW, 2, wpack, 2, wd, ws, wb, nullptr, rb);
    if (rc) return rc;
  }
  if (rb == 32) return launch_cfg<32, 3>(tmX, tmW, a, st);
  if (rb == 64) return launch_cfg<64, 3>(tmX, tmW, a, st);
  return launch_cfg<128, 1>(tmX, tmW, a, st);
}

bool shape_ok(int Ci, int Co) {
  auto okc = [](int c) { return c == 16 || c == 32 || c == 64 || c == 128; };
  return okc(Ci) && okc(Co);
}

}  // namespace

int omr_conv3x3_fwd_tc(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Ci, int Co,
                       int sh, int sw, int relu, cudaStream_t st) {
  if (!shape_ok(Ci, Co) || N < 1) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return OMR_TC_NOT_ELIGIBLE;
  ConvTcArgs a{};
  a.y = (bf16*)y; a.bias = bias; a.relu = relu; a.N = N;
  a.GH = (H + sh - 1) / sh; a.GW = (W + sw - 1) / sw;
  a.ish = sh; a.isw = sw;
  a.ntaps = 9;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      int t = kh * 3 + kw;
      a.dh[t] = kh - 1; a.dw[t] = kw - 1; a.widx[t] = t;
    }
  a.OH = a.GH; a.OW = a.GW; a.osh = 1; a.osw = 1; a.oph = 0; a.opw = 0;
  return run_taps(x, N, H, W, Ci, w, Co, a, st);
}

// dx[N,H,W,Ci] from dy[N,Ho,Wo,Co] and the transposed pack wT[Ci, 9*Co]:
//   dx[h,w,ci] = sum_{kh,kw,co} dy[(h+1-kh)/sh, (w+1-kw)/sw, co] * w[co,ci,kh,kw]   (only exact divisions)
int omr_conv3x3_dgrad_tc(const void* dy, const void* wT, void* dx, int N, int H, int W, int Ci, int Co, int sh, int sw,
                         const void* mask, float mask_scale, cudaStream_t st) {
  if (!shape_ok(Ci, Co) || N < 1 || sh > 2 || sw > 2) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(wT) & 15) || (reinterpret_cast<uintptr_t>(dx) & 15) ||
      (reinterpret_cast<uintptr_t>(mask) & 15))
    return OMR_TC_NOT_ELIGIBLE;
  const int Ho = (H + sh - 1) / sh, Wo = (W + sw - 1) / sw;
  for (int ph = 0; ph < sh; ++ph)
    for (int pw = 0; pw < sw; ++pw) {
      ConvTcArgs a{};
      a.y = (bf16*)dx; a.bias = nullptr; a.relu = 0; a.N = N;
      a.mask = (const bf16*)mask; a.mask_scale = mask_scale;
      a.GH = (H - ph + sh - 1) / sh; a.GW = (W - pw + sw - 1) / sw;  // pixels of this parity class
      if (a.GH <= 0 || a.GW <= 0) continue;
      a.ish = 1; a.isw = 1;
      a.ntaps = 0;
      for (int kh = 0; kh < 3; ++kh) {
        if ((ph + 1 - kh) % sh != 0) continue;
        for (int kw = 0; kw < 3; ++kw) {
          if ((pw + 1 - kw) % sw != 0) continue;
          int t = a.ntaps++;
          // h = i*sh + ph  ->  source row (h + 1 - kh) / sh = i + (ph + 1 - kh) / sh   (exact; may be -1 -> zero fill)
          a.dh[t] = (ph + 1 - kh) / sh; a.dw[t] = (pw + 1 - kw) / sw; a.widx[t] = kh * 3 + kw;
        }
      }
      a.OH = H; a.OW = W; a.osh = sh; a.osw = sw; a.oph = ph; a.opw = pw;
      int rc = run_taps(dy, N, Ho, Wo, Co, wT, Ci, a, st);
      if (rc) return rc;
    }
  return OMR_OK;
}
