// wgrad_tc.cu -- weight gradient of the 3x3 convolutions on the tcgen05 tensor cores.
//
//   dW[co, ci, kh, kw] (+)= sum over pixels p = (n, oh, ow) of  dY[p, co] * X[n, oh*sh + kh - 1, ow*sw + kw - 1, ci]
//
// i.e. nine GEMMs  [Co x P] * [P x Ci]  whose reduction dimension is the (huge) pixel axis.  Both operands are
// NHWC activations, so both are "MN-major" for the tensor core (the channel index is contiguous inside a pixel
// row); one shared-memory row = one pixel, fetched by 4-D TMA boxes -- the X box of tap (kh,kw) starts at the
// shifted coordinate, with zero fill for the padding and element strides for strided convolutions, and the dY
// box rows that fall outside the image are zero-filled, which makes every pixel tile a clean multiple of 16.
// Each CTA owns a contiguous range of pixel tiles and a group of taps (all 9 when 9*Ci fp32 columns fit in
// TMEM, else one kernel row of 3), accumulates them in TMEM across its whole range and finally adds its partial
// sums to the fp32 gradient in the parameter's own [Co,Ci,3,3] layout with atomics.
// Warp roles: 0-3 epilogue, 4 TMA producer, 5 MMA issuer + TMEM allocator.
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

struct WgradArgs {
  float* dw;
  float* ws;  // optional fp32 scratch [9][Co][Ci] (zeroed): partial sums go there with 16-byte vector reductions and a small
              // second kernel adds them into dw's [Co][Ci][3][3] layout -- for the wide layers the epilogue's scalar atomics
              // (one 4-byte reduction per element and CTA, 36 bytes apart) took ~60 % of the kernel (ncu, round 2)
  int N, Ci, Co;
  int TH, TW, KP;  // pixel tile (KP = TH*TW, multiple of 16)
  int tiles_h, tiles_w, num_tiles;
  int sh, sw;
  int taps_per_cta, tap_groups, ctas_per_group;
  int rba, rbb, chunks_a, chunks_b;  // row bytes / channel chunks of the dY and X operands
  int stage_bytes, stages;
  // halo mode (stride 1): ONE X box of (TH+2) x (TW+2) pixels per chunk and stage; the nine taps are row-shifted
  // descriptor windows into it (start row = (row + kh) * pitch + col + kw), so X is fetched ~1.6x instead of 9x
  int halo, pitch, halo_rows;
};

__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                          const __grid_constant__ CUtensorMap tmX, WgradArgs g) {
  omr_pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + g.stages * g.stage_bytes);
  uint64_t* empty_bar = full_bar + g.stages;
  uint64_t* accum_bar = empty_bar + g.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  uint32_t* s_aoff = tmem_slot + 2;  // [8]   A start-address offsets per k16 step (16-byte units)
  uint32_t* s_boff = s_aoff + 8;     // [72]  B start-address offsets per (MMA group, k16 step)

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int group = blockIdx.x / g.ctas_per_group;        // tap group
  const int member = blockIdx.x - group * g.ctas_per_group;
  const int tap0 = group * g.taps_per_cta;
  const int ntap = g.taps_per_cta;
  // contiguous range of pixel tiles of this CTA
  const int per = (g.num_tiles + g.ctas_per_group - 1) / g.ctas_per_group;
  const int t_begin = member * per;
  int t_end = t_begin + per;
  if (t_end > g.num_tiles) t_end = g.num_tiles;
  const int ntiles = t_end > t_begin ? t_end - t_begin : 0;
  const int cols = ntap * g.Ci;
  const uint32_t tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
  const int a_sub = g.KP * g.rba;  // bytes of one dY chunk tile
  const int b_sub = (g.halo ? g.halo_rows : g.KP) * g.rbb;  // bytes of one X (tap, chunk) tile / of one halo chunk
  const int a_bytes = g.chunks_a * a_sub;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);

  if (warp == 4) {
    {  // TMA producer: warp-uniform loop (see tc_common.cuh), one elected lane issues
      const uint32_t tx = g.halo ? (uint32_t)(a_bytes + g.chunks_b * (g.TH + 2) * g.pitch * g.rbb)
                                 : (uint32_t)(a_bytes + ntap * g.chunks_b * b_sub);
      int s = 0;
      uint32_t ph = 1;
      for (int i = 0; i < ntiles; ++i) {
        const int tile = t_begin + i;
        const int tw = tile % g.tiles_w;
        const int th = (tile / g.tiles_w) % g.tiles_h;
        const int n = tile / (g.tiles_w * g.tiles_h);
        const int oh0 = th * g.TH, ow0 = tw * g.TW;
        mbar_wait(&empty_bar[s], ph);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], tx);
          uint8_t* stage = smem + s * g.stage_bytes;
          for (int c = 0; c < g.chunks_a; ++c) tma_load_4d(stage + c * a_sub, &tmDY, &full_bar[s], c * 64, ow0, oh0, n);
          if (g.halo) {
            for (int c = 0; c < g.chunks_b; ++c) tma_load_4d(stage + a_bytes + c * b_sub, &tmX, &full_bar[s], c * 64, ow0 - 1, oh0 - 1, n);
          } else {
            for (int t = 0; t < ntap; ++t) {
              const int tap = tap0 + t, kh = tap / 3, kw = tap - kh * 3;
              for (int c = 0; c < g.chunks_b; ++c)
                tma_load_4d(stage + a_bytes + (t * g.chunks_b + c) * b_sub, &tmX, &full_bar[s], c * 64, ow0 * g.sw + kw - 1,
                            oh0 * g.sh + kh - 1, n);
            }
          }
        }
        __syncwarp();
        if (++s == g.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    // The MMA issuer is ONE thread: everything that can be hoisted out of its loop is.  All descriptors of a stage
    // differ only in their 14-bit start-address field, so the per-(group, k-step) address offsets (in 16-byte units)
    // are tabulated once by the whole warp and an MMA costs two table loads and two adds.
    // "group" = one tcgen05.mma target: with the halo layout and Ci <= 64 the three kw taps of a kernel row are ONE
    // MMA of N = 3*Ci (MN-major B whose three Ci-wide chunks are the same window shifted by one pixel row each:
    // leading byte offset = one row), otherwise one tap (N = Ci).
    const bool merge = g.halo && g.chunks_b == 1;
    const int ngroup = merge ? ntap / 3 : ntap;
    const int nj = g.KP / 16;  // <= 8
    {
      const uint32_t nB = merge ? 3 * g.Ci : g.Ci;
      // Co <= 64 -> M = 64: the MN-major A fetch costs one shared-memory wavefront per (k row, M chunk), and with 16/32
      // channel rows most chunks of an M = 128 tile would be aliases of the real one; M = 64 halves that traffic
      const uint32_t idesc = make_idesc_bf16(g.Co > 64 ? 128 : 64, nB, 1, 1);
      const uint32_t lbo_a = g.chunks_a > 1 ? (uint32_t)a_sub : 0u;  // Co <= 64: every 64-wide M chunk aliases the real one
      const uint32_t lbo_b = merge ? (uint32_t)g.rbb : (g.chunks_b > 1 ? (uint32_t)b_sub : 0u);
      const uint64_t a_hi = make_smem_desc(0, lbo_a, 8 * g.rba, g.rba);  // everything but the start address
      const uint64_t b_hi = make_smem_desc(0, lbo_b, 8 * g.rbb, g.rbb);
      // start-address offsets (16-byte units) kept in (uniform) registers: A per k16 step, B = per-step part + per-group part
      const uint32_t a_step = (uint32_t)(16 * g.rba) >> 4;
      uint32_t jb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (g.halo) {
          const int pr = (j * 16) / g.TW, pc = (j * 16) - pr * g.TW;
          jb[j] = (uint32_t)((pr * g.pitch + pc) * g.rbb) >> 4;
        } else {
          jb[j] = (uint32_t)(j * 16 * g.rbb) >> 4;
        }
      }
      uint32_t gb[9];
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        const int tap = tap0 + (merge ? 3 * m : m), kh = tap / 3, kw = tap - kh * 3;
        gb[m] = g.halo ? (uint32_t)((kh * g.pitch + kw) * g.rbb) >> 4 : (uint32_t)(m * g.chunks_b * b_sub) >> 4;
      }
      const uint32_t smem_lo = smem_u32(smem) >> 4, stage_step = (uint32_t)g.stage_bytes >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < ntiles; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_lo = smem_lo + (uint32_t)s * stage_step;
        const uint32_t b_lo = a_lo + ((uint32_t)a_bytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int m = 0; m < 9; ++m) {
            if (m < ngroup) {
              const uint32_t d = tmem_base + (uint32_t)m * nB;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nj) {
                  if (j > 0)
                    umma_bf16_acc(d, a_hi | (uint64_t)(a_lo + j * a_step), b_hi | (uint64_t)(b_lo + gb[m] + jb[j]), idesc);
                  else
                    umma_bf16(d, a_hi | (uint64_t)(a_lo), b_hi | (uint64_t)(b_lo + gb[m] + jb[0]), idesc, i > 0 ? 1u : 0u);
                }
            }
          }
          umma_commit(&empty_bar[s]);
          if (i == ntiles - 1) umma_commit(accum_bar);
        }
        __syncwarp();
        if (++s == g.stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (ntiles == 0 && elect_one()) umma_commit(accum_bar);
    }
  } else if (ntiles > 0) {
    // ---- epilogue: thread = output channel co; add this CTA's partial sums into dW[co, ci, kh, kw] ----
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    // accumulator rows: M = 128 -> lane = row; M = 64 -> row r lives in lane (r % 16) + 32 * (r / 16) (16 per warp)
    const bool m64 = g.Co <= 64;
    const int co = m64 ? warp * 16 + lane : warp * 32 + lane;
    const bool ok = co < g.Co && (!m64 || lane < 16);
    for (int t = 0; t < ntap; ++t) {
      const int tap = tap0 + t;
      for (int c0 = 0; c0 < g.Ci; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * g.Ci + c0), v);
        tmem_ld_wait();
        if (ok) {
          if (g.ws) {
            float* dst = g.ws + ((long long)tap * g.Co + co) * g.Ci + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
          } else {
            float* dst = g.dw + ((long long)co * g.Ci + c0) * 9 + tap;
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + j * 9, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dw[co][ci][tap] += ws[tap][co][ci]
__global__ void wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Co, int Ci) {
  omr_pdl_enter();
  const int total = Co * Ci * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, cc = i / 9;  // cc = co * Ci + ci
    dw[i] += ws[(long long)tap * Co * Ci + cc];
  }
}

int g_wgrad_halo = -1;
bool wgrad_halo_enabled() {
  if (g_wgrad_halo < 0) {
    const char* e = getenv("OMR_WGRAD_HALO");
    g_wgrad_halo = (e && e[0] == '0') ? 0 : 1;  // on unless explicitly disabled
  }
  return g_wgrad_halo == 1;
}
int g_sms = 0;
int sms() {
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

}  // namespace

int omr_conv3x3_wgrad_tc(const void* x, const void* dy, float* dw, int N, int H, int W, int Ci, int Co, int sh, int sw,
                         int accumulate, float* ws, cudaStream_t st) {
  auto okc = [](int c) { return c == 16 || c == 32 || c == 64 || c == 128; };
  if (!okc(Ci) || !okc(Co) || N < 1) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15)) return OMR_TC_NOT_ELIGIBLE;
  const int Ho = (H + sh - 1) / sh, Wo = (W + sw - 1) / sw;
  WgradArgs g{};
  g.dw = dw; g.N = N; g.Ci = Ci; g.Co = Co; g.sh = sh; g.sw = sw;
  // the scratch path pays two extra small launches: worth it where the epilogue's atomics dominate (>= 64 x 64 channels)
  g.ws = (ws && Ci * Co >= 64 * 64 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) ? ws : nullptr;
  g.halo = (sh == 1 && sw == 1 && wgrad_halo_enabled()) ? 1 : 0;
  g.KP = (Ci == 16 || g.halo) ? 128 : 64;
  int TW = ((Wo + 15) / 16) * 16;
  if (TW > g.KP) TW = g.KP;
  if (g.halo && TW > 32 && Ho >= 4) TW = 32;  // 4 x 32 patches: the (TH+2) x (TW+2) halo is 1.6x the patch, not 3x
  g.TW = TW; g.TH = g.KP / TW;
  if (g.KP % TW != 0) {  // TW in {16,32,48,64,...}: keep TH*TW a multiple of 16 that fits
    g.TH = g.KP / TW;
    g.KP = g.TH * TW;
  }
  g.pitch = g.TW + 2;
  g.halo_rows = (g.TH + 2) * g.pitch;
  // a shifted 16-row window may start up to 2*pitch+2 rows into the box: it always ends inside it
  // ((TH-1+2)*pitch + (TW-16) + 2 + 16 <= (TH+2)*pitch)
  if (g.TH * sh > 256 || g.TW * sw > 256) return OMR_TC_NOT_ELIGIBLE;
  g.tiles_w = (Wo + g.TW - 1) / g.TW;
  g.tiles_h = (Ho + g.TH - 1) / g.TH;
  g.num_tiles = N * g.tiles_h * g.tiles_w;
  g.taps_per_cta = (9 * Ci <= 512) ? 9 : 3;
  g.tap_groups = 9 / g.taps_per_cta;
  g.ctas_per_group = sms() / g.tap_groups;
  if (g.ctas_per_group > g.num_tiles) g.ctas_per_group = g.num_tiles;
  g.rba = (Co >= 64 ? 64 : Co) * 2; g.chunks_a = Co > 64 ? 2 : 1;
  g.rbb = (Ci >= 64 ? 64 : Ci) * 2; g.chunks_b = Ci > 64 ? 2 : 1;
  const int raw = g.halo ? g.KP * g.rba * g.chunks_a + ((g.halo_rows * g.rbb + 1023) / 1024 * 1024) * g.chunks_b
                         : g.KP * (g.rba * g.chunks_a + g.taps_per_cta * g.rbb * g.chunks_b);
  if (g.halo) g.halo_rows = ((g.halo_rows * g.rbb + 1023) / 1024 * 1024) / g.rbb;  // chunk tiles stay 1 KB aligned
  g.stage_bytes = (raw + 1023) / 1024 * 1024;
  g.stages = (200 * 1024) / g.stage_bytes;
  // narrow channel rows make every TMA box hundreds of 32/64-byte requests: the pipeline is bound by their latency,
  // so it runs as deep as shared memory allows
  static int max_stages = 0;
  if (!max_stages) {
    const char* e = getenv("OMR_WGRAD_STAGES");
    max_stages = e ? atoi(e) : 12;
    if (max_stages < 2) max_stages = 2;
    if (max_stages > 16) max_stages = 16;
  }
  if (g.stages > max_stages) g.stages = max_stages;
  if (g.stages < 2) return OMR_TC_NOT_ELIGIBLE;
  // every sub-tile must start on a swizzle-atom boundary (8 rows): KP is a multiple of 16 rows, so a_sub/b_sub are
  // multiples of 16 * rb >= 512 B; the 128 B swizzle needs 1024 B: 16 rows * 128 B = 2048 ok, 64 B: 16*64 = 1024 ok,
  // 32 B: atom is 256 B, 16*32 = 512 ok.
  const int smem_bytes = g.stages * g.stage_bytes + 1024 + 1024;

  CUtensorMap tmDY, tmX;
  {
    unsigned long long dims[4] = {(unsigned long long)Co, (unsigned long long)Wo, (unsigned long long)Ho, (unsigned long long)N};
    unsigned long long strides[3] = {(unsigned long long)Co * 2, (unsigned long long)Wo * Co * 2, (unsigned long long)Ho * Wo * Co * 2};
    unsigned int box[4] = {(unsigned)(g.rba / 2), (unsigned)g.TW, (unsigned)g.TH, 1u};
    int rc = omr_make_tensor_map(&tmDY, 2, dy, 4, dims, strides, box, nullptr, g.rba);
    if (rc) return rc;
    unsigned long long xd[4] = {(unsigned long long)Ci, (unsigned long long)W, (unsigned long long)H, (unsigned long long)N};
    unsigned long long xs[3] = {(unsigned long long)Ci * 2, (unsigned long long)W * Ci * 2, (unsigned long long)H * W * Ci * 2};
    unsigned int xb[4] = {(unsigned)(g.rbb / 2), (unsigned)(g.halo ? g.pitch : g.TW * sw), (unsigned)(g.halo ? g.TH + 2 : g.TH * sh), 1u};
    unsigned int es[4] = {1u, (unsigned)sw, (unsigned)sh, 1u};
    rc = omr_make_tensor_map(&tmX, 2, x, 4, xd, xs, xb, es, g.rbb);
    if (rc) return rc;
  }
  if (!accumulate) OMR_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Co * Ci * 9, st));
  if (g.ws) OMR_CUDA(cudaMemsetAsync(g.ws, 0, sizeof(float) * (size_t)Co * Ci * 9, st));
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  OmrLaunch(g.ctas_per_group * g.tap_groups, 192, smem_bytes, st)(wgrad_tc_kernel, tmDY, tmX, g);
  OMR_LAUNCHED();
  if (g.ws) {
    OmrLaunch((Co * Ci * 9 + 255) / 256, 256, 0, st)(wgrad_finalize_kernel, (const float*)g.ws, dw, Co, Ci);
    OMR_LAUNCHED();
  }
  return OMR_OK;
}
