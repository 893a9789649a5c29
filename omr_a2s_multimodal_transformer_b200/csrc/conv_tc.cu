// conv_tc.cu -- 3x3 convolution (forward and data gradient) as an implicit GEMM on the tcgen05 tensor cores.
//
//   D[pixel, co] = sum over taps t, channels c of  X[n, oh*ish + dh_t, ow*isw + dw_t, c] * Wt[co, widx_t, c]
//
// * activations are NHWC bf16; the A operand of tap t is fetched by 4-D TMA boxes whose start coordinate carries the
//   tap shift -- out-of-range coordinates (the zero padding) are filled with zeros by the TMA unit and strided
//   convolutions use the tensor map's element strides, so no thread ever computes an im2col address;
// * one smem row = one pixel = min(C,64) channels (32/64/128 bytes) in the TMA swizzle of that width, which is
//   exactly the K-major UMMA operand layout; the weight slice [Cout x chunk] of the tap is the B operand;
// * persistent CTAs (one per SM), each owning a CONTIGUOUS range of output tiles; the fp32 accumulator lives in TMEM and
//   is double buffered, so the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1;
// * the same kernels compute the data gradient: stride 1 -> taps mirrored, weights from the [Ci,3,3,Co] pack;
//   stride 2 -> one launch per output parity class (1, 2, 2 or 4 taps each), written with a strided epilogue.
//
// Warp roles (320 threads): 0-7 epilogue, 8 TMA producer, 9 MMA issuer + TMEM allocator.  Producer and issuer run their
// loops warp-wide on uniform values with the instruction under elect_one() (tc_common.cuh).  The epilogue was the
// bottleneck of the narrow (C = 16 / 32) layers when it had ONE warp per SM sub-partition (measured: ~2100 clk per 512-pixel
// tile, instruction-latency bound): it now has two warps per sub-partition -- warps w and w + 4 share the TMEM lanes
// 32 (w & 3) .. + 31 and split a tile's (row, 16-channel chunk) units between them -- batches its TMEM loads, takes the
// bias from shared memory and does ReLU / masking on packed bf16 pairs.
//
// Output path: the epilogue packs its bf16 results into a shared-memory staging tile laid out [out row][out col][channel] in
// the TMA swizzle of the pixel width (conflict-free 16-byte stores: a lane = a pixel = a staging row), and one elected thread
// writes the tile with a TMA store (cp.async.bulk.tensor ... bulk_group; edge tiles are clipped by the TMA unit).  Direct
// per-lane 32-byte global stores touched one 128-byte line per lane and instruction for C >= 32.  Two staging buffers
// alternate when shared memory allows.  (Strided parity-class launches of the generic kernel keep direct stores.)
//
// Fused per-channel reductions in the epilogue (MODE, output channels <= 64), accumulated per thread over the CTA's tile
// range and flushed with one atomic per (warp, channel) when the sample changes:
//   MODE 1  InstanceNorm statistics of the stored output: (sum y, sum y^2) per (n, c)     [conv2 of a block, forward]
//   MODE 2  column sums of the stored output per c = bias gradient of the producing layer  [data gradient with fused mask]
//   MODE 3  InstanceNorm backward sums (sum dy, sum dy * x) per (n, c), x = the norm's input [data gradient of conv3]
#define OMR_HAVE_TC_CONV 1
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int MAX_TAPS = 9;
constexpr int NTHREADS = 320;
constexpr int EPI_WARPS = 8, W_TMA = 8, W_MMA = 9;

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
  // halo kernel only: output classes.  A stride-2 data gradient is ONE launch over the dy grid: grid pixel (i, j) produces the
  // osh x osw output pixels (osh i + ph, osw j + pw), one accumulator per class (ph, pw); tap t feeds class tap_cls[t] and
  // tap_first[t] marks the first tap of its class.  The halo box starts at grid offset (box_h0, box_w0) and is
  // (TH + box_hr) x (TW + box_wr) pixels.  Plain stride-1 convolution: ncls = 1, box = (-1, -1, 2, 2).
  int ncls, tap_cls[MAX_TAPS], tap_first[MAX_TAPS], cls_ph[4], cls_pw[4];
  int box_h0, box_w0, box_hr, box_wr;
  const bf16* mask;                // optional: out = mask > 0 ? out * mask_scale : 0 (same layout as y)
  float mask_scale;
  // fused reductions (see the header): exactly one of them is non-null in MODE 1 / 2 / 3
  double* in_sums;     // MODE 1: [N][Cout][2] += (sum y, sum y^2)
  float* colsum;       // MODE 2: [Cout] += sum y
  const bf16* in_x;    // MODE 3: the InstanceNorm input (same layout as y) ...
  double* in_bsums;    //         ... and [N][Cout][2] += (sum y, sum y * x)
  // staged output: staging tile of st_rows x st_cols output pixels (0 = direct global stores), st_bufs buffers of st_bytes
  int st_rows, st_cols, st_bufs, st_bytes;
  long long* dbg;      // OMR_CONV_DEBUG & 8: clock64() stamps of CTA 0 / epilogue warp 0 (halo kernel): [tile][8]
  int debug;           // OMR_CONV_DEBUG (diagnostics, results are WRONG when set): 1 = no MMAs, 2 = no TMA loads, 4 = no global stores
};

#define DBG_STAMP(it, k)                                                                                   \
  do {                                                                                                     \
    if (g.dbg && blockIdx.x == 0 && threadIdx.x == 0 && (it) < 64) g.dbg[(it) * 8 + (k)] = clock64(); \
  } while (0)

// ---- epilogue building blocks ---------------------------------------------------------------------------------------
// contiguous, balanced tile range of this CTA
__device__ __forceinline__ void tile_range(int num_tiles, int& t0, int& cnt) {
  const int base = num_tiles / (int)gridDim.x, rem = num_tiles % (int)gridDim.x;
  const int b = (int)blockIdx.x;
  t0 = b * base + (b < rem ? b : rem);
  cnt = base + (b < rem ? 1 : 0);
}

// One (pixel, 16-channel chunk) unit: fp32 accumulators -> (+bias, ReLU | mask) -> bf16 -> global, plus the fused sums.
// a0 / a1: this chunk's 16 per-thread accumulators (MODE != 0).
// sdst: the pixel's row in the staging tile (nullptr: store to global), c0: first channel of the unit.
template <int MODE>
__device__ __forceinline__ void epi_unit(const ConvTcArgs& g, const uint32_t (&v)[16], long long off, const float* sbias, bool ok,
                                         float* a0, float* a1, uint8_t* sdst, int spix, int c0) {
  if (!ok) return;
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  if (g.bias) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b4 = reinterpret_cast<const float4*>(sbias)[q];
      f[4 * q] += b4.x; f[4 * q + 1] += b4.y; f[4 * q + 2] += b4.z; f[4 * q + 3] += b4.w;
    }
  }
  __nv_bfloat162 h[8];
  if (g.mask) {  // fused ReLU / dropout backward of the layer that produced this conv's forward input
    const uint4 m0 = reinterpret_cast<const uint4*>(g.mask + off)[0];
    const uint4 m1 = reinterpret_cast<const uint4*>(g.mask + off)[1];
    const __nv_bfloat162* mb0 = reinterpret_cast<const __nv_bfloat162*>(&m0);
    const __nv_bfloat162* mb1 = reinterpret_cast<const __nv_bfloat162*>(&m1);
    const __nv_bfloat162 zero = __float2bfloat162_rn(0.f);
    const float sc = g.mask_scale;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __hmul2(__floats2bfloat162_rn(f[2 * j] * sc, f[2 * j + 1] * sc), __hgt2(mb0[j], zero));
      h[4 + j] = __hmul2(__floats2bfloat162_rn(f[8 + 2 * j] * sc, f[8 + 2 * j + 1] * sc), __hgt2(mb1[j], zero));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    if (g.relu) {
      const __nv_bfloat162 zero = __float2bfloat162_rn(0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __hmax2(h[j], zero);
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float lo = __low2float(h[j]), hi = __high2float(h[j]);
      a0[2 * j] += lo; a0[2 * j + 1] += hi;
      a1[2 * j] = fmaf(lo, lo, a1[2 * j]); a1[2 * j + 1] = fmaf(hi, hi, a1[2 * j + 1]);
    }
  } else if (MODE == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[2 * j] += __low2float(h[j]);
      a0[2 * j + 1] += __high2float(h[j]);
    }
  } else if (MODE == 3) {
    const uint4 x0 = reinterpret_cast<const uint4*>(g.in_x + off)[0];
    const uint4 x1 = reinterpret_cast<const uint4*>(g.in_x + off)[1];
    const __nv_bfloat162* xb0 = reinterpret_cast<const __nv_bfloat162*>(&x0);
    const __nv_bfloat162* xb1 = reinterpret_cast<const __nv_bfloat162*>(&x1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 xv = j < 4 ? xb0[j] : xb1[j - 4];
      const float lo = __low2float(h[j]), hi = __high2float(h[j]);
      a0[2 * j] += lo; a0[2 * j + 1] += hi;
      a1[2 * j] = fmaf(lo, __low2float(xv), a1[2 * j]); a1[2 * j + 1] = fmaf(hi, __high2float(xv), a1[2 * j + 1]);
    }
  }
  if (sdst) {
    // staging row = min(Cout, 64) channels (rb bytes) in the TMA swizzle of that width: 16-byte chunk index XOR the
    // address bits [7, 7 + log2(rb / 16)); Cout = 128: two sub-tiles of 64 channels
    const int rb = g.Cout >= 64 ? 128 : g.Cout * 2;
    const int key = rb == 128 ? (spix & 7) : (rb == 64 ? ((spix >> 1) & 3) : ((spix >> 2) & 1));
    const int cc = (c0 & 63) >> 3;
    uint8_t* row = sdst + (c0 >> 6) * (g.st_bytes >> 1) + (long long)spix * rb;
    *reinterpret_cast<uint4*>(row + ((cc ^ key) << 4)) = *reinterpret_cast<const uint4*>(&h[0]);
    *reinterpret_cast<uint4*>(row + (((cc + 1) ^ key) << 4)) = *reinterpret_cast<const uint4*>(&h[4]);
  } else if (!(g.debug & 4)) {
    uint4* d4 = reinterpret_cast<uint4*>(g.y + off);
    d4[0] = *reinterpret_cast<const uint4*>(&h[0]);
    d4[1] = *reinterpret_cast<const uint4*>(&h[4]);
  }
}

// add this warp's per-thread partial sums of sample n to the global accumulators and clear them
template <int MODE>
__device__ __forceinline__ void epi_flush(const ConvTcArgs& g, int n, int hf, int nchunk, int lane, float (&a0)[32], float (&a1)[32]) {
  if (MODE == 0) return;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int chunk = nchunk == 1 ? 0 : hf + 2 * k;
    if (chunk < nchunk && (nchunk > 1 || k == 0)) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s0 = warp_sum(a0[k * 16 + j]);
        const float s1 = (MODE == 2) ? 0.f : warp_sum(a1[k * 16 + j]);
        if (lane == j) {
          const int c = chunk * 16 + j;
          if (MODE == 1) {
            atomicAdd(g.in_sums + ((long long)n * g.Cout + c) * 2, (double)s0);
            atomicAdd(g.in_sums + ((long long)n * g.Cout + c) * 2 + 1, (double)s1);
          } else if (MODE == 2) {
            atomicAdd(g.colsum + c, s0);
          } else {
            atomicAdd(g.in_bsums + ((long long)n * g.Cout + c) * 2, (double)s0);
            atomicAdd(g.in_bsums + ((long long)n * g.Cout + c) * 2 + 1, (double)s1);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      a0[k * 16 + j] = 0.f;
      a1[k * 16 + j] = 0.f;
    }
  }
}

// The epilogue of one tile: `rows` accumulator rows of 128 pixels x Cout channels in TMEM at t_base (+ r * Cout).
// Warp (q = warp & 3, hf = warp >> 2), lane -> pixel column px = 32 q + lane of every row; the (row, chunk) units are dealt
// to the two halves: Cout = 16 -> rows alternate, Cout >= 32 -> chunk parity = hf (so a thread's accumulators belong to fixed
// channels).  pix(r, ok) gives the element offset of (row r, this thread's pixel, channel 0) in y and whether it exists.
template <int MODE, typename PixFn>
__device__ __forceinline__ void epi_tile(const ConvTcArgs& g, uint32_t t_base, int rows, int q, int hf, const float* sbias, PixFn pix,
                                         float (&a0)[32], float (&a1)[32], uint8_t* sdst) {
  const int nchunk = g.Cout >> 4;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  if (nchunk == 1) {
    for (int r = hf; r < rows; r += 4) {  // two rows in flight (r, r + 2)
      uint32_t v0[16], v1[16];
      const bool two = r + 2 < rows;
      tmem_ld16(t_base + lane_addr + (uint32_t)(r * 16), v0);
      if (two) tmem_ld16(t_base + lane_addr + (uint32_t)((r + 2) * 16), v1);
      tmem_ld_wait();
      bool ok;
      int spix;
      long long off = pix(r, ok, spix);
      epi_unit<MODE>(g, v0, off, sbias, ok, a0, a1, sdst, spix, 0);
      if (two) {
        off = pix(r + 2, ok, spix);
        epi_unit<MODE>(g, v1, off, sbias, ok, a0, a1, sdst, spix, 0);
      }
    }
  } else {
    for (int r = 0; r < rows; ++r) {
      bool ok;
      int spix;
      const long long off = pix(r, ok, spix);
      const uint32_t t_row = t_base + lane_addr + (uint32_t)(r * g.Cout);
#pragma unroll
      for (int k = 0; k < 4; k += 2) {  // chunks hf + 2k and hf + 2k + 2 in flight
        const int c0 = (hf + 2 * k) * 16;
        if (hf + 2 * k < nchunk) {
          const bool two = hf + 2 * k + 2 < nchunk;
          uint32_t v0[16], v1[16];
          tmem_ld16(t_row + (uint32_t)c0, v0);
          if (two) tmem_ld16(t_row + (uint32_t)(c0 + 32), v1);
          tmem_ld_wait();
          // per-thread accumulators exist for a warp's first two chunks only (MODE != 0 requires Cout <= 64: k = 0)
          epi_unit<MODE>(g, v0, off + c0, sbias + c0, ok, &a0[0], &a1[0], sdst, spix, c0);
          if (two) epi_unit<MODE>(g, v1, off + c0 + 32, sbias + c0 + 32, ok, &a0[16], &a1[16], sdst, spix, c0 + 32);
        }
      }
    }
  }
}

// staging buffer hand-over among the 8 epilogue warps (256 threads, named barrier 1); thread 0 issues and tracks the stores
__device__ __forceinline__ void stage_acquire(const ConvTcArgs& g) {
  if (threadIdx.x == 0) {
    if (g.st_bufs > 1) tma_store_wait_read<1>();
    else tma_store_wait_read<0>();
  }
  named_barrier(1, EPI_WARPS * 32);
}
__device__ __forceinline__ void stage_store(const ConvTcArgs& g, const CUtensorMap* tmY, const uint8_t* buf, int ow0, int oh0, int n) {
  fence_proxy_async();
  named_barrier(1, EPI_WARPS * 32);
  if (threadIdx.x == 0) {
    if (!(g.debug & 4)) {
      tma_store_4d(tmY, buf, 0, ow0, oh0, n);
      if (g.Cout > 64) tma_store_4d(tmY, buf + (g.st_bytes >> 1), 64, ow0, oh0, n);
    }
    tma_store_commit();
  }
}

// RB: row bytes (= 2 * min(Cin, 64)); G: (tap, chunk) sub-tiles per pipeline stage
template <int RB, int G>
struct Cfg {
  static constexpr int A_SUB = 128 * RB;
  static constexpr int B_SUB = ((128 * RB) + 1023) / 1024 * 1024;  // room for Cout <= 128 rows, 1 KB aligned
  static constexpr int STAGE = G * (A_SUB + B_SUB);
  static constexpr int STAGES = (STAGE * 4 <= 160 * 1024) ? 4 : (STAGE * 3 <= 160 * 1024 ? 3 : 2);
  static constexpr int OUT_STAGE = 2 * 128 * 128 * 2;  // two staging tiles of 128 pixels x <= 128 channels
  static constexpr int SMEM = STAGES * STAGE + OUT_STAGE + 1024 + 1024;
};

template <int RB, int G, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                              const __grid_constant__ CUtensorMap tmW,
                                                              const __grid_constant__ CUtensorMap tmY, ConvTcArgs g) {
  omr_pdl_enter();
  using C = Cfg<RB, G>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOut = smem + STAGES * C::STAGE;  // staging tiles (1 KB aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOut + C::OUT_STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));  // [128], 16-byte aligned

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int chunks = (g.Cin * 2 + RB - 1) / RB;  // 64-channel chunks per tap (1 or 2)
  const int nsub = g.ntaps * chunks;
  const int ngroups = (nsub + G - 1) / G;
  const uint32_t tmem_cols = g.Cout * 2 <= 32 ? 32 : (g.Cout * 2 <= 64 ? 64 : (g.Cout * 2 <= 128 ? 128 : 256));
  int t0, tcnt;
  tile_range(g.num_tiles, t0, tcnt);

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    if (g.st_rows) tma_prefetch_desc(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, tmem_cols);
  if (threadIdx.x < 128) sbias[threadIdx.x] = (g.bias && (int)threadIdx.x < g.Cout) ? g.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t a_box_bytes = (uint32_t)(g.TH * g.TW) * RB;
  const uint32_t b_box_bytes = (uint32_t)g.Cout * RB;

  if (warp == W_TMA) {
    // ---- TMA producer: warp-uniform loop, one elected lane issues ----
    int s = 0;
    uint32_t ph = 1;
    for (int i = 0; i < tcnt; ++i) {
      const int tile = t0 + i;
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
  } else if (warp == W_MMA) {
    // ---- MMA issuer: warp-uniform loop, tcgen05.mma / commit under elect_one() ----
    const uint32_t idesc = make_idesc_bf16(128, g.Cout, 0, 0);
    const uint64_t d_hi = make_smem_desc(0, 16, 8 * RB, RB);
    const uint32_t smem_lo = smem_u32(smem) >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < tcnt; ++i) {
      const uint32_t a = i & 1, aph = (i >> 1) & 1;
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
    // ---- epilogue: thread = one output pixel of the TH x TW patch, its Cout channels split between warps w and w + 4 ----
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;  // tile row = pixel index inside the patch
    const int pr = r / g.TW, pc = r - pr * g.TW;
    float a0[32], a1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      a0[j] = 0.f;
      a1[j] = 0.f;
    }
    int cur_n = -1;
    for (int i = 0; i < tcnt; ++i) {
      const uint32_t a = i & 1, aph = (i >> 1) & 1;
      const int tile = t0 + i;
      const int tw = tile % g.tiles_w;
      const int th = (tile / g.tiles_w) % g.tiles_h;
      const int n = tile / (g.tiles_w * g.tiles_h);
      if (MODE != 0 && MODE != 2 && n != cur_n) {
        if (cur_n >= 0) epi_flush<MODE>(g, cur_n, hf, g.Cout >> 4, lane, a0, a1);
        cur_n = n;
      }
      const int gh = th * g.TH + pr, gw = tw * g.TW + pc;
      const int oh = gh * g.osh + g.oph, ow = gw * g.osw + g.opw;
      const bool okp = pr < g.TH && gh < g.GH && gw < g.GW && oh < g.OH && ow < g.OW;
      const long long offp = (((long long)n * g.OH + oh) * g.OW + ow) * g.Cout;
      uint8_t* sbuf = g.st_rows ? sOut + (i % g.st_bufs) * g.st_bytes : nullptr;
      if (g.st_rows) stage_acquire(g);
      mbar_wait(&tfull_bar[a], aph);
      tc_fence_after();
      // the patch is ONE accumulator row block: with Cout = 16 only the hf = 0 warps have a unit (rows = 1); the staging
      // tile is the patch itself (pixel r of the patch = staging pixel r)
      epi_tile<MODE>(g, tmem_base + a * (uint32_t)g.Cout, 1, q, hf, sbias,
                     [&](int, bool& ok, int& spix) {
                       ok = okp;
                       spix = r;
                       return offp;
                     },
                     a0, a1, sbuf);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
      if (g.st_rows) stage_store(g, &tmY, sbuf, tw * g.TW, th * g.TH, n);
    }
    if (MODE != 0 && (MODE == 2 || cur_n >= 0)) epi_flush<MODE>(g, cur_n, hf, g.Cout >> 4, lane, a0, a1);
    if (g.st_rows && threadIdx.x == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Halo variant for stride-1 convolutions with C_in <= 64 (the high-resolution layers, which are bound by HBM / L2->SM
// traffic, not by the tensor cores): per tile of TH output rows x TW pixels ONE TMA box brings the (TH+2) x (TW+2)
// input pixels, and the nine taps are nine UMMA descriptors whose start address is shifted by whole pixel rows
// ((dh+1)*(TW+2) + (dw+1) rows of RB bytes) inside that box -- the swizzle is a function of the shared-memory
// address, so a row-shifted window of a TMA-written tile is still a valid K-major operand.  All nine weight taps
// stay resident in shared memory for the life of the persistent CTA.
// ---------------------------------------------------------------------------------------------------------------
template <int RB, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                const __grid_constant__ CUtensorMap tmW,
                                                                const __grid_constant__ CUtensorMap tmY, ConvTcArgs g, int wsub,
                                                                int hsub, int stages) {
  omr_pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                // 9 taps x wsub
  uint8_t* sH = smem + 9 * wsub;     // stages x hsub
  uint8_t* sOut = sH + stages * hsub;  // st_bufs x st_bytes staging tiles (1 KB aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + g.st_bufs * g.st_bytes);
  uint64_t* w_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));  // [128], 16-byte aligned

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int TH = g.TH;
  const uint32_t acc_cols = (uint32_t)(TH * g.ncls * g.Cout);  // per accumulator buffer: one [128 x Cout] block per (row, class)
  const uint32_t need = 2 * acc_cols;
  const uint32_t tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
  const int pitch = g.TW + g.box_wr;  // pixel rows per input image row inside the halo box
  int t0, tcnt;
  tile_range(g.num_tiles, t0, tcnt);

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    if (g.st_rows) tma_prefetch_desc(&tmY);
    mbar_init(w_full, 1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, tmem_cols);
  if (threadIdx.x < 128) sbias[threadIdx.x] = (g.bias && (int)threadIdx.x < g.Cout) ? g.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);

  if (warp == W_TMA) {
    // ---- TMA producer: the whole warp walks the tiles (uniform state), one elected lane issues ----
    if (elect_one()) {
      mbar_expect_tx(w_full, 9u * (uint32_t)g.Cout * RB);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * wsub, &tmW, w_full, t * g.Cin, 0);
    }
    int s = 0;
    uint32_t ph = 1;  // parity to wait for on empty_bar[s]
    for (int i = 0; i < tcnt; ++i) {
      const int tile = t0 + i;
      const int tw = tile % g.tiles_w;
      const int th = (tile / g.tiles_w) % g.tiles_h;
      const int n = tile / (g.tiles_w * g.tiles_h);
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        if (g.debug & 2) {
          mbar_arrive(&full_bar[s]);
        } else {
          mbar_expect_tx(&full_bar[s], (uint32_t)(TH + g.box_hr) * (uint32_t)pitch * RB);
          tma_load_4d(sH + s * hsub, &tmX, &full_bar[s], 0, tw * g.TW + g.box_w0, th * TH + g.box_h0, n);
        }
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == W_MMA) {
    // ---- MMA issuer: warp-uniform loop, tcgen05.mma / commit under elect_one() ----
    const uint32_t idesc = make_idesc_bf16(128, g.Cout, 0, 0);
    const uint64_t d_hi = make_smem_desc(0, 16, 8 * RB, RB);
    // per-tap start offsets inside the halo (A) and the weight bank (B), in 16-byte units
    uint32_t ta[9], tb[9], tc_[9], tf[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int tt = t < g.ntaps ? t : 0;
      ta[t] = (uint32_t)(((g.dh[tt] - g.box_h0) * pitch + (g.dw[tt] - g.box_w0)) * RB) >> 4;
      tb[t] = (uint32_t)(g.widx[tt] * wsub) >> 4;
      tc_[t] = (uint32_t)(g.tap_cls[tt] * g.Cout);  // TMEM column offset of the tap's class inside a row's accumulators
      tf[t] = (uint32_t)g.tap_first[tt];
    }
    const uint32_t row_step = (uint32_t)(pitch * RB) >> 4;
    const uint32_t cout = (uint32_t)(g.ncls * g.Cout);
    const int ntaps = g.ntaps;
    const bool no_mma = (g.debug & 1) != 0;
    mbar_wait(w_full, 0);
    const uint32_t w_lo = smem_u32(sW) >> 4;
    const uint32_t h_base = smem_u32(sH) >> 4, h_step = (uint32_t)hsub >> 4;
    int s = 0;
    uint32_t ph = 0;  // parity to wait for on full_bar[s]
    for (int i = 0; i < tcnt; ++i) {
      const uint32_t a = i & 1, aph = (i >> 1) & 1;
      mbar_wait(&tempty_bar[a], aph ^ 1);
      mbar_wait(&full_bar[s], ph);
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
                  if (j > 0)
                    umma_bf16_acc(d_tmem + tc_[t], d_hi | (uint64_t)(hr + ta[t] + 2 * j), d_hi | (uint64_t)(w_lo + tb[t] + 2 * j), idesc);
                  else
                    umma_bf16(d_tmem + tc_[t], d_hi | (uint64_t)(hr + ta[t]), d_hi | (uint64_t)(w_lo + tb[t]), idesc, tf[t] ^ 1u);
                }
              }
            }
          }
        }
        umma_commit(&empty_bar[s]);
        umma_commit(&tfull_bar[a]);
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    // ---- epilogue: thread = pixel column px of every output row of the tile ----
    const int q = warp & 3, hf = warp >> 2;
    const int px = q * 32 + lane;
    float a0[32], a1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      a0[j] = 0.f;
      a1[j] = 0.f;
    }
    int cur_n = -1;
    for (int i = 0; i < tcnt; ++i) {
      const uint32_t a = i & 1, aph = (i >> 1) & 1;
      const int tile = t0 + i;
      const int tw = tile % g.tiles_w;
      const int th = (tile / g.tiles_w) % g.tiles_h;
      const int n = tile / (g.tiles_w * g.tiles_h);
      if (MODE != 0 && MODE != 2 && n != cur_n) {
        if (cur_n >= 0) epi_flush<MODE>(g, cur_n, hf, g.Cout >> 4, lane, a0, a1);
        cur_n = n;
      }
      const int gw = tw * g.TW + px;
      const bool okw = px < g.TW && gw < g.GW;
      const long long img = (long long)n * g.OH;
      const int gh0 = th * TH;
      uint8_t* sbuf = g.st_rows ? sOut + (i % g.st_bufs) * g.st_bytes : nullptr;
      DBG_STAMP(i, 0);
      if (g.st_rows) stage_acquire(g);
      DBG_STAMP(i, 1);
      mbar_wait(&tfull_bar[a], aph);
      DBG_STAMP(i, 2);
      tc_fence_after();
      // accumulator block rr = (tile row, class): grid pixel (gh0 + row, gw) -> output pixel (osh gh + ph, osw gw + pw);
      // staging tile = the dense block of (TH osh) x (TW osw) output pixels of this tile
      epi_tile<MODE>(g, tmem_base + a * acc_cols, TH * g.ncls, q, hf, sbias,
                     [&](int rr, bool& ok, int& spix) {
                       const int row = g.ncls == 1 ? rr : rr / g.ncls, cl = g.ncls == 1 ? 0 : rr - row * g.ncls;
                       const int gh = gh0 + row;
                       const int oh = gh * g.osh + g.cls_ph[cl], ow = gw * g.osw + g.cls_pw[cl];
                       ok = okw && gh < g.GH && oh < g.OH && ow < g.OW;
                       spix = (row * g.osh + g.cls_ph[cl]) * g.st_cols + px * g.osw + g.cls_pw[cl];
                       return ((img + oh) * g.OW + ow) * g.Cout;
                     },
                     a0, a1, sbuf);
      DBG_STAMP(i, 3);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
      DBG_STAMP(i, 4);
      if (g.st_rows) stage_store(g, &tmY, sbuf, tw * g.TW * g.osw, gh0 * g.osh, n);
      DBG_STAMP(i, 5);
    }
    if (MODE != 0 && (MODE == 2 || cur_n >= 0)) epi_flush<MODE>(g, cur_n, hf, g.Cout >> 4, lane, a0, a1);
    if (g.st_rows && threadIdx.x == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
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

int mode_of(const ConvTcArgs& a) { return a.in_sums ? 1 : (a.colsum ? 2 : (a.in_bsums ? 3 : 0)); }

template <int RB, int G, int MODE>
int launch_generic(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmY, const ConvTcArgs& a, cudaStream_t st) {
  auto kern = conv_tc_kernel<RB, G, MODE>;
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<RB, G>::SMEM));
    configured = true;
  }
  int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  OmrLaunch(grid, NTHREADS, Cfg<RB, G>::SMEM, st)(kern, tmX, tmW, tmY, a);
  OMR_LAUNCHED();
  return OMR_OK;
}
template <int RB, int G>
int launch_generic_mode(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmY, const ConvTcArgs& a, cudaStream_t st) {
  switch (mode_of(a)) {
    case 0: return launch_generic<RB, G, 0>(tmX, tmW, tmY, a, st);
    case 1: return launch_generic<RB, G, 1>(tmX, tmW, tmY, a, st);
    case 2: return launch_generic<RB, G, 2>(tmX, tmW, tmY, a, st);
    default: return launch_generic<RB, G, 3>(tmX, tmW, tmY, a, st);
  }
}

template <int RB, int MODE>
int launch_halo(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmY, const ConvTcArgs& h, int wsub, int hsub,
                int stages, int smem_bytes, cudaStream_t st) {
  auto kern = conv_halo_kernel<RB, MODE>;
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = h.num_tiles < num_sms() ? h.num_tiles : num_sms();
  if (h.debug & 8) {  // diagnostics only (synchronises): per-tile stamps of the epilogue
    static long long* dbg_dev = nullptr;
    static int dumps = 0;
    if (!dbg_dev) cudaMalloc(&dbg_dev, 64 * 8 * sizeof(long long));
    cudaMemsetAsync(dbg_dev, 0, 64 * 8 * sizeof(long long), st);
    ConvTcArgs hd = h;
    hd.dbg = dbg_dev;
    OmrLaunch(grid, NTHREADS, smem_bytes, st)(kern, tmX, tmW, tmY, hd, wsub, hsub, stages);
    OMR_LAUNCHED();
    cudaStreamSynchronize(st);
    static long long hb[64 * 8];
    cudaMemcpy(hb, dbg_dev, sizeof(hb), cudaMemcpyDeviceToHost);
    if (dumps++ < 2) {
      fprintf(stderr, "[conv_halo dbg] Cin %d Cout %d TH %d ncls %d tiles %d grid %d stages %d st_bufs %d mode %d\n", h.Cin, h.Cout, h.TH,
              h.ncls, h.num_tiles, grid, stages, h.st_bufs, MODE);
      for (int it = 8; it < 20; ++it)
        fprintf(stderr, "  it%2d period %6lld | acquire %5lld  wait_tfull %5lld  units %5lld  arrive %5lld  store %5lld\n", it,
                hb[it * 8] - hb[(it - 1) * 8], hb[it * 8 + 1] - hb[it * 8], hb[it * 8 + 2] - hb[it * 8 + 1], hb[it * 8 + 3] - hb[it * 8 + 2],
                hb[it * 8 + 4] - hb[it * 8 + 3], hb[it * 8 + 5] - hb[it * 8 + 4]);
    }
    return OMR_OK;
  }
  OmrLaunch(grid, NTHREADS, smem_bytes, st)(kern, tmX, tmW, tmY, h, wsub, hsub, stages);
  OMR_LAUNCHED();
  return OMR_OK;
}
template <int RB>
int launch_halo_mode(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmY, const ConvTcArgs& h, int wsub, int hsub,
                     int stages, int smem_bytes, cudaStream_t st) {
  switch (mode_of(h)) {
    case 0: return launch_halo<RB, 0>(tmX, tmW, tmY, h, wsub, hsub, stages, smem_bytes, st);
    case 1: return launch_halo<RB, 1>(tmX, tmW, tmY, h, wsub, hsub, stages, smem_bytes, st);
    case 2: return launch_halo<RB, 2>(tmX, tmW, tmY, h, wsub, hsub, stages, smem_bytes, st);
    default: return launch_halo<RB, 3>(tmX, tmW, tmY, h, wsub, hsub, stages, smem_bytes, st);
  }
}

// tensor map of the output y [N, OH, OW, Cout] for staged TMA stores: box = st_cols x st_rows pixels of min(Cout, 64) channels
int make_out_map(CUtensorMap* tm, const ConvTcArgs& a) {
  const int cb = a.Cout >= 64 ? 64 : a.Cout;
  unsigned long long dims[4] = {(unsigned long long)a.Cout, (unsigned long long)a.OW, (unsigned long long)a.OH, (unsigned long long)a.N};
  unsigned long long strides[3] = {(unsigned long long)a.Cout * 2, (unsigned long long)a.OW * a.Cout * 2,
                                   (unsigned long long)a.OH * a.OW * a.Cout * 2};
  unsigned int box[4] = {(unsigned)cb, (unsigned)a.st_cols, (unsigned)a.st_rows, 1u};
  return omr_make_tensor_map(tm, 2, a.y, 4, dims, strides, box, nullptr, cb * 2);
}

int g_stage_mode = -1;
int stage_bufs_wanted() {  // OMR_CONV_STAGE: 0 = direct global stores, 1 = one staging buffer, 2 (default) = two when they fit
  if (g_stage_mode < 0) {
    const char* e = getenv("OMR_CONV_STAGE");
    g_stage_mode = e ? atoi(e) : 0;  // measured: staged TMA stores are not faster than direct stores (round 2)
    if (g_stage_mode < 0 || g_stage_mode > 2) g_stage_mode = 2;
  }
  return g_stage_mode;
}

// One launch of the tap-GEMM.  x: [N, XH, XW, Cin] bf16; wpack: [Cout, 9*Cin] bf16 (tap-major, channels innermost).
int run_taps(const void* x, int N, int XH, int XW, int Cin, const void* wpack, int Cout, ConvTcArgs a, cudaStream_t st) {
  const int rb = (Cin >= 64 ? 64 : Cin) * 2;
  a.debug = conv_debug();
  if (mode_of(a) != 0 && Cout > 64) return OMR_TC_NOT_ELIGIBLE;  // the caller must not ask (fused sums need Cout <= 64)
  const bool upsample = a.ncls > 1;
  if (!upsample) {  // plain stride-1 convolution / one parity class: a single output class, 3 x 3 halo
    a.ncls = 1;
    a.box_h0 = -1; a.box_w0 = -1; a.box_hr = 2; a.box_wr = 2;
    a.cls_ph[0] = a.oph; a.cls_pw[0] = a.opw;
    for (int t = 0; t < MAX_TAPS; ++t) { a.tap_cls[t] = 0; a.tap_first[t] = t == 0; }
  }
  if (halo_enabled() && a.ish == 1 && a.isw == 1 && Cin <= 64 && (upsample || (a.ntaps == 9 && a.osh == 1 && a.osw == 1))) {
    const int TW = a.GW >= 128 ? 128 : a.GW;
    const int pitch = TW + a.box_wr;
    const int wsub = (Cout * rb + 1023) / 1024 * 1024;
    // TH output rows per tile: fewer TMA rows per output pixel ((TH+2)/TH instead of 3) and fewer barrier round trips;
    // bounded by TMEM (2 buffers x TH x classes x Cout columns <= 512) and by shared memory (>= 2 halo stages next to the weights)
    // Staged output (TMA store) needs a staging tile of (TH osh) x (TW osw) pixels; preference: two staging buffers, then one,
    // then direct stores, each with the largest TH that leaves >= 2 halo stages.
    int TH = 4, hsub = 0, stages = 0, st_bufs = 0, st_bytes = 0;
    bool found = false;
    for (int want = stage_bufs_wanted(); want >= 0 && !found; --want) {
      for (TH = 4; TH >= 1; TH >>= 1) {
        if (2 * TH * a.ncls * Cout > 512 || (TH > 1 && a.GH < TH)) continue;
        if (TH * a.osh > 256 || TW * a.osw > 256) continue;
        int rows = (TH + a.box_hr) * pitch;
        // a tap's 128-row operand window starts up to box_hr * pitch + box_wr rows into the last tile row's box row
        if (rows < (TH - 1 + a.box_hr) * pitch + a.box_wr + 128) rows = (TH - 1 + a.box_hr) * pitch + a.box_wr + 128;
        hsub = (rows * rb + 1023) / 1024 * 1024;
        st_bufs = want;
        st_bytes = want ? ((TH * a.osh) * (TW * a.osw) * Cout * 2 + 1023) / 1024 * 1024 : 0;
        stages = (225 * 1024 - 1024 - 1024 - 9 * wsub - st_bufs * st_bytes) / hsub;
        if (stages > 4) stages = 4;
        if (stages >= 2) {
          found = true;
          break;
        }
      }
    }
    if (found) {
      ConvTcArgs h = a;
      h.TH = TH; h.TW = TW;
      h.st_bufs = st_bufs; h.st_bytes = st_bytes;
      h.st_rows = st_bufs ? TH * a.osh : 0; h.st_cols = st_bufs ? TW * a.osw : 0;
      h.tiles_w = (a.GW + TW - 1) / TW;
      h.tiles_h = (a.GH + TH - 1) / TH;
      h.num_tiles = a.N * h.tiles_h * h.tiles_w;
      h.Cin = Cin; h.Cout = Cout;
      CUtensorMap tmX, tmW;
      unsigned long long dims[4] = {(unsigned long long)Cin, (unsigned long long)XW, (unsigned long long)XH, (unsigned long long)N};
      unsigned long long strides[3] = {(unsigned long long)Cin * 2, (unsigned long long)XW * Cin * 2, (unsigned long long)XH * XW * Cin * 2};
      unsigned int box[4] = {(unsigned)Cin, (unsigned)pitch, (unsigned)(TH + a.box_hr), 1u};
      int rc = omr_make_tensor_map(&tmX, 2, x, 4, dims, strides, box, nullptr, rb);
      if (rc) return rc;
      unsigned long long wd[2] = {(unsigned long long)9 * Cin, (unsigned long long)Cout};
      unsigned long long ws[1] = {(unsigned long long)9 * Cin * 2};
      unsigned int wb[2] = {(unsigned)Cin, (unsigned)Cout};
      rc = omr_make_tensor_map(&tmW, 2, wpack, 2, wd, ws, wb, nullptr, rb);
      if (rc) return rc;
      CUtensorMap tmY = tmX;
      if (h.st_rows) {
        rc = make_out_map(&tmY, h);
        if (rc) return rc;
      }
      const int smem_bytes = 9 * wsub + stages * hsub + st_bufs * st_bytes + 1024 + 1024;
      if (rb == 32) return launch_halo_mode<32>(tmX, tmW, tmY, h, wsub, hsub, stages, smem_bytes, st);
      if (rb == 64) return launch_halo_mode<64>(tmX, tmW, tmY, h, wsub, hsub, stages, smem_bytes, st);
      return launch_halo_mode<128>(tmX, tmW, tmY, h, wsub, hsub, stages, smem_bytes, st);
    }
  }
  if (upsample) return OMR_TC_NOT_ELIGIBLE;  // the caller falls back to one launch per parity class
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
  CUtensorMap tmY = tmX;
  if (stage_bufs_wanted() > 0 && a.osh == 1 && a.osw == 1) {  // dense output patch: staged TMA store
    a.st_rows = TH; a.st_cols = TW; a.st_bufs = stage_bufs_wanted();
    a.st_bytes = 128 * Cout * 2;  // <= 32 KB; Cfg reserves two of them
    int rc = make_out_map(&tmY, a);
    if (rc) return rc;
  }
  if (rb == 32) return launch_generic_mode<32, 3>(tmX, tmW, tmY, a, st);
  if (rb == 64) return launch_generic_mode<64, 3>(tmX, tmW, tmY, a, st);
  return launch_generic_mode<128, 1>(tmX, tmW, tmY, a, st);
}

bool shape_ok(int Ci, int Co) {
  auto okc = [](int c) { return c == 16 || c == 32 || c == 64 || c == 128; };
  return okc(Ci) && okc(Co);
}

}  // namespace

// in_sums (nullable, Co <= 64): [N][Co][2] += (sum y, sum y^2) of the stored output (InstanceNorm statistics)
int omr_conv3x3_fwd_tc(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Ci, int Co,
                       int sh, int sw, int relu, double* in_sums, cudaStream_t st) {
  if (!shape_ok(Ci, Co) || N < 1) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return OMR_TC_NOT_ELIGIBLE;
  if (in_sums && Co > 64) return OMR_TC_NOT_ELIGIBLE;
  ConvTcArgs a{};
  a.y = (bf16*)y; a.bias = bias; a.relu = relu; a.N = N;
  a.in_sums = in_sums;
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
// colsum (nullable, Ci <= 64): [Ci] += column sums of the stored dx; in_x / in_bsums (nullable, Ci <= 64): InstanceNorm
// backward sums [N][Ci][2] += (sum dx, sum dx * in_x) with in_x laid out like dx.
int omr_conv3x3_dgrad_tc(const void* dy, const void* wT, void* dx, int N, int H, int W, int Ci, int Co, int sh, int sw,
                         const void* mask, float mask_scale, float* colsum, const void* in_x, double* in_bsums, cudaStream_t st) {
  if (!shape_ok(Ci, Co) || N < 1 || sh > 2 || sw > 2) return OMR_TC_NOT_ELIGIBLE;
  if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(wT) & 15) || (reinterpret_cast<uintptr_t>(dx) & 15) ||
      (reinterpret_cast<uintptr_t>(mask) & 15) || (reinterpret_cast<uintptr_t>(in_x) & 15))
    return OMR_TC_NOT_ELIGIBLE;
  if ((colsum || in_bsums) && Ci > 64) return OMR_TC_NOT_ELIGIBLE;
  if (colsum && in_bsums) return OMR_TC_NOT_ELIGIBLE;
  const int Ho = (H + sh - 1) / sh, Wo = (W + sw - 1) / sw;
  if ((sh > 1 || sw > 1) && Ci <= 64 && Co <= 64) {
    // strided data gradient as ONE launch over the dy grid (fractionally-strided convolution): grid pixel (i, j) produces its
    // sh x sw output pixels, class (ph, pw) from the taps kh = ph + 1 (mod sh), kw = pw + 1 (mod sw) -- nine MMAs per
    // grid tile like a stride-1 convolution, dy fetched once instead of once per parity class
    ConvTcArgs a{};
    a.y = (bf16*)dx; a.N = N;
    a.mask = (const bf16*)mask; a.mask_scale = mask_scale;
    a.colsum = colsum; a.in_x = (const bf16*)in_x; a.in_bsums = in_bsums;
    a.GH = Ho; a.GW = Wo; a.ish = 1; a.isw = 1;
    a.OH = H; a.OW = W; a.osh = sh; a.osw = sw; a.oph = 0; a.opw = 0;
    a.ncls = sh * sw;
    int dhmin = 9, dhmax = -9, dwmin = 9, dwmax = -9;
    a.ntaps = 0;
    for (int ph = 0; ph < sh; ++ph)
      for (int pw = 0; pw < sw; ++pw) {
        const int cl = ph * sw + pw;
        a.cls_ph[cl] = ph; a.cls_pw[cl] = pw;
        bool first = true;
        for (int kh = 0; kh < 3; ++kh) {
          if ((ph + 1 - kh) % sh != 0) continue;
          for (int kw = 0; kw < 3; ++kw) {
            if ((pw + 1 - kw) % sw != 0) continue;
            const int t = a.ntaps++;
            a.dh[t] = (ph + 1 - kh) / sh; a.dw[t] = (pw + 1 - kw) / sw; a.widx[t] = kh * 3 + kw;
            a.tap_cls[t] = cl; a.tap_first[t] = first ? 1 : 0;
            first = false;
            dhmin = a.dh[t] < dhmin ? a.dh[t] : dhmin; dhmax = a.dh[t] > dhmax ? a.dh[t] : dhmax;
            dwmin = a.dw[t] < dwmin ? a.dw[t] : dwmin; dwmax = a.dw[t] > dwmax ? a.dw[t] : dwmax;
          }
        }
      }
    a.box_h0 = dhmin; a.box_w0 = dwmin; a.box_hr = dhmax - dhmin; a.box_wr = dwmax - dwmin;
    int rc = run_taps(dy, N, Ho, Wo, Co, wT, Ci, a, st);
    if (rc != OMR_TC_NOT_ELIGIBLE) return rc;
  }
  for (int ph = 0; ph < sh; ++ph)
    for (int pw = 0; pw < sw; ++pw) {
      ConvTcArgs a{};
      a.y = (bf16*)dx; a.bias = nullptr; a.relu = 0; a.N = N;
      a.mask = (const bf16*)mask; a.mask_scale = mask_scale;
      a.colsum = colsum; a.in_x = (const bf16*)in_x; a.in_bsums = in_bsums;
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
