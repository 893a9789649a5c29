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
// * the same kernels compute the data gradient: stride 1 -> taps mirrored, weights from the [Ci,3,3,Co] pack; stride 2 ->
//   ONE launch over the dy grid (a fractionally-strided convolution: grid pixel (i, j) produces its sh x sw output pixels,
//   one accumulator per output parity class) when the channel counts allow, else one launch per parity class.
//
// Warp roles (320 threads): 0-7 epilogue, 8 TMA producer, 9 MMA issuer + TMEM allocator.  Producer and issuer run their
// loops warp-wide on uniform values with the instruction under elect_one() (tc_common.cuh).
//
// The epilogue is what bounds the narrow (C = 16 / 32) high-resolution layers (ncu, round 2: 85 % of the executed
// instructions of the first version were address arithmetic, parameter re-loads and branches on run-time flags; ~550 clk per
// 16-channel unit and warp).  It is therefore specialised at compile time -- NCHUNK = Cout / 16, EPI = what is applied
// (bias + ReLU | fused ReLU/dropout-backward mask | nothing), MODE = which per-channel sums are accumulated -- keeps the
// bias in registers, walks tile coordinates and offsets incrementally (no divisions, 32-bit deltas), keeps two TMEM loads in
// flight and works on packed bf16 pairs.  Two warps per SM sub-partition: warps w and w + 4 share the TMEM lanes
// 32 (w & 3) .. + 31 and split a tile's (row, 16-channel chunk) units.
// (A staged TMA-store epilogue -- swizzled shared-memory tile + cp.async.bulk.tensor store -- was built and measured in
// round 2: no gain over the direct 2 x 16-byte stores per lane, so it was removed again.)
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
constexpr int EPI_FWD = 0, EPI_MASK = 1, EPI_PLAIN = 2;

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
  const bf16* mask;                // EPI_MASK: out = mask > 0 ? out * mask_scale : 0 (same layout as y)
  float mask_scale;
  // fused reductions (see the header): exactly one of them is non-null in MODE 1 / 2 / 3
  double* in_sums;     // MODE 1: [N][Cout][2] += (sum y, sum y^2)
  float* colsum;       // MODE 2: [Cout] += sum y
  const bf16* in_x;    // MODE 3: the InstanceNorm input (same layout as y) ...
  double* in_bsums;    //         ... and [N][Cout][2] += (sum y, sum y * x)
  // generic kernel: pipeline geometry (run-time, so that the kernel is instantiated per epilogue only)
  int rb, G, a_sub, b_sub, stage_bytes, stages;
  int debug;           // OMR_CONV_DEBUG (diagnostics, results are WRONG when set): 1 = no MMAs, 2 = no TMA loads, 4 = no global stores
};

// ---- epilogue building blocks ---------------------------------------------------------------------------------------
// contiguous, balanced tile range of this CTA
__device__ __forceinline__ void tile_range(int num_tiles, int& t0, int& cnt) {
  const int base = num_tiles / (int)gridDim.x, rem = num_tiles % (int)gridDim.x;
  const int b = (int)blockIdx.x;
  t0 = b * base + (b < rem ? b : rem);
  cnt = base + (b < rem ? 1 : 0);
}

__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// Per-thread epilogue state that does not change over the kernel: the bias of this warp's chunks (registers), the ReLU floor,
// the mask scale, and the per-thread partial sums of the fused reductions.
template <int NCHUNK, int MODE>
struct EpiRegs {
  static constexpr int K = NCHUNK == 1 ? 1 : NCHUNK / 2;  // chunks per warp half
  static constexpr int KA = MODE == 0 ? 1 : (K > 2 ? 2 : K);
  float bias[K][16];
  float a0[KA][16], a1[KA][16];
};

// One (pixel, 16-channel chunk) unit: fp32 accumulators -> (+bias, ReLU | mask | nothing) -> bf16 -> global, plus the sums.
template <int EPI, int MODE>
__device__ __forceinline__ void epi_unit(const uint32_t (&v)[16], const float (&b)[16], __nv_bfloat162 floor2, float mscale,
                                         const bf16* __restrict__ maskp, const bf16* __restrict__ xinp, bf16* __restrict__ dst,
                                         bool store, float (&a0)[16], float (&a1)[16]) {
  __nv_bfloat162 h[8];
  if (EPI == EPI_FWD) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      h[j] = __hmax2(__floats2bfloat162_rn(__uint_as_float(v[2 * j]) + b[2 * j], __uint_as_float(v[2 * j + 1]) + b[2 * j + 1]), floor2);
  } else if (EPI == EPI_MASK) {
    const uint4 m0 = ldg16(maskp), m1 = ldg16(maskp + 8);
    const __nv_bfloat162* mb0 = reinterpret_cast<const __nv_bfloat162*>(&m0);
    const __nv_bfloat162* mb1 = reinterpret_cast<const __nv_bfloat162*>(&m1);
    const __nv_bfloat162 zero = __float2bfloat162_rn(0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 m = j < 4 ? mb0[j] : mb1[j - 4];
      h[j] = __hmul2(__floats2bfloat162_rn(__uint_as_float(v[2 * j]) * mscale, __uint_as_float(v[2 * j + 1]) * mscale), __hgt2(m, zero));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
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
    const uint4 x0 = ldg16(xinp), x1 = ldg16(xinp + 8);
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
  if (store) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = *reinterpret_cast<const uint4*>(&h[0]);
    d4[1] = *reinterpret_cast<const uint4*>(&h[4]);
  }
}

// add this warp's per-thread partial sums of sample n to the global accumulators and clear them
template <int NCHUNK, int MODE>
__device__ __forceinline__ void epi_flush(const ConvTcArgs& g, int n, int hf, int lane, EpiRegs<NCHUNK, MODE>& e) {
  if (MODE == 0) return;
  constexpr int Cout = NCHUNK * 16;
#pragma unroll
  for (int k = 0; k < EpiRegs<NCHUNK, MODE>::KA; ++k) {
    const int chunk = NCHUNK == 1 ? 0 : hf + 2 * k;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float s0 = warp_sum(e.a0[k][j]);
      const float s1 = (MODE == 2) ? 0.f : warp_sum(e.a1[k][j]);
      if (lane == j) {
        const int c = chunk * 16 + j;
        if (MODE == 1) {
          atomicAdd(g.in_sums + ((long long)n * Cout + c) * 2, (double)s0);
          atomicAdd(g.in_sums + ((long long)n * Cout + c) * 2 + 1, (double)s1);
        } else if (MODE == 2) {
          atomicAdd(g.colsum + c, s0);
        } else {
          atomicAdd(g.in_bsums + ((long long)n * Cout + c) * 2, (double)s0);
          atomicAdd(g.in_bsums + ((long long)n * Cout + c) * 2 + 1, (double)s1);
        }
      }
      e.a0[k][j] = 0.f;
      e.a1[k][j] = 0.f;
    }
  }
}

template <int NCHUNK, int MODE>
__device__ __forceinline__ void epi_init(const ConvTcArgs& g, int hf, EpiRegs<NCHUNK, MODE>& e) {
  using E = EpiRegs<NCHUNK, MODE>;
#pragma unroll
  for (int k = 0; k < E::K; ++k) {
    const int chunk = NCHUNK == 1 ? 0 : hf + 2 * k;
#pragma unroll
    for (int j = 0; j < 16; ++j) e.bias[k][j] = g.bias ? __ldg(g.bias + chunk * 16 + j) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < E::KA; ++k)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      e.a0[k][j] = 0.f;
      e.a1[k][j] = 0.f;
    }
}

// The epilogue of one tile: `rows` accumulator blocks of 128 pixels x Cout channels in TMEM at t_base (+ rr * Cout).
// Warp (q = warp & 3, hf = warp >> 2), lane -> TMEM lane 32 q + lane of every block; the (block, 16-channel chunk) units are
// dealt to the two halves: NCHUNK = 1 -> blocks alternate, NCHUNK >= 2 -> chunk parity = hf (a thread's accumulators and bias
// registers belong to fixed channels).  unit(rr, ok, off): validity and element offset (32-bit, relative to base) of this
// thread's pixel of block rr.  Two TMEM loads are in flight per wait.
template <int NCHUNK, int EPI, int MODE, typename OffFn>
__device__ __forceinline__ void epi_tile(const ConvTcArgs& g, uint32_t t_base, int rows, int q, int hf, long long base, bool st_ok,
                                         __nv_bfloat162 floor2, EpiRegs<NCHUNK, MODE>& e, OffFn unit) {
  constexpr int Cout = NCHUNK * 16;
  const uint32_t t_lane = t_base + ((uint32_t)(q * 32) << 16);
  bf16* const yb = g.y + base;
  const bf16* const mb = EPI == EPI_MASK ? g.mask + base : nullptr;
  const bf16* const xb = MODE == 3 ? g.in_x + base : nullptr;
  const float msc = g.mask_scale;
  if (NCHUNK == 1) {
    for (int rr = hf; rr < rows; rr += 4) {  // blocks rr and rr + 2 in flight
      uint32_t v0[16], v1[16];
      const bool two = rr + 2 < rows;
      tmem_ld16(t_lane + (uint32_t)(rr * 16), v0);
      if (two) tmem_ld16(t_lane + (uint32_t)((rr + 2) * 16), v1);
      bool ok0, ok1 = false;
      int o0, o1 = 0;
      unit(rr, ok0, o0);
      if (two) unit(rr + 2, ok1, o1);
      tmem_ld_wait();
      if (ok0) epi_unit<EPI, MODE>(v0, e.bias[0], floor2, msc, mb + o0, xb + o0, yb + o0, st_ok, e.a0[0], e.a1[0]);
      if (ok1) epi_unit<EPI, MODE>(v1, e.bias[0], floor2, msc, mb + o1, xb + o1, yb + o1, st_ok, e.a0[0], e.a1[0]);
    }
  } else {
    constexpr int K = NCHUNK / 2;
    for (int rr = 0; rr < rows; ++rr) {
      bool ok;
      int o;
      unit(rr, ok, o);
      const uint32_t t_row = t_lane + (uint32_t)(rr * Cout) + (uint32_t)(hf * 16);
#pragma unroll
      for (int k = 0; k < K; k += 2) {  // chunks hf + 2k and hf + 2k + 2 in flight
        uint32_t v0[16], v1[16];
        tmem_ld16(t_row + (uint32_t)(k * 32), v0);
        if (k + 1 < K) tmem_ld16(t_row + (uint32_t)(k * 32 + 32), v1);
        tmem_ld_wait();
        if (ok) {
          const int c0 = o + (hf + 2 * k) * 16;
          constexpr int KA = EpiRegs<NCHUNK, MODE>::KA;
          epi_unit<EPI, MODE>(v0, e.bias[k], floor2, msc, mb + c0, xb + c0, yb + c0, st_ok, e.a0[k < KA ? k : 0], e.a1[k < KA ? k : 0]);
          if (k + 1 < K)
            epi_unit<EPI, MODE>(v1, e.bias[k + 1 < K ? k + 1 : 0], floor2, msc, mb + c0 + 32, xb + c0 + 32, yb + c0 + 32, st_ok,
                                e.a0[k + 1 < KA ? k + 1 : 0], e.a1[k + 1 < KA ? k + 1 : 0]);
        }
      }
    }
  }
}

// MMA issuer of the halo kernel: warp-uniform loop, tcgen05.mma / commit under elect_one().  NJ = k-steps (of 16 channels) per tap.
template <int NJ, int Cout>
__device__ __forceinline__ void halo_mma_role(const ConvTcArgs& g, uint8_t* sW, uint8_t* sH, int wsub, int hsub, int stages, int pitch,
                                              int TH, uint32_t acc_cols, uint32_t tmem_base, int t0, int tcnt, uint64_t* w_full,
                                              uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tfull_bar, uint64_t* tempty_bar) {
  (void)t0;
  constexpr int RB = NJ * 32;
  const uint32_t idesc = make_idesc_bf16(128, Cout, 0, 0);
  const uint64_t d_hi = make_smem_desc(0, 16, 8 * RB, RB);
  // per-tap start offsets inside the halo (A) and the weight bank (B), in 16-byte units
  uint32_t ta[9], tb[9], tc_[9], tf[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int tt = t < g.ntaps ? t : 0;
    ta[t] = (uint32_t)(((g.dh[tt] - g.box_h0) * pitch + (g.dw[tt] - g.box_w0)) * RB) >> 4;
    tb[t] = (uint32_t)(g.widx[tt] * wsub) >> 4;
    tc_[t] = (uint32_t)(g.tap_cls[tt] * Cout);  // TMEM column offset of the tap's class inside a row's accumulators
    tf[t] = (uint32_t)g.tap_first[tt];
  }
  const uint32_t row_step = (uint32_t)(pitch * RB) >> 4;
  const uint32_t cout = (uint32_t)(g.ncls * Cout);
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
              for (int j = 0; j < NJ; ++j) {
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
}

// MMA issuer of the generic kernel (NJ k-steps per sub-tile, G sub-tiles per stage: the host picks G = 3 for 32/64-byte rows, 1 for 128)
template <int NJ, int G, int Cout>
__device__ __forceinline__ void generic_mma_role(const ConvTcArgs& g, uint8_t* smem, int nsub, int ngroups, uint32_t tmem_base, int tcnt,
                                                 uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tfull_bar, uint64_t* tempty_bar) {
  constexpr int RB = NJ * 32;
  const uint32_t idesc = make_idesc_bf16(128, Cout, 0, 0);
  const uint64_t d_hi = make_smem_desc(0, 16, 8 * RB, RB);
  const uint32_t smem_lo = smem_u32(smem) >> 4;
  const uint32_t a_sub16 = (uint32_t)g.a_sub >> 4, b_sub16 = (uint32_t)g.b_sub >> 4, st16 = (uint32_t)g.stage_bytes >> 4;
  const int STAGES = g.stages;
  int s = 0;
  uint32_t ph = 0;
  for (int i = 0; i < tcnt; ++i) {
    const uint32_t a = i & 1, aph = (i >> 1) & 1;
    mbar_wait(&tempty_bar[a], aph ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + a * (uint32_t)Cout;
    for (int grp = 0; grp < ngroups; ++grp) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const int sub0 = grp * G;
      const int cnt = (nsub - sub0) < G ? (nsub - sub0) : G;
      const uint32_t stage = smem_lo + (uint32_t)s * st16;
      if (elect_one()) {
#pragma unroll
        for (int q = 0; q < G; ++q) {
          if (q < cnt) {
            const uint32_t a_addr = stage + q * a_sub16;
            const uint32_t b_addr = stage + G * a_sub16 + q * b_sub16;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              if (q == 0 && j == 0)
                umma_bf16(d_tmem, d_hi | (uint64_t)(a_addr), d_hi | (uint64_t)(b_addr), idesc, grp == 0 ? 0u : 1u);
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
}

// ---------------------------------------------------------------------------------------------------------------
// Generic tap-GEMM kernel: one 128-pixel TH x TW patch per tile, every (tap, 64-channel chunk) sub-tile fetched by its own
// TMA box (strided convolutions, C_in = 128, parity-class data gradients).  Pipeline geometry (rb = row bytes, G sub-tiles
// per stage, stage count) is passed at run time so that the kernel is instantiated per epilogue only.
// ---------------------------------------------------------------------------------------------------------------
template <int NCHUNK, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                              const __grid_constant__ CUtensorMap tmW, ConvTcArgs g) {
  omr_pdl_enter();
  constexpr int Cout = NCHUNK * 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = g.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * g.stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int RB = g.rb, G = g.G;
  const int chunks = (g.Cin * 2 + RB - 1) / RB;  // 64-channel chunks per tap (1 or 2)
  const int nsub = g.ntaps * chunks;
  const int ngroups = (nsub + G - 1) / G;
  constexpr uint32_t tmem_cols = Cout * 2 <= 32 ? 32 : (Cout * 2 <= 64 ? 64 : (Cout * 2 <= 128 ? 128 : 256));
  int t0, tcnt;
  tile_range(g.num_tiles, t0, tcnt);

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t a_box_bytes = (uint32_t)(g.TH * g.TW) * RB;
  const uint32_t b_box_bytes = (uint32_t)Cout * RB;

  if (warp == W_TMA) {
    // ---- TMA producer: warp-uniform loop, one elected lane issues ----
    int s = 0;
    uint32_t ph = 1;
    int tw = t0 % g.tiles_w, th = (t0 / g.tiles_w) % g.tiles_h, n = t0 / (g.tiles_w * g.tiles_h);
    for (int i = 0; i < tcnt; ++i) {
      const int oh0 = th * g.TH, ow0 = tw * g.TW;
      for (int grp = 0; grp < ngroups; ++grp) {
        mbar_wait(&empty_bar[s], ph);
        const int sub0 = grp * G;
        const int cnt = (nsub - sub0) < G ? (nsub - sub0) : G;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], (uint32_t)cnt * (a_box_bytes + b_box_bytes));
          uint8_t* stage = smem + s * g.stage_bytes;
          for (int q = 0; q < cnt; ++q) {
            const int sub = sub0 + q;
            const int tap = sub / chunks, ch = sub - tap * chunks;
            const int c0 = ch * (RB / 2);
            tma_load_4d(stage + q * g.a_sub, &tmX, &full_bar[s], c0, ow0 * g.isw + g.dw[tap], oh0 * g.ish + g.dh[tap], n);
            tma_load_2d(stage + G * g.a_sub + q * g.b_sub, &tmW, &full_bar[s], g.widx[tap] * g.Cin + c0, 0);
          }
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      if (++tw == g.tiles_w) {
        tw = 0;
        if (++th == g.tiles_h) {
          th = 0;
          ++n;
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---- MMA issuer: warp-uniform loop specialised on (k-steps per sub-tile, sub-tiles per stage) ----
    if (RB == 32) generic_mma_role<1, 3, Cout>(g, smem, nsub, ngroups, tmem_base, tcnt, full_bar, empty_bar, tfull_bar, tempty_bar);
    else if (RB == 64) generic_mma_role<2, 3, Cout>(g, smem, nsub, ngroups, tmem_base, tcnt, full_bar, empty_bar, tfull_bar, tempty_bar);
    else generic_mma_role<4, 1, Cout>(g, smem, nsub, ngroups, tmem_base, tcnt, full_bar, empty_bar, tfull_bar, tempty_bar);
  } else {
    // ---- epilogue: thread = one output pixel of the TH x TW patch, its Cout channels split between warps w and w + 4 ----
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;  // tile row = pixel index inside the patch
    const int pr = r / g.TW, pc = r - pr * g.TW;
    EpiRegs<NCHUNK, 0> e;
    epi_init<NCHUNK, 0>(g, hf, e);
    const __nv_bfloat162 floor2 = __float2bfloat162_rn(g.relu ? 0.f : -INFINITY);
    const bool st_ok = !(g.debug & 4);
    int tw = t0 % g.tiles_w, th = (t0 / g.tiles_w) % g.tiles_h, n = t0 / (g.tiles_w * g.tiles_h);
    for (int i = 0; i < tcnt; ++i) {
      const uint32_t a = i & 1, aph = (i >> 1) & 1;
      const int gh = th * g.TH + pr, gw = tw * g.TW + pc;
      const int oh = gh * g.osh + g.oph, ow = gw * g.osw + g.opw;
      const bool okp = pr < g.TH && gh < g.GH && gw < g.GW && oh < g.OH && ow < g.OW;
      const long long offp = (((long long)n * g.OH + oh) * g.OW + ow) * Cout;
      mbar_wait(&tfull_bar[a], aph);
      tc_fence_after();
      // the patch is ONE accumulator block: with Cout = 16 only the hf = 0 warps have a unit
      epi_tile<NCHUNK, EPI, 0>(g, tmem_base + a * (uint32_t)Cout, 1, q, hf, okp ? offp : 0, st_ok, floor2, e,
                               [&](int, bool& ok, int& o) {
                                 ok = okp;
                                 o = 0;
                               });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
      if (++tw == g.tiles_w) {
        tw = 0;
        if (++th == g.tiles_h) {
          th = 0;
          ++n;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Halo variant for C_in <= 64 (the high-resolution layers, which are bound by HBM / L2->SM traffic and by the epilogue, not by
// the tensor cores): per tile of TH grid rows x TW pixels ONE TMA box brings the (TH + box_hr) x (TW + box_wr) input pixels,
// and the taps are UMMA descriptors whose start address is shifted by whole pixel rows inside that box -- the swizzle is a
// function of the shared-memory address, so a row-shifted window of a TMA-written tile is still a valid K-major operand.
// All nine weight taps stay resident in shared memory for the life of the persistent CTA.
// ---------------------------------------------------------------------------------------------------------------
template <int NCHUNK, int EPI, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                const __grid_constant__ CUtensorMap tmW, ConvTcArgs g, int wsub,
                                                                int hsub, int stages) {
  omr_pdl_enter();
  constexpr int Cout = NCHUNK * 16;
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
  const int TH = g.TH, RB = g.rb;
  const uint32_t acc_cols = (uint32_t)(TH * g.ncls * Cout);  // per accumulator buffer: one [128 x Cout] block per (row, class)
  const uint32_t need = 2 * acc_cols;
  const uint32_t tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
  const int pitch = g.TW + g.box_wr;  // pixel rows per input image row inside the halo box
  int t0, tcnt;
  tile_range(g.num_tiles, t0, tcnt);

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);

  if (warp == W_TMA) {
    // ---- TMA producer: the whole warp walks the tiles (uniform state), one elected lane issues ----
    if (elect_one()) {
      mbar_expect_tx(w_full, 9u * (uint32_t)Cout * (uint32_t)RB);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * wsub, &tmW, w_full, t * g.Cin, 0);
    }
    int s = 0;
    uint32_t ph = 1;  // parity to wait for on empty_bar[s]
    int tw = t0 % g.tiles_w, th = (t0 / g.tiles_w) % g.tiles_h, n = t0 / (g.tiles_w * g.tiles_h);
    const uint32_t box_bytes = (uint32_t)(TH + g.box_hr) * (uint32_t)pitch * (uint32_t)RB;
    for (int i = 0; i < tcnt; ++i) {
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        if (g.debug & 2) {
          mbar_arrive(&full_bar[s]);
        } else {
          mbar_expect_tx(&full_bar[s], box_bytes);
          tma_load_4d(sH + s * hsub, &tmX, &full_bar[s], 0, tw * g.TW + g.box_w0, th * TH + g.box_h0, n);
        }
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
      if (++tw == g.tiles_w) {
        tw = 0;
        if (++th == g.tiles_h) {
          th = 0;
          ++n;
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---- MMA issuer: warp-uniform loop specialised on the k-steps per tap (row bytes / 32) ----
    if (RB == 32) halo_mma_role<1, Cout>(g, sW, sH, wsub, hsub, stages, pitch, TH, acc_cols, tmem_base, t0, tcnt, w_full, full_bar, empty_bar, tfull_bar, tempty_bar);
    else if (RB == 64) halo_mma_role<2, Cout>(g, sW, sH, wsub, hsub, stages, pitch, TH, acc_cols, tmem_base, t0, tcnt, w_full, full_bar, empty_bar, tfull_bar, tempty_bar);
    else halo_mma_role<4, Cout>(g, sW, sH, wsub, hsub, stages, pitch, TH, acc_cols, tmem_base, t0, tcnt, w_full, full_bar, empty_bar, tfull_bar, tempty_bar);
  } else {
    // ---- epilogue: thread = grid pixel column px of every row of the tile ----
    const int q = warp & 3, hf = warp >> 2;
    const int px = q * 32 + lane;
    EpiRegs<NCHUNK, MODE> e;
    epi_init<NCHUNK, MODE>(g, hf, e);
    const __nv_bfloat162 floor2 = __float2bfloat162_rn(g.relu ? 0.f : -INFINITY);
    const bool st_ok = !(g.debug & 4);
    const int ncls = g.ncls, cls_shift = ncls == 1 ? 0 : (ncls == 2 ? 1 : 2), cls_mask = ncls - 1;
    const int osh = g.osh, osw = g.osw, OW = g.OW, OH = g.OH, GH = g.GH, GW = g.GW, TW = g.TW;
    // per-class output offsets (elements, relative to the tile's first output pixel of this thread) and coordinates
    int cdo[4], cph[4], cpw[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cph[c] = c < ncls ? g.cls_ph[c] : 0;
      cpw[c] = c < ncls ? g.cls_pw[c] : 0;
      cdo[c] = (cph[c] * OW + cpw[c]) * Cout;
    }
    const int row_do = osh * OW * Cout;  // element offset between consecutive grid rows of the tile
    int tw = t0 % g.tiles_w, th = (t0 / g.tiles_w) % g.tiles_h, n = t0 / (g.tiles_w * g.tiles_h);
    int cur_n = -1;
    for (int i = 0; i < tcnt; ++i) {
      const uint32_t a = i & 1, aph = (i >> 1) & 1;
      if (MODE != 0 && MODE != 2 && n != cur_n) {
        if (cur_n >= 0) epi_flush<NCHUNK, MODE>(g, cur_n, hf, lane, e);
        cur_n = n;
      }
      const int gw = tw * TW + px, gh0 = th * TH;
      const bool okw = px < TW && gw < GW;
      const int ow0 = gw * osw, oh0 = gh0 * osh;
      const long long base = okw ? (((long long)n * OH + oh0) * OW + ow0) * Cout : 0;
      mbar_wait(&tfull_bar[a], aph);
      tc_fence_after();
      // accumulator block rr = (tile row, class): grid pixel (gh0 + row, gw) -> output pixel (osh gh + ph, osw gw + pw)
      epi_tile<NCHUNK, EPI, MODE>(g, tmem_base + a * acc_cols, TH * ncls, q, hf, base, st_ok, floor2, e,
                                  [&](int rr, bool& ok, int& o) {
                                    const int row = rr >> cls_shift, cl = rr & cls_mask;
                                    int ph_ = cph[0], pw_ = cpw[0], d_ = cdo[0];
#pragma unroll
                                    for (int c = 1; c < 4; ++c)
                                      if (cl == c) {
                                        ph_ = cph[c];
                                        pw_ = cpw[c];
                                        d_ = cdo[c];
                                      }
                                    ok = okw && gh0 + row < GH && oh0 + row * osh + ph_ < OH && ow0 + pw_ < OW;
                                    o = row * row_do + d_;
                                  });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
      if (++tw == g.tiles_w) {
        tw = 0;
        if (++th == g.tiles_h) {
          th = 0;
          ++n;
        }
      }
    }
    if (MODE != 0 && (MODE == 2 || cur_n >= 0)) epi_flush<NCHUNK, MODE>(g, cur_n, hf, lane, e);
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
// shared-memory budget of a halo CTA (OMR_CONV_SMEM_KB, default 225): with <= 110 KB two CTAs -- of this kernel or of the other
// encoder's kernel running on its own stream -- share an SM
int conv_smem_cap() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("OMR_CONV_SMEM_KB");
    v = (e ? atoi(e) : 225) * 1024;
    if (v < 64 * 1024) v = 225 * 1024;
  }
  return v;
}
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
int epi_of(const ConvTcArgs& a) { return a.mask ? EPI_MASK : ((a.bias || a.relu) ? EPI_FWD : EPI_PLAIN); }

// ---- launchers: one instantiation per (NCHUNK, EPI, MODE) that the encoders use ------------------------------------
template <int NCHUNK, int EPI>
int launch_generic(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvTcArgs& a, int smem_bytes, cudaStream_t st) {
  auto kern = conv_tc_kernel<NCHUNK, EPI>;
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  OmrLaunch(grid, NTHREADS, smem_bytes, st)(kern, tmX, tmW, a);
  OMR_LAUNCHED();
  return OMR_OK;
}
template <int NCHUNK>
int launch_generic_epi(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvTcArgs& a, int smem_bytes, cudaStream_t st) {
  switch (epi_of(a)) {
    case EPI_FWD: return launch_generic<NCHUNK, EPI_FWD>(tmX, tmW, a, smem_bytes, st);
    case EPI_MASK: return launch_generic<NCHUNK, EPI_MASK>(tmX, tmW, a, smem_bytes, st);
    default: return launch_generic<NCHUNK, EPI_PLAIN>(tmX, tmW, a, smem_bytes, st);
  }
}

template <int NCHUNK, int EPI, int MODE>
int launch_halo(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvTcArgs& h, int wsub, int hsub, int stages, int smem_bytes,
                cudaStream_t st) {
  auto kern = conv_halo_kernel<NCHUNK, EPI, MODE>;
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = h.num_tiles < num_sms() ? h.num_tiles : num_sms();
  OmrLaunch(grid, NTHREADS, smem_bytes, st)(kern, tmX, tmW, h, wsub, hsub, stages);
  OMR_LAUNCHED();
  return OMR_OK;
}
// the (EPI, MODE) pairs that occur: forward (+ InstanceNorm statistics), masked data gradient (+ bias column sums), plain
// data gradient (+ InstanceNorm backward sums)
template <int NCHUNK>
int launch_halo_epi(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvTcArgs& h, int wsub, int hsub, int stages, int smem_bytes,
                    cudaStream_t st) {
  const int epi = epi_of(h), mode = mode_of(h);
  if (epi == EPI_FWD && mode == 0) return launch_halo<NCHUNK, EPI_FWD, 0>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
  if (epi == EPI_MASK && mode == 0) return launch_halo<NCHUNK, EPI_MASK, 0>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
  if (epi == EPI_PLAIN && mode == 0) return launch_halo<NCHUNK, EPI_PLAIN, 0>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
  if constexpr (NCHUNK <= 4) {
    if (epi == EPI_FWD && mode == 1) return launch_halo<NCHUNK, EPI_FWD, 1>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
    if (epi == EPI_MASK && mode == 2) return launch_halo<NCHUNK, EPI_MASK, 2>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
    if (epi == EPI_PLAIN && mode == 3) return launch_halo<NCHUNK, EPI_PLAIN, 3>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
  }
  return OMR_TC_NOT_ELIGIBLE;  // a combination the encoders never produce: the dispatcher falls back to separate passes
}

// One launch of the tap-GEMM.  x: [N, XH, XW, Cin] bf16; wpack: [Cout, 9*Cin] bf16 (tap-major, channels innermost).
int run_taps(const void* x, int N, int XH, int XW, int Cin, const void* wpack, int Cout, ConvTcArgs a, cudaStream_t st) {
  const int rb = (Cin >= 64 ? 64 : Cin) * 2;
  a.debug = conv_debug();
  a.rb = rb;
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
    // TH grid rows per tile: fewer TMA rows per output pixel ((TH+2)/TH instead of 3) and fewer barrier round trips;
    // bounded by TMEM (2 buffers x TH x classes x Cout columns <= 512) and by shared memory (>= 2 halo stages next to the weights)
    int TH = 4, hsub = 0, stages = 0;
    for (; TH >= 1; TH >>= 1) {
      if (2 * TH * a.ncls * Cout > 512 || (TH > 1 && a.GH < TH)) continue;
      int rows = (TH + a.box_hr) * pitch;
      // a tap's 128-row operand window starts up to box_hr * pitch + box_wr rows into the last tile row's box row
      if (rows < (TH - 1 + a.box_hr) * pitch + a.box_wr + 128) rows = (TH - 1 + a.box_hr) * pitch + a.box_wr + 128;
      hsub = (rows * rb + 1023) / 1024 * 1024;
      stages = (conv_smem_cap() - 1024 - 1024 - 9 * wsub) / hsub;
      if (stages > 4) stages = 4;
      if (stages >= 2) break;
    }
    if (TH >= 1 && stages >= 2) {
      ConvTcArgs h = a;
      h.TH = TH; h.TW = TW;
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
      const int smem_bytes = 9 * wsub + stages * hsub + 1024 + 1024;
      switch (Cout) {
        case 16: return launch_halo_epi<1>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
        case 32: return launch_halo_epi<2>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
        case 64: return launch_halo_epi<4>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
        default: return launch_halo_epi<8>(tmX, tmW, h, wsub, hsub, stages, smem_bytes, st);
      }
    }
  }
  if (upsample) return OMR_TC_NOT_ELIGIBLE;  // the caller falls back to one launch per parity class
  if (mode_of(a) != 0) return OMR_TC_NOT_ELIGIBLE;  // the generic kernel has no fused sums: separate pass
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
  // pipeline: G (tap, chunk) sub-tiles per stage, as many stages (<= 4) as fit 160 KB
  a.G = rb == 128 ? 1 : 3;
  a.a_sub = 128 * rb;
  a.b_sub = (128 * rb + 1023) / 1024 * 1024;  // room for Cout <= 128 rows, 1 KB aligned
  a.stage_bytes = a.G * (a.a_sub + a.b_sub);
  a.stages = (a.stage_bytes * 4 <= 160 * 1024) ? 4 : (a.stage_bytes * 3 <= 160 * 1024 ? 3 : 2);
  const int smem_bytes = a.stages * a.stage_bytes + 1024 + 1024;
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
  switch (Cout) {
    case 16: return launch_generic_epi<1>(tmX, tmW, a, smem_bytes, st);
    case 32: return launch_generic_epi<2>(tmX, tmW, a, smem_bytes, st);
    case 64: return launch_generic_epi<4>(tmX, tmW, a, smem_bytes, st);
    default: return launch_generic_epi<8>(tmX, tmW, a, smem_bytes, st);
  }
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
  if (in_sums && (Co > 64 || sh != 1 || sw != 1 || Ci > 64)) return OMR_TC_NOT_ELIGIBLE;
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
  if ((colsum || in_bsums) && (Ci > 64 || Co > 64)) return OMR_TC_NOT_ELIGIBLE;  // fused sums live in the halo kernel
  if (colsum && in_bsums) return OMR_TC_NOT_ELIGIBLE;
  if (colsum && !mask) return OMR_TC_NOT_ELIGIBLE;
  if (in_bsums && mask) return OMR_TC_NOT_ELIGIBLE;
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
  if (colsum || in_bsums) {
    if (sh > 1 || sw > 1) return OMR_TC_NOT_ELIGIBLE;  // parity-class launches carry no fused sums
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
