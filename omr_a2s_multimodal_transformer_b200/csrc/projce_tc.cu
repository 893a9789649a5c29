// projce_tc.cu -- the vocabulary classifier FUSED with the softmax cross-entropy (reference decoder.py:145-146 out_layer +
// model.py:109,166,444,588 CrossEntropyLoss), forward and backward, without ever writing the [B*T, V] logits:
//     forward    S = X W^T (+ bias) tile by tile in TMEM; online (max, sum-exp) per row in the epilogue warps -> row lse,
//                row loss = lse - logit[target]
//     dX         the same S tiles are recomputed; P = exp(S - lse) goes back to TENSOR memory as packed bf16 and is the A
//                operand of  dX += P W_tile  (accumulated over the vocabulary in TMEM);  the one-hot term is subtracted in
//                the final epilogue:  dX[r] = scale * (P W - W[target_r])
//     dW, db     the transposed problem: S^T = W_tile X_blk^T per block of rows, P^T -> TMEM -> dW_tile += P^T X_blk; db is
//                the row sum of P^T; the one-hot term is a tiny scatter kernel
// One kernel template serves all three.  A CTA keeps its "resident" operand R (128 x 256 bf16: a block of rows of X, or a
// vocabulary tile of W) in shared memory and streams tiles T (64 x 256) of the other matrix through a 4-stage ring:
//     S[128 x 64] = R T^T        tcgen05.mma M=128 N=64, 16 x K16          (4 TMEM buffers of 64 columns)
//     E = f(S)                   8 epilogue warps: warp w = TMEM lane quarter w & 3, column half w >> 2 (32 columns a thread)
//     ACC[128 x 256] += E T      A = E from TMEM (packed bf16, written over the S columns the thread has consumed),
//                                B = the same shared-memory T tile read MN-major            (256 TMEM columns)
// Round 1 wrote 16384 x 7040 bf16 logits (230 MB), re-read them for the lse, rewrote them as dlogits and re-read those
// twice (VERDICT "missing 1").  Restricted to d_model = 256 (R must fit next to the ring).
#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int PD = 256;                    // model width (K of the score GEMM, N of the accumulation GEMM)
constexpr int PR = 128, PT_ = 64;          // resident rows, streamed rows per tile
constexpr int P_RCHUNK = PR * 128;         // bytes of a [128 x 64] bf16 chunk (one of the 4 K-chunks of R)
constexpr int P_TCHUNK = PT_ * 128;        // bytes of a [64 x 64] chunk
constexpr int P_TSTAGE = 4 * P_TCHUNK;     // 32 KB
constexpr int P_NST = 4;                   // T stages
constexpr int P_NSB = 4;                   // S buffers in TMEM
constexpr int P_OFF_T = 4 * P_RCHUNK, P_OFF_AUX = P_OFF_T + P_NST * P_TSTAGE, P_OFF_X = P_OFF_AUX + P_NST * 64 * 4,
              P_OFF_BAR = P_OFF_X + 128 * 4 * 4, P_SMEM = P_OFF_BAR + 256;
constexpr float P_LOG2E = 1.4426950408889634f, P_LN2 = 0.6931471805599453f;

struct ProjCeArgs {
  const float* bias;          // [V] or NULL
  const long long* targets;   // [M]
  long long ignore_index;
  int M, V;
  int tiles_per_item;         // MODE 2: blocks of 64 rows per work item
  // MODE 0 outputs
  float* row_loss; float* row_lse;
  // MODE 1 / 2 inputs
  const float* row_lse_in;    // [M] natural log
  const float* loss_out;      // [2]: (mean loss, n_valid)
  const float* gscale;        // [1] upstream gradient
  // MODE 1
  bf16* dx; long long dx_ld;
  const bf16* w; long long w_ld;  // the classifier matrix again, for the one-hot term
  // MODE 2
  float* dw; float* db;       // fp32 [V, 256] / [V], accumulated into
};

__device__ __forceinline__ float p_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float p_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void p_red4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// MODE 0: forward (lse, loss)   1: dX   2: dW, db
template <int MODE>
__global__ void __launch_bounds__(320, 1) projce_kernel(const __grid_constant__ CUtensorMap tmR,
                                                        const __grid_constant__ CUtensorMap tmT, ProjCeArgs g) {
  omr_pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* sR = smem;                                                  // 4 K-chunks of [128 x 64]
  uint8_t* sT = smem + P_OFF_T;                                        // [P_NST] x 4 K-chunks of [64 x 64]
  float* sAux = reinterpret_cast<float*>(smem + P_OFF_AUX);            // [P_NST][64] per-column term of a T tile (log2 units)
  float* sX = reinterpret_cast<float*>(smem + P_OFF_X);                // [128][4] end-of-item exchange between the column halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_OFF_BAR);
  uint64_t* r_full = bars;
  uint64_t* t_full = bars + 1;                  // [P_NST]
  uint64_t* t_empty = bars + 1 + P_NST;         // [P_NST]
  uint64_t* s_full = bars + 1 + 2 * P_NST;      // [P_NSB] scores of a tile in TMEM buffer i % P_NSB
  uint64_t* e_done = s_full + P_NSB;            // [P_NSB] epilogue warps are done with the buffer (and have written E) (count 8)
  uint64_t* acc_done = e_done + P_NSB;          // all MMAs of the item complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  // the work item
  int r0, t_begin, t_end;  // first resident row; streamed tiles [t_begin, t_end) of 64 rows
  if (MODE == 2) {
    const int nsplit = gridDim.y;
    r0 = blockIdx.x * PR;  // vocabulary tile
    const int nrb = (g.M + PT_ - 1) / PT_;
    t_begin = blockIdx.y * g.tiles_per_item;
    t_end = t_begin + g.tiles_per_item < nrb ? t_begin + g.tiles_per_item : nrb;
    (void)nsplit;
  } else {
    r0 = blockIdx.x * PR;  // block of rows
    t_begin = 0;
    t_end = (g.V + PT_ - 1) / PT_;
  }
  const int ntiles = t_end - t_begin;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmT);
    mbar_init(r_full, 1);
    for (int s = 0; s < P_NST; ++s) {
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 1);
    }
    for (int s = 0; s < P_NSB; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&e_done[s], 8);
    }
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t tmem_S = tmem_base, tmem_ACC = tmem_base + 256;

  if (warp == 8) {
    // ---- producer: R once, then the T tiles with their per-column term ----
    if (ntiles > 0) {
      if (elect_one()) {
        mbar_expect_tx(r_full, 4 * P_RCHUNK);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) tma_load_2d(sR + kc * P_RCHUNK, &tmR, r_full, kc * 64, r0);
      }
      __syncwarp();
      for (int i = 0; i < ntiles; ++i) {
        const int s = i % P_NST, c0 = (t_begin + i) * PT_;
        // per-column term (log2 units).  MODE 0/1: columns are classes: bias * log2e, -inf past V (those columns vanish);
        // MODE 2: columns are rows of X: -lse * log2e, -inf for ignored / out-of-range rows (their probabilities vanish)
        float aux[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c = c0 + k * 32 + lane;
          if (MODE == 2) {
            const bool ok = c < g.M && g.targets[c] != g.ignore_index;
            aux[k] = ok ? -g.row_lse_in[c] * P_LOG2E : -INFINITY;
          } else {
            aux[k] = c < g.V ? (g.bias ? g.bias[c] * P_LOG2E : 0.f) : -INFINITY;
          }
        }
        mbar_wait(&t_empty[s], ((i / P_NST) & 1) ^ 1);
        sAux[s * 64 + lane] = aux[0];
        sAux[s * 64 + 32 + lane] = aux[1];
        __syncwarp();
        if (elect_one()) {
          mbar_expect_tx(&t_full[s], P_TSTAGE);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) tma_load_2d(sT + s * P_TSTAGE + kc * P_TCHUNK, &tmT, &t_full[s], kc * 64, c0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ---- MMA issuer: scores of tile i as soon as its T tile has landed and S buffer i % 4 is free; the accumulation MMA of
    // tile i - 1 once the epilogue warps have written E(i - 1).  In-order execution of this thread's MMAs makes the reuse of
    // an S buffer by the scores of tile i + 4 safe against the accumulation MMA of tile i that reads E from it. ----
    if (ntiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 256, 0, 1);  // A from TMEM, B MN-major
      mbar_wait(r_full, 0);
      const uint32_t r_addr = smem_u32(sR), t_base = smem_u32(sT);
      auto issue_acc = [&](int i) {  // ACC += E(i) T(i)
        const int s = i % P_NST, b = i % P_NSB;
        mbar_wait(&e_done[b], (i / P_NSB) & 1);
        tc_fence_after();
        const uint32_t t_addr = t_base + (uint32_t)s * P_TSTAGE, e_tmem = tmem_S + (uint32_t)b * 64;
        if (elect_one()) {
          if (MODE != 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)  // K16 slices of the 64 streamed rows; E columns: halves at +0 and +32, 8 per slice
              umma_bf16_ta(tmem_ACC, e_tmem + (j >> 1) * 32 + (j & 1) * 8, make_smem_desc(t_addr + j * 2048, P_TCHUNK, 1024, 128), idesc_acc,
                           (i > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(&t_empty[s]);
        }
        __syncwarp();
      };
      for (int i = 0; i < ntiles; ++i) {
        const int s = i % P_NST, b = i % P_NSB;
        mbar_wait(&t_full[s], (i / P_NST) & 1);
        if (i >= P_NSB) mbar_wait(&e_done[b], ((i / P_NSB) - 1) & 1);  // (already seen by issue_acc(i - 4); kept for MODE 0 clarity)
        tc_fence_after();
        const uint32_t t_addr = t_base + (uint32_t)s * P_TSTAGE, d_tmem = tmem_S + (uint32_t)b * 64;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const uint32_t kc = kk >> 2, j = kk & 3;
            umma_bf16(d_tmem, make_smem_desc(r_addr + kc * P_RCHUNK + j * 32, 16, 1024, 128),
                      make_smem_desc(t_addr + kc * P_TCHUNK + j * 32, 16, 1024, 128), idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[b]);
        }
        __syncwarp();
        if (i >= 1) issue_acc(i - 1);
      }
      issue_acc(ntiles - 1);
      if (elect_one()) umma_commit(acc_done);
      __syncwarp();
    }
  } else {
    // ---- 8 epilogue warps: thread = (resident row lq * 32 + lane, column half hf) ----
    const int lq = warp & 3, hf = warp >> 2;
    const int rr = lq * 32 + lane;     // row inside R
    const int row = r0 + rr;           // MODE 0/1: row of X; MODE 2: class
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    // per-row constants
    float row_term = 0.f;              // MODE 1: -lse (log2); MODE 2: bias (log2); MODE 0: unused
    long long tgt = -1;
    bool row_ok = false;
    if (MODE == 2) {
      row_ok = row < g.V;
      row_term = row_ok ? (g.bias ? g.bias[row] * P_LOG2E : 0.f) : -INFINITY;
    } else {
      if (row < g.M) {
        tgt = g.targets[row];
        row_ok = tgt != g.ignore_index;
      }
      if (MODE == 1) row_term = row_ok ? -g.row_lse_in[row] * P_LOG2E : -INFINITY;
    }
    float m_run = -INFINITY, l_run = 0.f, x_t = 0.f, psum = 0.f;
    for (int i = 0; i < ntiles; ++i) {
      const int s = i % P_NST, b = i % P_NSB;
      const int c0 = (t_begin + i) * PT_ + hf * 32;  // first streamed row (= column of S) of this thread
      mbar_wait(&s_full[b], (i / P_NSB) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_S + (uint32_t)b * 64 + lane_addr + hf * 32, v);
      tmem_ld_wait();
      const float* aux = sAux + s * 64 + hf * 32;
      float x[32];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(aux + e);
        x[e] = fmaf(__uint_as_float(v[e]), P_LOG2E, a4.x + row_term);
        x[e + 1] = fmaf(__uint_as_float(v[e + 1]), P_LOG2E, a4.y + row_term);
        x[e + 2] = fmaf(__uint_as_float(v[e + 2]), P_LOG2E, a4.z + row_term);
        x[e + 3] = fmaf(__uint_as_float(v[e + 3]), P_LOG2E, a4.w + row_term);
      }
      if (MODE == 0) {
        // online log-sum-exp over this thread's columns (log2 units); the target's logit when it passes by
        float mx = m_run;
#pragma unroll
        for (int e = 0; e < 32; e += 2) mx = p_max3(mx, x[e], x[e + 1]);
        if (mx > m_run) {
          l_run *= p_ex2(m_run - mx);  // m_run = -inf: l_run is 0 and stays 0
          m_run = mx;
        }
        if (mx > -INFINITY) {
          float sum = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) sum += p_ex2(x[e] - m_run);
          l_run += sum;
        }
        if (row_ok && tgt >= c0 && tgt < c0 + 32) {
          const int te = (int)(tgt - c0);
#pragma unroll
          for (int e = 0; e < 32; ++e) x_t = (e == te) ? x[e] : x_t;
        }
      } else {
        // E = exp2(x): probabilities (the row / column terms already carry -lse and the bias); packed bf16 back into the S
        // columns this thread has just consumed: 16 columns at offset hf * 32 of the buffer
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = p_ex2(x[e]), p1 = p_ex2(x[e + 1]);
          if (MODE == 2) psum += p0 + p1;
          pk[e >> 1] = pack_bf16(p0, p1);
        }
        tmem_st16(tmem_S + (uint32_t)b * 64 + lane_addr + hf * 32, pk);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&e_done[b]);
    }
    // ---- end of item ----
    if (MODE == 0) {
      // combine the two column halves of a row: lse = log(l_a 2^m_a + l_b 2^m_b), loss = lse - logit[target]
      sX[rr * 4 + hf * 2] = m_run;
      sX[rr * 4 + hf * 2 + 1] = l_run;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float m_o = sX[rr * 4 + (hf ^ 1) * 2], l_o = sX[rr * 4 + (hf ^ 1) * 2 + 1];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool has_t = row_ok && tgt >= 0 && ((tgt / 32) & 1) == hf && tgt < g.V;
      sX[rr * 4 + hf] = has_t ? x_t : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hf == 0 && row < g.M) {
        const float mm = fmaxf(m_run, m_o);
        const float lt = (m_run > -INFINITY ? l_run * p_ex2(m_run - mm) : 0.f) + (m_o > -INFINITY ? l_o * p_ex2(m_o - mm) : 0.f);
        const float lse2 = mm + log2f(lt);
        const float xt2 = sX[rr * 4] + sX[rr * 4 + 1];
        g.row_lse[row] = lse2 * P_LN2;
        g.row_loss[row] = row_ok ? (lse2 - xt2) * P_LN2 : 0.f;
      }
    } else if (ntiles > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
      const float nv = g.loss_out[1];
      const float scale = nv > 0.f ? g.gscale[0] / nv : 0.f;
      if (MODE == 1) {
        // dX[row, 128 hf .. 128 hf + 128) = scale * (ACC - W[target])
        const bool live = row < g.M;
        bf16* dst = g.dx + (long long)row * g.dx_ld + hf * 128;
        const bf16* wt = (row_ok && tgt >= 0 && tgt < g.V) ? g.w + tgt * g.w_ld + hf * 128 : nullptr;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_ACC + lane_addr + hf * 128 + c * 32, v);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float w8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              if (wt) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(wt + c * 32 + u * 8);
                const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&ww[k]);
                  w8[2 * k] = __low2float(h2);
                  w8[2 * k + 1] = __high2float(h2);
                }
              }
              uint4 o4;
              o4.x = pack_bf16((__uint_as_float(v[8 * u]) - w8[0]) * scale, (__uint_as_float(v[8 * u + 1]) - w8[1]) * scale);
              o4.y = pack_bf16((__uint_as_float(v[8 * u + 2]) - w8[2]) * scale, (__uint_as_float(v[8 * u + 3]) - w8[3]) * scale);
              o4.z = pack_bf16((__uint_as_float(v[8 * u + 4]) - w8[4]) * scale, (__uint_as_float(v[8 * u + 5]) - w8[5]) * scale);
              o4.w = pack_bf16((__uint_as_float(v[8 * u + 6]) - w8[6]) * scale, (__uint_as_float(v[8 * u + 7]) - w8[7]) * scale);
              *reinterpret_cast<uint4*>(dst + c * 32 + u * 8) = o4;
            }
          }
        }
      } else {
        // dW[class, 128 hf ..) += scale * ACC ; db[class] += scale * row sum of the probabilities (this column half's share)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_ACC + lane_addr + hf * 128 + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float* dst = g.dw + (long long)row * PD + hf * 128 + c * 32;
#pragma unroll
            for (int u = 0; u < 8; ++u)
              p_red4(dst + 4 * u, __uint_as_float(v[4 * u]) * scale, __uint_as_float(v[4 * u + 1]) * scale,
                     __uint_as_float(v[4 * u + 2]) * scale, __uint_as_float(v[4 * u + 3]) * scale);
          }
        }
        if (row_ok && g.db) atomicAdd(g.db + row, psum * scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// the one-hot term of the weight / bias gradient: dW[target_r] -= scale x_r, db[target_r] -= scale (one warp per row)
__global__ void projce_onehot_kernel(const bf16* __restrict__ x, long long x_ld, const long long* __restrict__ targets,
                                     long long ignore_index, const float* __restrict__ loss_out,
                                     const float* __restrict__ gscale, float* __restrict__ dw, float* __restrict__ db, int M,
                                     int V) {
  omr_pdl_enter();
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= M) return;
  const long long t = targets[r];
  if (t == ignore_index || t < 0 || t >= V) return;
  const float nv = loss_out[1];
  const float scale = nv > 0.f ? gscale[0] / nv : 0.f;
  const uint4 x4 = *reinterpret_cast<const uint4*>(x + r * x_ld + lane * 8);
  const uint32_t xx[4] = {x4.x, x4.y, x4.z, x4.w};
  float* dst = dw + t * PD + lane * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&xx[k]);
    atomicAdd(dst + 2 * k, -scale * __low2float(h2));
    atomicAdd(dst + 2 * k + 1, -scale * __high2float(h2));
  }
  if (lane == 0 && db) atomicAdd(db + t, -scale);
}

int make_mat_map(CUtensorMap* m, const void* base, long long rows, long long ld, int box_rows) {
  unsigned long long dims[2] = {(unsigned long long)PD, (unsigned long long)rows};
  unsigned long long strides[1] = {(unsigned long long)ld * 2};
  unsigned int box[2] = {64u, (unsigned)box_rows};
  return omr_make_tensor_map(m, 2, base, 2, dims, strides, box, nullptr, 128);
}
bool pal(const void* p, long long ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * 2) % 16 == 0; }

template <int MODE>
int launch(const CUtensorMap& tmR, const CUtensorMap& tmT, const ProjCeArgs& g, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(projce_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    configured = true;
  }
  OmrLaunch(grid, 320, P_SMEM, st)(projce_kernel<MODE>, tmR, tmT, g);
  OMR_LAUNCHED();
  return OMR_OK;
}

}  // namespace

// x [M, 256] bf16 (row stride x_ld), w [V, 256] bf16 (row stride w_ld), bias fp32 [V] or NULL
int omr_proj_ce_fwd_tc(const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                       const long long* targets, long long M, int V, int D, long long ignore_index, float* row_loss,
                       float* row_lse, cudaStream_t st) {
  if (D != PD || M < 1 || V < 1 || M > 0x7fffffffll || !pal(x, x_ld) || !pal(w, w_ld)) return OMR_TC_NOT_ELIGIBLE;
  CUtensorMap tmR, tmT;
  int rc = make_mat_map(&tmR, x, M, x_ld, PR);
  if (rc) return rc;
  rc = make_mat_map(&tmT, w, V, w_ld, PT_);
  if (rc) return rc;
  ProjCeArgs g{};
  g.bias = bias; g.targets = targets; g.ignore_index = ignore_index; g.M = (int)M; g.V = V;
  g.row_loss = row_loss; g.row_lse = row_lse;
  return launch<0>(tmR, tmT, g, dim3((unsigned)((M + PR - 1) / PR)), st);
}

int omr_proj_ce_bwd_tc(const void* x, long long x_ld, const void* w, long long w_ld, const float* bias,
                       const long long* targets, const float* row_lse, const float* loss_out, const float* gscale,
                       long long M, int V, int D, long long ignore_index, void* dx, long long dx_ld, float* dw, float* db,
                       cudaStream_t st_dx, cudaStream_t st_dw) {
  if (D != PD || M < 1 || V < 1 || M > 0x7fffffffll || !pal(x, x_ld) || !pal(w, w_ld) || (dx && !pal(dx, dx_ld)) ||
      (dw && (reinterpret_cast<uintptr_t>(dw) & 15)))
    return OMR_TC_NOT_ELIGIBLE;
  ProjCeArgs g{};
  g.bias = bias; g.targets = targets; g.ignore_index = ignore_index; g.M = (int)M; g.V = V;
  g.row_lse_in = row_lse; g.loss_out = loss_out; g.gscale = gscale;
  if (dx) {
    CUtensorMap tmR, tmT;
    int rc = make_mat_map(&tmR, x, M, x_ld, PR);
    if (rc) return rc;
    rc = make_mat_map(&tmT, w, V, w_ld, PT_);
    if (rc) return rc;
    ProjCeArgs a = g;
    a.dx = (bf16*)dx; a.dx_ld = dx_ld; a.w = (const bf16*)w; a.w_ld = w_ld;
    rc = launch<1>(tmR, tmT, a, dim3((unsigned)((M + PR - 1) / PR)), st_dx);
    if (rc) return rc;
  }
  if (dw) {
    CUtensorMap tmR, tmT;
    int rc = make_mat_map(&tmR, w, V, w_ld, PR);
    if (rc) return rc;
    rc = make_mat_map(&tmT, x, M, x_ld, PT_);
    if (rc) return rc;
    ProjCeArgs a = g;
    a.dw = dw; a.db = db;
    const int nvt = (V + PR - 1) / PR;
    const int nrb = (int)((M + PT_ - 1) / PT_);
    int nsm = 148;
    {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
      if (nsm <= 0) nsm = 148;
    }
    // work items = (vocabulary tile, chunk of row blocks): about three full waves of CTAs
    int nsplit = (3 * nsm) / nvt;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > nrb) nsplit = nrb;
    a.tiles_per_item = (nrb + nsplit - 1) / nsplit;
    nsplit = (nrb + a.tiles_per_item - 1) / a.tiles_per_item;
    rc = launch<2>(tmR, tmT, a, dim3((unsigned)nvt, (unsigned)nsplit), st_dw);
    if (rc) return rc;
    OmrLaunch((unsigned)((M * 32 + 255) / 256), 256, 0, st_dw)(projce_onehot_kernel, (const bf16*)x, x_ld, targets, ignore_index, loss_out,
                                                              gscale, dw, db, (int)M, V);
    OMR_LAUNCHED();
  }
  return OMR_OK;
}
