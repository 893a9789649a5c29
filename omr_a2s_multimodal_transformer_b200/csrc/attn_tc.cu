// attn_tc.cu -- flash-style attention forward (head_dim 64, bf16) on the tcgen05 tensor cores.
//
// One CTA = one (batch, head, 128-query tile); it walks the visible 128-key tiles:
//     S  = Q K^T          tcgen05.mma  M=128 (queries) N=128 (keys) K=64,  fp32 S in TMEM
//     P  = online softmax of  scale*S + key_bias (+ causal / sliding-window structure)   -- 4 softmax warps,
//                         one query row per thread, exp2 with the running row max, bf16 P written to shared memory
//                         in the 128B-swizzled K-major operand layout
//     O += P V            tcgen05.mma  M=128 N=64 K=128 (V rows are the reduction index: MN-major B), fp32 in TMEM,
//                         folded into the per-thread fp32 output row with the usual rescaling
// Q/K/V tiles arrive by TMA straight out of the packed projection buffers ([B,T,3D] / [B,S,2D]; the head is a
// column offset), K/V double buffered.  Two CTAs fit on an SM (112 KB smem, 256 TMEM columns each), so one CTA's
// softmax overlaps the other's MMAs.  The mask algebra is the reference's: additive fp32
// key bias per (b, key) (0, +1.0 or -inf), keys visible to query t iff k <= t + Tk - Tq (causal) and
// k >= t + Tk - Tq - window.  lse (natural log) is saved for the backward.
// Warp roles: 0-3 softmax + epilogue (TMEM lanes 32w..32w+31), 4 TMA producer, 5 MMA issuer + TMEM allocator.
#include <type_traits>

#include "attn_drop.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE = 128 * 128;  // bytes of a [128 x 64] bf16 tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnTcArgs {
  bf16* o;
  long long o_bs, o_rs;
  float* lse;
  const float* key_bias;
  int B, H, Tq, Tk;
  float scale_log2;  // scale * log2(e)
  int causal, window;
  AttnDrop drop;     // attention-probability dropout (thr == 0: off)
  // mixer block mask (reference model.py:340-352): pairs with query >= q_len[s] AND key >= kv_len[s] are excluded, where
  // s = (b * H + h) % quirk_mod reproduces the reference's head-major mask repeat (quirk_mod = 0: s = b); NULL: none
  const int* q_len;
  const int* kv_len;
  int quirk_mod;
};
__device__ __forceinline__ void block_mask_of(const AttnTcArgs& a, int b, int h, int& lq, int& lkv) {
  lq = a.Tq;
  lkv = a.Tk;
  if (a.q_len && a.kv_len) {
    const int s = a.quirk_mod > 0 ? (b * a.H + h) % a.quirk_mod : b;
    lq = a.q_len[s];
    lkv = a.kv_len[s];
  }
}

// Q, K[2], V[2], P (2 tiles) + bias tile + barriers: 115,456 B, so that two CTAs (+1 KB system reserve each) fit
// in the 228 KB of an SM; the dynamic window is declared 1024-aligned (no static shared memory in this kernel)
constexpr int FWD_SMEM = TILE * 7 + 512 + 256;

__device__ __forceinline__ void kv_tile_range(const AttnTcArgs& a, int q0, int& kt0, int& kt1) {
  const int nkt = (a.Tk + BKV - 1) / BKV;
  kt0 = 0; kt1 = nkt;
  if (a.causal) {
    const int off = a.Tk - a.Tq;
    int t_last = q0 + BQ - 1;
    if (t_last > a.Tq - 1) t_last = a.Tq - 1;
    const int jmax = t_last + off;
    if (jmax < 0) { kt1 = 0; return; }
    const int e = jmax / BKV + 1;
    if (e < kt1) kt1 = e;
    if (a.window > 0) {
      const int jmin = q0 + off - a.window;
      if (jmin > 0) kt0 = jmin / BKV;
    }
  }
}

__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void softmax_bar2() { asm volatile("bar.sync 1, 256;" ::: "memory"); }


__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- the per-tile work of a softmax thread (one query row, 128 keys of S in TMEM), specialised so that the common
// tile -- every key visible to every row -- pays for no interval tests, and a call without key bias for no bias loads.
// x = scale_log2 * S + bias (log2 units).  MASKED tiles (causal diagonal, window edge, ragged key tail) test the row's
// visible interval [k_lo, k_hi] per key.
struct SmRow {
  uint32_t tmem_s;      // TMEM address of this row's S values
  const float* bias;    // shared-memory bias tile of the 128 keys (pre-multiplied by log2 e)
  float scale_log2;
  int j0, k_lo, k_hi;   // first key of the tile, visible interval of the row
};

template <bool MASKED, bool BIAS>
__device__ __forceinline__ float softmax_row_max(const SmRow& w) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(w.tmem_s + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      float bq[4] = {0.f, 0.f, 0.f, 0.f};
      if (BIAS || MASKED) {
        const float4 b4 = *reinterpret_cast<const float4*>(w.bias + c * 32 + e);
        bq[0] = b4.x; bq[1] = b4.y; bq[2] = b4.z; bq[3] = b4.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = w.j0 + c * 32 + e + k;
        float x = __uint_as_float(v[e + k]);
        if (BIAS || MASKED) x = fmaf(x, w.scale_log2, bq[k]);  // otherwise the (positive) scale is applied to the max
        if (MASKED) x = (j >= w.k_lo && j <= w.k_hi) ? x : -INFINITY;
        mx = fmaxf(mx, x);
      }
    }
  }
  return (BIAS || MASKED) ? mx : mx * w.scale_log2;
}

// P = exp2(x - m) -> row sum (returned) and the bf16 tile in the 128B-swizzled K-major operand layout; with DROP the
// stored probabilities carry the keep mask / (1-p) while the sum stays that of the full row
template <bool MASKED, bool BIAS, bool DROP>
__device__ __forceinline__ float softmax_row_exp(const SmRow& w, float m_safe, uint8_t* sP, int r, int t, uint32_t dstream,
                                                 uint32_t dkp, uint32_t thr, float inv_keep) {
  float rs = 0.f;
  const float nm = -m_safe;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(w.tmem_s + c * 32, v);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      float bq[4] = {nm, nm, nm, nm};
      if (BIAS || MASKED) {
        const float4 b4 = *reinterpret_cast<const float4*>(w.bias + c * 32 + e);
        bq[0] += b4.x; bq[1] += b4.y; bq[2] += b4.z; bq[3] += b4.w;
      }
      float pr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = w.j0 + c * 32 + e + k;
        pr[k] = ex2_approx(fmaf(__uint_as_float(v[e + k]), w.scale_log2, bq[k]));
        if (MASKED) pr[k] = (j >= w.k_lo && j <= w.k_hi) ? pr[k] : 0.f;
        rs += pr[k];
      }
      if (DROP) {
        // rows t and t^1 (neighbouring lanes) share their 2 x 2 blocks: the even lane hashes the block of keys
        // (j, j+1), the odd lane that of (j+2, j+3), and each hands the other the word of its row parity
        const int j = w.j0 + c * 32 + e;
        const uint2 mine = attn_drop_block(dstream, (uint32_t)(t >> 1), (uint32_t)((j >> 1) + (t & 1)), dkp);
        const uint32_t ox = __shfl_xor_sync(0xffffffffu, mine.x, 1), oy = __shfl_xor_sync(0xffffffffu, mine.y, 1);
        const uint32_t w0 = (t & 1) ? oy : mine.x;  // keys (j, j+1) for this row
        const uint32_t w1 = (t & 1) ? mine.y : ox;  // keys (j+2, j+3)
        pr[0] = (w0 & 0xFFFFu) >= thr ? pr[0] * inv_keep : 0.f;
        pr[1] = (w0 >> 16) >= thr ? pr[1] * inv_keep : 0.f;
        pr[2] = (w1 & 0xFFFFu) >= thr ? pr[2] * inv_keep : 0.f;
        pr[3] = (w1 >> 16) >= thr ? pr[3] * inv_keep : 0.f;
      }
      pk[e >> 1] = pack_bf16(pr[0], pr[1]);
      pk[(e >> 1) + 1] = pack_bf16(pr[2], pr[3]);
    }
    // keys c*32 .. c*32+31 = 64 bytes = 4 sixteen-byte units of row r in chunk c/2
    uint8_t* rowp = sP + (c >> 1) * TILE + r * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int unit = (c & 1) * 4 + u;
      *reinterpret_cast<uint4*>(rowp + ((unit ^ (r & 7)) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
    }
  }
  return rs;
}

__global__ void __launch_bounds__(192, 2) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                             const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV, AttnTcArgs a) {
  omr_pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();  // the swizzled tiles need 1 KB alignment
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE;      // [2]
  uint8_t* sV = smem + 3 * TILE;  // [2]
  uint8_t* sP = smem + 5 * TILE;  // two 64-key chunks
  float* sBias = reinterpret_cast<float*>(smem + 7 * TILE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TILE + 512);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* pv_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint32_t* live = reinterpret_cast<uint32_t*>(bars + 9);  // [32] bit i: key tile kt0 + i holds a key that is not masked out

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  int kt0, kt1;
  kv_tile_range(a, q0, kt0, kt1);
  const int ntiles = kt1 - kt0;
  // Key tiles whose every key carries bias = -inf (the padded tail of the memory, reference decoder.py:150-189 with
  // the concat mixer's bool mask) contribute exactly zero: they are skipped by all three roles.
  const bool use_live = a.key_bias != nullptr && ntiles <= 1024;
  if (use_live) {
    for (int i = threadIdx.x; i < 32; i += blockDim.x) live[i] = 0u;
    __syncthreads();
    const float* kbp = a.key_bias + (long long)b * a.Tk;
    const int jend = kt1 * BKV < a.Tk ? kt1 * BKV : a.Tk;
    for (int j = kt0 * BKV + (int)threadIdx.x; j < jend; j += blockDim.x)
      if (kbp[j] > -INFINITY) atomicOr(&live[(j / BKV - kt0) >> 5], 1u << ((j / BKV - kt0) & 31));
  }
  auto tile_live = [&](int i) { return !use_live || ((live[i >> 5] >> (i & 31)) & 1u); };

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t tmem_S = tmem_base, tmem_PV = tmem_base + 128;

  // Producer and MMA issuer run their loops WARP-wide on uniform values and issue under elect_one() (tc_common.cuh): a
  // lone lane under `if (lane == 0)` made the compiler wrap every tcgen05.mma / TMA in an ELECT / R2UR / BRA.U.ANY waterfall.
  if (warp == 4) {
    if (ntiles > 0) {
      if (elect_one()) {
        mbar_expect_tx(q_full, TILE);
        tma_load_3d(sQ, &tmQ, q_full, h * HD, q0, b);
      }
      for (int i = 0, n = 0; i < ntiles; ++i) {
        if (!bcast0(tile_live(i) ? 1u : 0u)) continue;
        const int s = n & 1;
        mbar_wait(&kv_empty[s], ((n >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[s], 2 * TILE);
          tma_load_3d(sK + s * TILE, &tmK, &kv_full[s], h * HD, (kt0 + i) * BKV, b);
          tma_load_3d(sV + s * TILE, &tmV, &kv_full[s], h * HD, (kt0 + i) * BKV, b);
        }
        __syncwarp();
        ++n;
      }
    }
  } else if (warp == 5) {
    if (ntiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      mbar_wait(q_full, 0);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP), k_base = smem_u32(sK), v_base = smem_u32(sV);
      for (int i = 0, n = 0; i < ntiles; ++i) {
        if (!bcast0(tile_live(i) ? 1u : 0u)) continue;
        const int s = n & 1;
        mbar_wait(&kv_full[s], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = k_base + (uint32_t)s * TILE, v_addr = v_base + (uint32_t)s * TILE;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j > 0)
              umma_bf16_acc(tmem_S, make_smem_desc(q_addr + j * 32, 16, 1024, 128), make_smem_desc(k_addr + j * 32, 16, 1024, 128), idesc_s);
            else
              umma_bf16_new(tmem_S, make_smem_desc(q_addr, 16, 1024, 128), make_smem_desc(k_addr, 16, 1024, 128), idesc_s);
          }
          umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_full, n & 1);
        ++n;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j > 0)
              umma_bf16_acc(tmem_PV, make_smem_desc(p_addr + (j >> 2) * TILE + (j & 3) * 32, 16, 1024, 128),
                            make_smem_desc(v_addr + j * 2048, 0, 1024, 128), idesc_pv);
            else
              umma_bf16_new(tmem_PV, make_smem_desc(p_addr, 16, 1024, 128), make_smem_desc(v_addr, 0, 1024, 128), idesc_pv);
          }
          umma_commit(pv_full);
          umma_commit(&kv_empty[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- softmax + epilogue: thread = query row ----
    const int r = warp * 32 + lane;
    const int t = q0 + r;
    const int off = a.Tk - a.Tq;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o_acc[d] = 0.f;
    const float* kb = a.key_bias ? a.key_bias + (long long)b * a.Tk : nullptr;
    const uint32_t dstream = a.drop.thr ? attn_drop_stream(a.drop, b * a.H + h) : 0u;
    const uint32_t dkp = (uint32_t)((a.Tk + 1) >> 1);
    // visible key interval of this row
    int k_hi = a.Tk - 1, k_lo = 0;
    if (a.causal) {
      k_hi = min(k_hi, t + off);
      if (a.window > 0) k_lo = max(0, t + off - a.window);
    }
    if (t >= a.Tq) k_hi = -1;
    int lq, lkv;
    block_mask_of(a, b, h, lq, lkv);
    if (t >= lq) k_hi = min(k_hi, lkv - 1);
    int n = 0;  // live tiles processed so far (the barrier phases count these)
    for (int i = 0; i < ntiles; ++i) {
      if (!tile_live(i)) continue;
      const int j0 = (kt0 + i) * BKV;
      // key-bias tile (pre-multiplied by log2 e) -> smem, shared by the 128 rows
      softmax_bar();  // previous tile's readers are done with sBias
      {
        const int j = j0 + r;
        sBias[r] = (kb && j < a.Tk) ? kb[j] * LOG2E : 0.f;
      }
      softmax_bar();
      mbar_wait(s_full, n & 1);
      tc_fence_after();
      // does every row of this CTA see every key of the tile?  (rows past Tq are never stored: they may see anything)
      bool full = j0 + BKV <= a.Tk;
      if (a.causal) {
        full = full && (j0 + BKV - 1 <= q0 + off) && (a.window <= 0 || j0 >= q0 + BQ - 1 + off - a.window);
      }
      full = full && (q0 + BQ <= lq || j0 + BKV <= lkv);  // block mask: no masked row in this CTA, or the tile lies below the cut
      const SmRow w{tmem_S + lane_addr, sBias, a.scale_log2, j0, k_lo, k_hi};
      // pass 1: row max
      float mx;
      if (!full) mx = softmax_row_max<true, true>(w);
      else if (kb) mx = softmax_row_max<false, true>(w);
      else mx = softmax_row_max<false, false>(w);
      const float m_new = fmaxf(m_run, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_safe);
      if (n > 0) {
        mbar_wait(pv_full, (n - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_PV + lane_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o_acc[c * 32 + e] = (o_acc[c * 32 + e] + __uint_as_float(v[e])) * alpha;
        }
      }
      // pass 2: P = exp2(x - m), row sum, bf16 into the swizzled A-operand tile
      float rs;
      const uint32_t thr = a.drop.thr;
      const float ik = a.drop.inv_keep;
      if (thr) {
        if (!full) rs = softmax_row_exp<true, true, true>(w, m_safe, sP, r, t, dstream, dkp, thr, ik);
        else if (kb) rs = softmax_row_exp<false, true, true>(w, m_safe, sP, r, t, dstream, dkp, thr, ik);
        else rs = softmax_row_exp<false, false, true>(w, m_safe, sP, r, t, dstream, dkp, thr, ik);
      } else {
        if (!full) rs = softmax_row_exp<true, true, false>(w, m_safe, sP, r, t, dstream, dkp, thr, ik);
        else if (kb) rs = softmax_row_exp<false, true, false>(w, m_safe, sP, r, t, dstream, dkp, thr, ik);
        else rs = softmax_row_exp<false, false, false>(w, m_safe, sP, r, t, dstream, dkp, thr, ik);
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      ++n;
    }
    if (n > 0) {
      mbar_wait(pv_full, (n - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_PV + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o_acc[c * 32 + e] += __uint_as_float(v[e]);
      }
    }
    if (t < a.Tq) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* op = a.o + (long long)b * a.o_bs + (long long)t * a.o_rs + h * HD;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint4 o4;
        o4.x = pack_bf16(o_acc[8 * u] * inv, o_acc[8 * u + 1] * inv);
        o4.y = pack_bf16(o_acc[8 * u + 2] * inv, o_acc[8 * u + 3] * inv);
        o4.z = pack_bf16(o_acc[8 * u + 4] * inv, o_acc[8 * u + 5] * inv);
        o4.w = pack_bf16(o_acc[8 * u + 6] * inv, o_acc[8 * u + 7] * inv);
        reinterpret_cast<uint4*>(op)[u] = o4;
      }
      if (a.lse) a.lse[((long long)b * a.H + h) * a.Tq + t] = l_run > 0.f ? (m_run + log2f(l_run)) * LN2 : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int make_head_map(CUtensorMap* m, const void* base, long long bs, long long rs, int B, int T, int H, int rows) {
  unsigned long long dims[3] = {(unsigned long long)H * HD, (unsigned long long)T, (unsigned long long)B};
  unsigned long long strides[2] = {(unsigned long long)rs * 2, (unsigned long long)bs * 2};
  unsigned int box[3] = {(unsigned)HD, (unsigned)rows, 1u};
  return omr_make_tensor_map(m, 2, base, 3, dims, strides, box, nullptr, 128);
}

}  // namespace

int omr_attn_fwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                    const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse,
                    const float* key_bias, int B, int H, int Tq, int Tk, int hd, float scale, int causal, int window,
                    const int* q_len, const int* kv_len, int quirk_mod, cudaStream_t st) {
  if (hd != HD || (q_len == nullptr) != (kv_len == nullptr) || B < 1 || H < 1 || Tq < 1 || Tk < 1) return OMR_TC_NOT_ELIGIBLE;
  auto al = [](const void* p, long long bs, long long rs) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (bs * 2) % 16 == 0 && (rs * 2) % 16 == 0;
  };
  if (!al(q, q_bs, q_rs) || !al(k, k_bs, k_rs) || !al(v, v_bs, v_rs) || !al(o, o_bs, o_rs)) return OMR_TC_NOT_ELIGIBLE;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_head_map(&tmQ, q, q_bs, q_rs, B, Tq, H, BQ);
  if (rc) return rc;
  rc = make_head_map(&tmK, k, k_bs, k_rs, B, Tk, H, BKV);
  if (rc) return rc;
  rc = make_head_map(&tmV, v, v_bs, v_rs, B, Tk, H, BKV);
  if (rc) return rc;
  AttnTcArgs a{(bf16*)o, o_bs, o_rs, lse, key_bias, B, H, Tq, Tk, scale * LOG2E, causal, window, omr_attn_cur_dropout(),
               q_len, kv_len, quirk_mod};
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    configured = true;
  }
  dim3 grid((unsigned)((Tq + BQ - 1) / BQ), (unsigned)H, (unsigned)B);
  OmrLaunch(grid, 192, FWD_SMEM, st)(attn_fwd_tc_kernel, tmQ, tmK, tmV, a);
  OMR_LAUNCHED();
  return OMR_OK;
}

// =================================================================================================================
// Backward.  One CTA = one (batch, head, 128-key tile); K and V stay resident in shared memory and the CTA walks the
// query tiles that can see them.  Everything is computed TRANSPOSED (rows = keys) so that each softmax thread owns
// one key row and the probabilities land in shared memory directly in the operand layouts of the three gradient
// GEMMs:
//     S^T  = K Q^T                 M=128 keys, N=128 queries, K=64          (TMEM, recomputed)
//     dP^T = V dO^T                same shape
//     P^T  = exp2(scale*S^T + bias_k - lse_q)   dS^T = P^T o (dP^T - delta_q)   -- bf16 into two swizzled smem tiles
//     dV  += P^T  dO               A = P^T  (K-major),  B = dO tile (MN-major)   accumulated in TMEM over the q tiles
//     dK  += dS^T Q                A = dS^T (K-major),  B = Q  tile (MN-major)   accumulated in TMEM over the q tiles
//     dQ_t = dS K                  A = dS^T read as an MN-major operand, B = K tile (MN-major); per q tile, added to an
//                                  fp32 accumulation buffer with vector atomics (other key tiles add to the same rows)
// A second tiny kernel scales the fp32 dQ sums and writes them in the caller's (strided, bf16) layout.
// =================================================================================================================
namespace {

constexpr int DQ_LD = 68;  // floats per staged dQ row: 272 B, so that 16-byte stores of 8 lanes hit 8 different bank groups
constexpr int BWD_SMEM = TILE * 10 + 1024 + 256 + 128 * DQ_LD * 4;

__device__ __forceinline__ void q_tile_range(const AttnTcArgs& a, int j0, int& qt0, int& qt1) {
  const int nqt = (a.Tq + BQ - 1) / BQ;
  qt0 = 0; qt1 = nqt;
  if (a.causal) {
    const int off = a.Tk - a.Tq;
    int j_last = j0 + BKV - 1;
    if (j_last > a.Tk - 1) j_last = a.Tk - 1;
    const int tmin = j0 - off;  // t >= j - off
    if (tmin > 0) qt0 = tmin / BQ;
    if (a.window > 0) {
      const int tmax = j_last - off + a.window;  // t <= j - off + window
      if (tmax < 0) { qt1 = 0; return; }
      const int e = tmax / BQ + 1;
      if (e < qt1) qt1 = e;
    }
    if (qt0 > qt1) qt0 = qt1;
  }
}

struct AttnBwdArgs {
  AttnTcArgs f;        // o/lse fields: lse is the saved forward statistic
  const float* delta;  // [B,H,Tq]
  float* dq_acc;       // [B,H,Tq,64] fp32, zeroed
  bf16* dk; long long dk_bs, dk_rs;
  bf16* dv; long long dv_bs, dv_rs;
  float scale;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(320, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                             const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV,
                                                             const __grid_constant__ CUtensorMap tmDO, AttnBwdArgs g) {
  omr_pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  const AttnTcArgs& a = g.f;
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE;
  uint8_t* sQ = smem + 2 * TILE;   // [2]
  uint8_t* sDO = smem + 4 * TILE;  // [2]
  uint8_t* sPT = smem + 6 * TILE;  // 2 chunks of 64 queries
  uint8_t* sDS = smem + 8 * TILE;  // 2 chunks of 64 queries
  float* sLse = reinterpret_cast<float*>(smem + 10 * TILE);  // [128] (already * log2e)
  float* sDelta = sLse + 128;                               // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 10 * TILE + 1024);
  float* sDQ = reinterpret_cast<float*>(smem + 10 * TILE + 1024 + 256);  // [128][DQ_LD] staging of a dQ tile
  uint64_t* kv_full = bars;
  uint64_t* qd_full = bars + 1;   // [2]
  uint64_t* qd_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;  // S^T and dP^T ready
  uint64_t* pds_full = bars + 6;  // P^T and dS^T written (count 4)
  uint64_t* mma2_done = bars + 7; // dV, dK, dQ MMAs of this q tile complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
  int qt0, qt1;
  q_tile_range(a, j0, qt0, qt1);
  // a key tile whose every key is masked out (bias = -inf: the padded tail of the memory) has P = 0 throughout:
  // dK = dV = 0 and no contribution to dQ -- the CTA only writes the zeros
  bool row_dead = true;
  if (warp < 8) {
    const int jr = j0 + (warp & 3) * 32 + lane;
    row_dead = jr >= a.Tk || (a.key_bias && !(a.key_bias[(long long)b * a.Tk + jr] > -INFINITY));
  }
  const int ntiles = __syncthreads_and(row_dead) ? 0 : qt1 - qt0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(mma2_done, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t tmem_ST = tmem_base, tmem_DPT = tmem_base + 128, tmem_DV = tmem_base + 256, tmem_DK = tmem_base + 320,
                 tmem_DQ = tmem_base + 384;

  // warp-uniform producer / issuer loops, instructions under elect_one() (see tc_common.cuh)
  if (warp == 8) {
    if (ntiles > 0) {
      if (elect_one()) {
        mbar_expect_tx(kv_full, 2 * TILE);
        tma_load_3d(sK, &tmK, kv_full, h * HD, j0, b);
        tma_load_3d(sV, &tmV, kv_full, h * HD, j0, b);
      }
      for (int i = 0; i < ntiles; ++i) {
        const int s = i & 1;
        mbar_wait(&qd_empty[s], ((i >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&qd_full[s], 2 * TILE);
          tma_load_3d(sQ + s * TILE, &tmQ, &qd_full[s], h * HD, (qt0 + i) * BQ, b);
          tma_load_3d(sDO + s * TILE, &tmDO, &qd_full[s], h * HD, (qt0 + i) * BQ, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    if (ntiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);  // A K-major (P^T / dS^T), B MN-major
      constexpr uint32_t idesc_mn = make_idesc_bf16(128, 64, 1, 1);  // A MN-major (dS), B MN-major
      mbar_wait(kv_full, 0);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), pt_addr = smem_u32(sPT), ds_addr = smem_u32(sDS);
      const uint32_t q_base = smem_u32(sQ), do_base = smem_u32(sDO);
      for (int i = 0; i < ntiles; ++i) {
        const int s = i & 1;
        mbar_wait(&qd_full[s], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t q_addr = q_base + (uint32_t)s * TILE, do_addr = do_base + (uint32_t)s * TILE;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j > 0)
              umma_bf16_acc(tmem_ST, make_smem_desc(k_addr + j * 32, 16, 1024, 128), make_smem_desc(q_addr + j * 32, 16, 1024, 128), idesc_s);
            else
              umma_bf16_new(tmem_ST, make_smem_desc(k_addr, 16, 1024, 128), make_smem_desc(q_addr, 16, 1024, 128), idesc_s);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j > 0)
              umma_bf16_acc(tmem_DPT, make_smem_desc(v_addr + j * 32, 16, 1024, 128), make_smem_desc(do_addr + j * 32, 16, 1024, 128), idesc_s);
            else
              umma_bf16_new(tmem_DPT, make_smem_desc(v_addr, 16, 1024, 128), make_smem_desc(do_addr, 16, 1024, 128), idesc_s);
          }
          umma_commit(sdp_full);
        }
        __syncwarp();
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // dV += P^T dO
            if (j > 0)
              umma_bf16_acc(tmem_DV, make_smem_desc(pt_addr + (j >> 2) * TILE + (j & 3) * 32, 16, 1024, 128),
                            make_smem_desc(do_addr + j * 2048, 0, 1024, 128), idesc_kn);
            else
              umma_bf16(tmem_DV, make_smem_desc(pt_addr, 16, 1024, 128), make_smem_desc(do_addr, 0, 1024, 128), idesc_kn, i > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // dK += dS^T Q
            if (j > 0)
              umma_bf16_acc(tmem_DK, make_smem_desc(ds_addr + (j >> 2) * TILE + (j & 3) * 32, 16, 1024, 128),
                            make_smem_desc(q_addr + j * 2048, 0, 1024, 128), idesc_kn);
            else
              umma_bf16(tmem_DK, make_smem_desc(ds_addr, 16, 1024, 128), make_smem_desc(q_addr, 0, 1024, 128), idesc_kn, i > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // dQ_tile = dS K : A = dS^T tile read MN-major (M = queries: 2 chunks, K = key rows)
            if (j > 0)
              umma_bf16_acc(tmem_DQ, make_smem_desc(ds_addr + j * 2048, TILE, 1024, 128), make_smem_desc(k_addr + j * 2048, 0, 1024, 128), idesc_mn);
            else
              umma_bf16_new(tmem_DQ, make_smem_desc(ds_addr, TILE, 1024, 128), make_smem_desc(k_addr, 0, 1024, 128), idesc_mn);
          }
          umma_commit(mma2_done);
          umma_commit(&qd_empty[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- 8 warps: thread = (key row r, column half hf) of the S^T / dP^T / dV / dK accumulators, and (query row r,
    // column half) of the dQ accumulator.  Warps w and w+4 share TMEM lanes 32(w&3)..+31 and split the columns, so the
    // exp/convert/store work of a tile is spread over twice the issue slots of a 4-warp epilogue. ----
    const int r = (warp & 3) * 32 + lane;
    const int hf = warp >> 2;
    const int j = j0 + r;
    const int off = a.Tk - a.Tq;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const float bias = j < a.Tk ? (a.key_bias ? a.key_bias[(long long)b * a.Tk + j] * LOG2E : 0.f) : -INFINITY;
    // queries that can see key j: t in [t_lo, t_hi]
    int t_lo = 0, t_hi = a.Tq - 1;
    if (a.causal) {
      t_lo = max(0, j - off);
      if (a.window > 0) t_hi = min(t_hi, j - off + a.window);
    }
    if (j >= a.Tk) t_hi = -1;
    int lq, lkv;
    block_mask_of(a, b, h, lq, lkv);
    if (j >= lkv) t_hi = min(t_hi, lq - 1);
    const long long stat_base = ((long long)b * a.H + h) * a.Tq;
    const uint32_t dstream = a.drop.thr ? attn_drop_stream(a.drop, b * a.H + h) : 0u;
    const uint32_t dkp = (uint32_t)((a.Tk + 1) >> 1), dsh = (uint32_t)(j & 1) * 16u;

    // dQ tile -> fp32 accumulation buffer.  The accumulator row of a thread is 128 contiguous bytes, but a warp-wide
    // red of one register group would touch 32 different lines; the tile is therefore transposed through shared memory
    // (padded rows: conflict-free both ways) and added with reds that cover 512 contiguous bytes per warp instruction.
    auto dq_epilogue = [&](int q_tile) {
      softmax_bar2();  // the previous tile's readers are done with sDQ
      {
        uint32_t v[32];
        tmem_ld32(tmem_DQ + lane_addr + hf * 32, v);
        tmem_ld_wait();
        float* rowp = sDQ + r * DQ_LD + hf * 32;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<float4*>(rowp + 4 * u) = make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]),
                                                                  __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
      }
      softmax_bar2();
      const int tid8 = threadIdx.x;  // 0..255: the eight epilogue warps
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int L = it * 256 + tid8, row = L >> 4, c4 = L & 15;
        const int t = q_tile * BQ + row;
        const float4 x = *reinterpret_cast<const float4*>(sDQ + row * DQ_LD + c4 * 4);
        if (t < a.Tq) red_add_v4(g.dq_acc + (stat_base + t) * HD + c4 * 4, x.x, x.y, x.z, x.w);
      }
    };

    for (int i = 0; i < ntiles; ++i) {
      const int q0 = (qt0 + i) * BQ;
      softmax_bar2();  // readers of the previous tile's statistics are done
      if (hf == 0) {
        const int t = q0 + r;
        sLse[r] = t < a.Tq ? a.lse[stat_base + t] * LOG2E : 0.f;
        sDelta[r] = t < a.Tq ? g.delta[stat_base + t] : 0.f;
      }
      softmax_bar2();
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      if (i > 0) {
        mbar_wait(mma2_done, (i - 1) & 1);  // P^T / dS^T tiles are free again, dQ of the previous q tile is complete
        tc_fence_after();
        dq_epilogue(qt0 + i - 1);
      }
      // every (query, key) pair of this tile visible?  (then no interval tests; rows past Tk carry bias = -inf)
      bool full = q0 + BQ <= a.Tq && j0 + BKV <= a.Tk;
      if (a.causal) full = full && q0 >= j0 + BKV - 1 - off && (a.window <= 0 || q0 + BQ - 1 <= j0 - off + a.window);
      full = full && (j0 + BKV <= lkv || q0 + BQ <= lq);
      auto tile_half = [&](auto masked_tag, auto drop_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value, DROP = decltype(drop_tag)::value;
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
          uint32_t sv[32], dv[32];
          tmem_ld32(tmem_ST + lane_addr + c * 32, sv);
          tmem_ld32(tmem_DPT + lane_addr + c * 32, dv);
          tmem_ld_wait();
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(sLse + c * 32 + e);
            const float4 d4 = *reinterpret_cast<const float4*>(sDelta + c * 32 + e);
            const float ls[4] = {bias - l4.x, bias - l4.y, bias - l4.z, bias - l4.w};
            const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
            float pr[4], fk[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int t = q0 + c * 32 + e + k;
              pr[k] = ex2_approx(fmaf(__uint_as_float(sv[e + k]), a.scale_log2, ls[k]));
              if (MASKED) pr[k] = (t >= t_lo && t <= t_hi) ? pr[k] : 0.f;
            }
            if (DROP) {  // mask / (1-p) of the pairs (t, j).  Keys j and j^1 (neighbouring lanes) share their 2 x 2 blocks:
              // the even lane hashes the block of queries (t, t+1), the odd lane that of (t+2, t+3); both words travel
              const int t = q0 + c * 32 + e;
              const uint2 mine = attn_drop_block(dstream, (uint32_t)((t >> 1) + (j & 1)), (uint32_t)(j >> 1), dkp);
              const uint32_t ox = __shfl_xor_sync(0xffffffffu, mine.x, 1), oy = __shfl_xor_sync(0xffffffffu, mine.y, 1);
              const uint2 b0 = (j & 1) ? make_uint2(ox, oy) : mine;  // queries (t, t+1)
              const uint2 b1 = (j & 1) ? mine : make_uint2(ox, oy);  // queries (t+2, t+3)
              fk[0] = ((b0.x >> dsh) & 0xFFFFu) >= a.drop.thr ? a.drop.inv_keep : 0.f;
              fk[1] = ((b0.y >> dsh) & 0xFFFFu) >= a.drop.thr ? a.drop.inv_keep : 0.f;
              fk[2] = ((b1.x >> dsh) & 0xFFFFu) >= a.drop.thr ? a.drop.inv_keep : 0.f;
              fk[3] = ((b1.y >> dsh) & 0xFFFFu) >= a.drop.thr ? a.drop.inv_keep : 0.f;
            }
            float ds[4], pd[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ds[k] = pr[k] * (DROP ? fmaf(__uint_as_float(dv[e + k]), fk[k], -dl[k]) : __uint_as_float(dv[e + k]) - dl[k]);
              pd[k] = DROP ? pr[k] * fk[k] : pr[k];
            }
            pk[e >> 1] = pack_bf16(pd[0], pd[1]);
            pk[(e >> 1) + 1] = pack_bf16(pd[2], pd[3]);
            dk[e >> 1] = pack_bf16(ds[0], ds[1]);
            dk[(e >> 1) + 1] = pack_bf16(ds[2], ds[3]);
          }
          uint8_t* prow = sPT + (c >> 1) * TILE + r * 128;
          uint8_t* drow = sDS + (c >> 1) * TILE + r * 128;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int sw = (((c & 1) * 4 + u) ^ (r & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + sw) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            *reinterpret_cast<uint4*>(drow + sw) = make_uint4(dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
          }
        }
      };
      if (a.drop.thr) {
        if (full) tile_half(std::false_type{}, std::true_type{});
        else tile_half(std::true_type{}, std::true_type{});
      } else {
        if (full) tile_half(std::false_type{}, std::false_type{});
        else tile_half(std::true_type{}, std::false_type{});
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    if (ntiles > 0) {
      mbar_wait(mma2_done, (ntiles - 1) & 1);
      tc_fence_after();
      dq_epilogue(qt0 + ntiles - 1);
    }
    // dK (x scale) and dV rows of this key
    if (j < a.Tk) {
      bf16* dkp = g.dk + (long long)b * g.dk_bs + (long long)j * g.dk_rs + h * HD;
      bf16* dvp = g.dv + (long long)b * g.dv_bs + (long long)j * g.dv_rs + h * HD;
      if (ntiles == 0) {
#pragma unroll
        for (int u = 4 * hf; u < 4 * hf + 4; ++u) {
          reinterpret_cast<uint4*>(dkp)[u] = make_uint4(0, 0, 0, 0);
          reinterpret_cast<uint4*>(dvp)[u] = make_uint4(0, 0, 0, 0);
        }
      }
    }
    if (ntiles > 0) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const float mul = which == 0 ? g.scale : 1.f;
        bf16* dst = which == 0 ? g.dk + (long long)b * g.dk_bs + (long long)j * g.dk_rs + h * HD
                               : g.dv + (long long)b * g.dv_bs + (long long)j * g.dv_rs + h * HD;
        {
          const int c = hf;
          uint32_t v[32];
          tmem_ld32((which == 0 ? tmem_DK : tmem_DV) + lane_addr + c * 32, v);
          tmem_ld_wait();
          if (j < a.Tk) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 o4;
              o4.x = pack_bf16(__uint_as_float(v[8 * u]) * mul, __uint_as_float(v[8 * u + 1]) * mul);
              o4.y = pack_bf16(__uint_as_float(v[8 * u + 2]) * mul, __uint_as_float(v[8 * u + 3]) * mul);
              o4.z = pack_bf16(__uint_as_float(v[8 * u + 4]) * mul, __uint_as_float(v[8 * u + 5]) * mul);
              o4.w = pack_bf16(__uint_as_float(v[8 * u + 6]) * mul, __uint_as_float(v[8 * u + 7]) * mul);
              reinterpret_cast<uint4*>(dst)[c * 4 + u] = o4;
            }
          }
        }
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

// delta[b,h,t] = sum_d dO * O ; one warp per (b,h,t)
__global__ void attn_delta_tc_kernel(const bf16* __restrict__ o, long long o_bs, long long o_rs, const bf16* __restrict__ dO,
                                     long long do_bs, long long do_rs, float* __restrict__ delta, int B, int H, int Tq) {
  omr_pdl_enter();
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * H * Tq) return;
  const int t = (int)(w % Tq);
  const long long rr = w / Tq;
  const int h = (int)(rr % H), b = (int)(rr / H);
  const __nv_bfloat162 ov = *reinterpret_cast<const __nv_bfloat162*>(o + (long long)b * o_bs + (long long)t * o_rs + h * HD + 2 * lane);
  const __nv_bfloat162 dv = *reinterpret_cast<const __nv_bfloat162*>(dO + (long long)b * do_bs + (long long)t * do_rs + h * HD + 2 * lane);
  float s = __low2float(ov) * __low2float(dv) + __high2float(ov) * __high2float(dv);
  s = warp_sum(s);
  if (lane == 0) delta[w] = s;
}

// dq[b,t,h,:] = bf16(scale * dq_acc[b,h,t,:])
__global__ void attn_dq_finalize_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long dq_bs, long long dq_rs,
                                        int B, int H, int Tq, float scale) {
  omr_pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread = 8 channels
  if (i >= (long long)B * H * Tq * 8) return;
  const int u = (int)(i & 7);
  const long long row = i >> 3;
  const int t = (int)(row % Tq);
  const long long rr = row / Tq;
  const int h = (int)(rr % H), b = (int)(rr / H);
  const float4 x = *reinterpret_cast<const float4*>(acc + row * HD + u * 8);
  const float4 y = *reinterpret_cast<const float4*>(acc + row * HD + u * 8 + 4);
  uint4 o4;
  o4.x = pack_bf16(x.x * scale, x.y * scale); o4.y = pack_bf16(x.z * scale, x.w * scale);
  o4.z = pack_bf16(y.x * scale, y.y * scale); o4.w = pack_bf16(y.z * scale, y.w * scale);
  *reinterpret_cast<uint4*>(dq + (long long)b * dq_bs + (long long)t * dq_rs + h * HD + u * 8) = o4;
}

}  // namespace

// ws: fp32 scratch of B*H*Tq*(1 + 64) floats: delta [B,H,Tq] followed by the dQ accumulators [B,H,Tq,64]
int omr_attn_bwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                    const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                    const void* dout, long long do_bs, long long do_rs, const float* lse, void* dq, long long dq_bs,
                    long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv, long long dv_bs, long long dv_rs,
                    float* ws, const float* key_bias, int B, int H, int Tq, int Tk, int hd, float scale, int causal,
                    int window, const int* q_len, const int* kv_len, int quirk_mod, cudaStream_t st) {
  if (hd != HD || (q_len == nullptr) != (kv_len == nullptr) || B < 1 || H < 1 || Tq < 1 || Tk < 1) return OMR_TC_NOT_ELIGIBLE;
  auto al = [](const void* p, long long bs, long long rs) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (bs * 2) % 16 == 0 && (rs * 2) % 16 == 0;
  };
  if (!al(q, q_bs, q_rs) || !al(k, k_bs, k_rs) || !al(v, v_bs, v_rs) || !al(o, o_bs, o_rs) || !al(dout, do_bs, do_rs) ||
      !al(dq, dq_bs, dq_rs) || !al(dk, dk_bs, dk_rs) || !al(dv, dv_bs, dv_rs))
    return OMR_TC_NOT_ELIGIBLE;
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc = make_head_map(&tmQ, q, q_bs, q_rs, B, Tq, H, BQ);
  if (rc) return rc;
  rc = make_head_map(&tmK, k, k_bs, k_rs, B, Tk, H, BKV);
  if (rc) return rc;
  rc = make_head_map(&tmV, v, v_bs, v_rs, B, Tk, H, BKV);
  if (rc) return rc;
  rc = make_head_map(&tmDO, dout, do_bs, do_rs, B, Tq, H, BQ);
  if (rc) return rc;
  const long long rows = (long long)B * H * Tq;
  float* delta = ws;
  float* dq_acc = ws + ((rows + 3) / 4) * 4;  // keep the accumulators 16-byte aligned
  OMR_CUDA(cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)rows * HD, st));
  OmrLaunch((unsigned)((rows * 32 + 255) / 256), 256, 0, st)(attn_delta_tc_kernel, (const bf16*)o, o_bs, o_rs, (const bf16*)dout, do_bs, do_rs,
                                                                            delta, B, H, Tq);
  OMR_LAUNCHED();
  AttnBwdArgs g{};
  g.f = AttnTcArgs{nullptr, 0, 0, const_cast<float*>(lse), key_bias, B, H, Tq, Tk, scale * LOG2E, causal, window,
                   omr_attn_cur_dropout(), q_len, kv_len, quirk_mod};
  g.delta = delta; g.dq_acc = dq_acc;
  g.dk = (bf16*)dk; g.dk_bs = dk_bs; g.dk_rs = dk_rs;
  g.dv = (bf16*)dv; g.dv_bs = dv_bs; g.dv_rs = dv_rs;
  g.scale = scale;
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    configured = true;
  }
  dim3 grid((unsigned)((Tk + BKV - 1) / BKV), (unsigned)H, (unsigned)B);
  OmrLaunch(grid, 320, BWD_SMEM, st)(attn_bwd_tc_kernel, tmQ, tmK, tmV, tmDO, g);
  OMR_LAUNCHED();
  OmrLaunch((unsigned)((rows * 8 + 255) / 256), 256, 0, st)(attn_dq_finalize_kernel, dq_acc, (bf16*)dq, dq_bs, dq_rs, B, H, Tq, scale);
  OMR_LAUNCHED();
  return OMR_OK;
}
