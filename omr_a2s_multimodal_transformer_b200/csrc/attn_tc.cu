// attn_tc.cu -- flash-style attention (head_dim 64, bf16) on the tcgen05 tensor cores: forward, backward.
//
// FORWARD.  One CTA = one (batch, head, 128-query tile); it walks the visible 128-key tiles:
//     S  = Q K^T          tcgen05.mma  M=128 (queries) N=128 (keys) K=64,  fp32 S in TMEM
//     P  = online softmax of  scale*S + key_bias (+ causal / sliding-window structure)   -- 8 softmax warps: the two warps
//                         of a TMEM lane quarter split the tile's keys (64 each), exp2 against a lazily updated reference
//                         maximum, bf16 P written to shared memory in the 128B-swizzled K-major operand layout
//     O += P V            tcgen05.mma  M=128 N=64 K=128 (V rows are the reduction index: MN-major B), accumulated in TMEM
//                         across the key tiles (rescaled there only when a row's maximum runs away)
// Q/K/V tiles arrive by TMA straight out of the packed projection buffers ([B,T,3D] / [B,S,2D]; the head is a
// column offset), K/V double buffered.  Two CTAs fit on an SM (113 KB smem, 256 TMEM columns each), so one CTA's
// softmax overlaps the other's MMAs.  The mask algebra is the reference's: additive fp32
// key bias per (b, key) (0, +1.0 or -inf), keys visible to query t iff k <= t + Tk - Tq (causal) and
// k >= t + Tk - Tq - window.  lse (natural log) is saved for the backward.
// Warp roles: 0-7 softmax + epilogue (TMEM lanes 32(w&3)..+31, key / channel half w>>2), 8 TMA producer, 9 MMA issuer +
// TMEM allocator.  Measured per key tile and CTA (OMR_ATTN_DEBUG=512, scripts/attn_stamps_fwd.py, C3 cross shape): 3500 clk =
// 250 (S MMA + hop) + 640 (row maximum + exchange) + 1800 (exp pass; MUFU.EX2 at 16 / clk / SM is ~60 % busy with two CTAs per
// SM) + 520 (P V issue) + 350 (next S issue).
// BACKWARD: see the comment in front of the backward kernel below.
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
  unsigned long long* stamps;  // forward, OMR_ATTN_DEBUG & 512: clock stamps of CTA (0,0,0) ([role][64]); NULL = off
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

// Q, K[3], V[3] + bias tile + barriers: 115,456 B, so that two CTAs (+1 KB system reserve each) fit in the 228 KB of an
// SM (256 bytes to spare); the dynamic window is declared 1024-aligned (no static shared memory in this kernel)
constexpr int FWD_KVST = 3;  // K/V stages (a tile lands ~2500 clk after its load is issued)
constexpr int FWD_OFF_BIAS = TILE * (1 + 2 * FWD_KVST), FWD_OFF_BAR = FWD_OFF_BIAS + 512, FWD_SMEM = FWD_OFF_BAR + 256;
static_assert(2 * (FWD_SMEM + 1024) <= 228 * 1024, "forward attention: two CTAs per SM");

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

__device__ __forceinline__ void softmax_bar2() { asm volatile("bar.sync 1, 256;" ::: "memory"); }


__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- the per-tile work of a softmax thread (one query row, 64 of the tile's 128 keys of S in TMEM), specialised so that
// the common tile -- every key visible to every row -- pays for no interval tests, and a tile whose key bias is zero
// throughout for no bias loads.  x = scale_log2 * S + bias (log2 units).  MASKED tiles (causal diagonal, window edge,
// ragged key tail) test the row's visible interval [k_lo, k_hi] per key.
struct SmRow {
  uint32_t tmem_s;      // TMEM address of this thread's 64 S values
  const float* bias;    // shared-memory bias values of its 64 keys (pre-multiplied by log2 e)
  float scale_log2;
  int j0, k_lo, k_hi;   // first key of the thread's 64, visible interval of the row
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <bool MASKED, bool BIAS>
__device__ __forceinline__ float softmax_row_max(const SmRow& w) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(w.tmem_s + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      float x[4];
      float bq[4] = {0.f, 0.f, 0.f, 0.f};
      if (BIAS || MASKED) {
        const float4 b4 = *reinterpret_cast<const float4*>(w.bias + c * 32 + e);
        bq[0] = b4.x; bq[1] = b4.y; bq[2] = b4.z; bq[3] = b4.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = w.j0 + c * 32 + e + k;
        x[k] = __uint_as_float(v[e + k]);
        if (BIAS || MASKED) x[k] = fmaf(x[k], w.scale_log2, bq[k]);  // otherwise the (positive) scale is applied to the max
        if (MASKED) x[k] = (j >= w.k_lo && j <= w.k_hi) ? x[k] : -INFINITY;
      }
      mx = fmax3(mx, x[0], x[1]);
      mx = fmax3(mx, x[2], x[3]);
    }
  }
  return (BIAS || MASKED) ? mx : mx * w.scale_log2;
}

// P = exp2(x - m) -> row sum over the thread's 64 keys (returned) and the packed bf16 values back into TENSOR memory (key k of
// the tile in column k / 2 of the P region: the A operand of the P V MMA is read from TMEM -- round 2, second half: the smem
// round trip of P was 45 % of the kernel's shared-memory traffic); with DROP the stored probabilities carry the keep mask
// (1/(1-p) is applied to the output row at the end) while the sum stays that of the full row
template <bool MASKED, bool BIAS, bool DROP>
__device__ __forceinline__ float softmax_row_exp(const SmRow& w, float m_safe, uint32_t tmem_p, int t, uint32_t dstream,
                                                 uint32_t dkp, uint32_t thr2) {
  float rs = 0.f;
  const float nm = -m_safe;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
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
      uint32_t m0 = 0xFFFFFFFFu, m1 = 0xFFFFFFFFu;
      if (DROP) {
        // rows t and t^1 (neighbouring lanes) share their 2 x 2 blocks: the even lane hashes the block of keys
        // (j, j+1), the odd lane that of (j+2, j+3), and each hands the other the word of its row parity.  Both fields of
        // a word are compared at once and the result becomes an AND mask of the packed bf16 pair.
        const int j = w.j0 + c * 32 + e;
        const uint2 mine = attn_drop_block(dstream, (uint32_t)(t >> 1), (uint32_t)((j >> 1) + (t & 1)), dkp);
        const uint32_t ox = __shfl_xor_sync(0xffffffffu, mine.x, 1), oy = __shfl_xor_sync(0xffffffffu, mine.y, 1);
        const uint32_t w0 = (t & 1) ? oy : mine.x;  // keys (j, j+1) for this row
        const uint32_t w1 = (t & 1) ? mine.y : ox;  // keys (j+2, j+3)
        m0 = attn_drop_mask_bf16x2(attn_drop_flags(w0, thr2));
        m1 = attn_drop_mask_bf16x2(attn_drop_flags(w1, thr2));
      }
      pk[e >> 1] = pack_bf16(pr[0], pr[1]) & m0;
      pk[(e >> 1) + 1] = pack_bf16(pr[2], pr[3]) & m1;
    }
    tmem_st16(tmem_p + c * 16, pk);  // keys c*32 .. c*32+31 of the thread's 64 -> 16 packed columns
  }
  tmem_st_wait();
  return rs;
}

constexpr int FWD_SW = 8;                      // softmax warps: warp w = TMEM lane quarter w & 3, key / channel half w >> 2
constexpr float FWD_RESCALE_THRESHOLD = 8.f;  // log2 units: the running reference maximum moves only when a row's true
                                              // maximum exceeds it by more than this (P <= 2^8, harmless in bf16 / fp32)

// OR over the 256 softmax threads (also a barrier among them)
__device__ __forceinline__ bool softmax_bar_or(bool mine) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "barrier.cta.red.or.pred p, 1, 256, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"((uint32_t)mine)
      : "memory");
  return out != 0;
}
// the two warps that share TMEM lane quarter lq (64 threads)
__device__ __forceinline__ void pair_bar(int lq) { asm volatile("bar.sync %0, 64;" ::"r"(2 + lq) : "memory"); }
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
// a value per row travels between the two warps of a lane quarter through a TMEM column of the row's own lane (shared
// memory is full: two CTAs per SM leave 256 bytes to spare).  The second barrier keeps the partner's read ahead of the next
// writer of that column (the exp pass, which stores P there).
__device__ __forceinline__ float pair_exchange(uint32_t col_mine, uint32_t col_other, int lq, float v) {
  tmem_st1(col_mine, __float_as_uint(v));
  tmem_st_wait();
  tc_fence_before();
  pair_bar(lq);
  tc_fence_after();
  const uint32_t o = tmem_ld1(col_other);
  tmem_ld_wait();
  tc_fence_before();
  pair_bar(lq);
  tc_fence_after();
  return __uint_as_float(o);
}

// Forward, round 2: EIGHT softmax warps per CTA (two CTAs per SM -> four warps per scheduler; round 1 had two, and the
// kernel sat at 32 % issue utilisation waiting on tcgen05.ld / MUFU latencies).  The two warps of a lane quarter split
// the tile's keys (64 each) and the output channels (32 each).  O stays in TMEM and accumulates across key tiles; the
// reference maximum of a row moves only when the row's true maximum has run away by more than 2^8 (then the row's O
// accumulator and running sum are rescaled through tcgen05.ld / st), so the common tile neither reads nor rescales O.
__global__ void __launch_bounds__(32 * (FWD_SW + 2), 2) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                          const __grid_constant__ CUtensorMap tmK,
                                                                          const __grid_constant__ CUtensorMap tmV,
                                                                          AttnTcArgs a) {
  omr_pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();  // the swizzled tiles need 1 KB alignment
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE;                   // [FWD_KVST]
  uint8_t* sV = smem + (1 + FWD_KVST) * TILE;  // [FWD_KVST]
  float* sBias = reinterpret_cast<float*>(smem + FWD_OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;              // [FWD_KVST]
  uint64_t* kv_empty = bars + 1 + FWD_KVST;  // [FWD_KVST]
  uint64_t* s_full = bars + 1 + 2 * FWD_KVST;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_full = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);
  uint32_t* live = reinterpret_cast<uint32_t*>(s_full + 4);  // [32] bit i: key tile kt0 + i holds a key that is not masked out

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

  if (warp == FWD_SW && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < FWD_KVST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, FWD_SW);
    mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == FWD_SW + 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t tmem_S = tmem_base, tmem_PV = tmem_base + 128, tmem_P = tmem_base + 192;  // P: 64 columns of packed bf16

  // Producer and MMA issuer run their loops WARP-wide on uniform values and issue under elect_one() (tc_common.cuh): a
  // lone lane under `if (lane == 0)` made the compiler wrap every tcgen05.mma / TMA in an ELECT / R2UR / BRA.U.ANY waterfall.
  if (warp == FWD_SW) {
    if (ntiles > 0) {
      if (elect_one()) {
        mbar_expect_tx(q_full, TILE);
        tma_load_3d(sQ, &tmQ, q_full, h * HD, q0, b);
      }
      for (int i = 0, n = 0; i < ntiles; ++i) {
        if (!bcast0(tile_live(i) ? 1u : 0u)) continue;
        const int s = n % FWD_KVST;
        mbar_wait(&kv_empty[s], ((n / FWD_KVST) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[s], 2 * TILE);
          tma_load_3d(sK + s * TILE, &tmK, &kv_full[s], h * HD, (kt0 + i) * BKV, b);
          tma_load_3d(sV + s * TILE, &tmV, &kv_full[s], h * HD, (kt0 + i) * BKV, b);
        }
        __syncwarp();
        if (a.stamps && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 && n < 64) a.stamps[0 * 64 + n] = clock64();
        ++n;
      }
    }
  } else if (warp == FWD_SW + 1) {
    if (ntiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      mbar_wait(q_full, 0);
      const uint32_t q_addr = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV);
      for (int i = 0, n = 0; i < ntiles; ++i) {
        if (!bcast0(tile_live(i) ? 1u : 0u)) continue;
        const int s = n % FWD_KVST;
        mbar_wait(&kv_full[s], (n / FWD_KVST) & 1);
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
        if (a.stamps && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 && n < 64) a.stamps[1 * 64 + n] = clock64();
        mbar_wait(p_full, n & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 8; ++j)  // O += P V: A = P from TMEM (keys 16j .. 16j+15 in columns 8j .. 8j+7); the softmax
                                       // warps have rescaled O if the reference maximum moved
            umma_bf16_ta(tmem_PV, tmem_P + j * 8, make_smem_desc(v_addr + j * 2048, 0, 1024, 128), idesc_pv, (n > 0 || j > 0) ? 1u : 0u);
          umma_commit(pv_full);
          umma_commit(&kv_empty[s]);
        }
        __syncwarp();
        if (a.stamps && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 && n < 64) a.stamps[2 * 64 + n] = clock64();
        ++n;
      }
    }
  } else {
    // ---- softmax + epilogue: thread = (query row, key half / channel half) ----
    const int lq = warp & 3, hf = warp >> 2;
    const int r = lq * 32 + lane;
    const int t = q0 + r;
    const int off = a.Tk - a.Tq;
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    // the pair exchange uses two columns of the P region, which is idle whenever an exchange happens (the P V MMA that read
    // it has completed: s_full / pv_full have been waited for) and is rewritten by the exp pass afterwards
    const uint32_t x_mine = tmem_P + lane_addr + (uint32_t)hf, x_other = tmem_P + lane_addr + (uint32_t)(hf ^ 1);
    float m_used = -INFINITY, l_run = 0.f;
    const float* kb = a.key_bias ? a.key_bias + (long long)b * a.Tk : nullptr;
    const uint32_t dstream = a.drop.thr ? attn_drop_stream(a.drop, b * a.H + h) : 0u;
    const uint32_t dkp = (uint32_t)((a.Tk + 1) >> 1), thr2 = a.drop.thr * 0x00010001u;
    // visible key interval of this row
    int k_hi = a.Tk - 1, k_lo = 0;
    if (a.causal) {
      k_hi = min(k_hi, t + off);
      if (a.window > 0) k_lo = max(0, t + off - a.window);
    }
    if (t >= a.Tq) k_hi = -1;
    int lqm, lkv;
    block_mask_of(a, b, h, lqm, lkv);
    if (t >= lqm) k_hi = min(k_hi, lkv - 1);
    const uint32_t tmem_p = tmem_P + lane_addr + (uint32_t)hf * 32;  // this thread's 64 keys = 32 packed columns
    int n = 0;  // live tiles processed so far (the barrier phases count these)
    for (int i = 0; i < ntiles; ++i) {
      if (!tile_live(i)) continue;
      const int j0 = (kt0 + i) * BKV;
      // key-bias tile (pre-multiplied by log2 e) -> smem, shared by the 128 rows; is any of it non-zero?
      softmax_bar2();  // previous tile's readers are done with sBias
      bool nz = false;
      if (hf == 0) {
        const int j = j0 + r;
        const float bv = (kb && j < a.Tk) ? kb[j] * LOG2E : 0.f;
        sBias[r] = bv;
        nz = bv != 0.f;
      }
      const bool any_bias = softmax_bar_or(nz);
      mbar_wait(s_full, n & 1);
      tc_fence_after();
      const bool stamper = a.stamps && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && threadIdx.x == 0 && n < 64;
      if (stamper) a.stamps[3 * 64 + n] = clock64();
      // does every row of this CTA see every key of the tile?  (rows past Tq are never stored: they may see anything)
      bool full = j0 + BKV <= a.Tk;
      if (a.causal) {
        full = full && (j0 + BKV - 1 <= q0 + off) && (a.window <= 0 || j0 >= q0 + BQ - 1 + off - a.window);
      }
      full = full && (q0 + BQ <= lqm || j0 + BKV <= lkv);  // block mask: no masked row in this CTA, or the tile lies below the cut
      const SmRow w{tmem_S + lane_addr + hf * 64, sBias + hf * 64, a.scale_log2, j0 + hf * 64, k_lo, k_hi};
      // pass 1: maximum of the thread's 64 keys, then of the row (exchange with the other half's warp)
      float mx;
      if (!full) mx = softmax_row_max<true, true>(w);
      else if (any_bias) mx = softmax_row_max<false, true>(w);
      else mx = softmax_row_max<false, false>(w);
      mx = fmaxf(mx, pair_exchange(x_mine, x_other, lq, mx));
      const bool need = mx > m_used + FWD_RESCALE_THRESHOLD;  // also the first finite maximum of a row (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2_approx(m_used - mx) : 1.f;  // m_used = -inf -> 0 (nothing accumulated yet)
        if (n > 0) {  // O accumulator rows of this warp's 32 channels
          mbar_wait(pv_full, (n - 1) & 1);
          tc_fence_after();
          uint32_t v[32];
          tmem_ld32(tmem_PV + lane_addr + hf * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st32(tmem_PV + lane_addr + hf * 32, v);
          tmem_st_wait();
        }
        l_run *= alpha;
        if (need) m_used = mx;
      }
      const float m_safe = (m_used == -INFINITY) ? 0.f : m_used;
      if (stamper) a.stamps[4 * 64 + n] = clock64();
      // pass 2: P = exp2(x - m), row sum, bf16 into the swizzled A-operand tile
      float rs;
      if (thr2) {
        if (!full) rs = softmax_row_exp<true, true, true>(w, m_safe, tmem_p, t, dstream, dkp, thr2);
        else if (any_bias) rs = softmax_row_exp<false, true, true>(w, m_safe, tmem_p, t, dstream, dkp, thr2);
        else rs = softmax_row_exp<false, false, true>(w, m_safe, tmem_p, t, dstream, dkp, thr2);
      } else {
        if (!full) rs = softmax_row_exp<true, true, false>(w, m_safe, tmem_p, t, dstream, dkp, thr2);
        else if (any_bias) rs = softmax_row_exp<false, true, false>(w, m_safe, tmem_p, t, dstream, dkp, thr2);
        else rs = softmax_row_exp<false, false, false>(w, m_safe, tmem_p, t, dstream, dkp, thr2);
      }
      l_run += rs;
      tc_fence_before();  // P went to tensor memory (tcgen05.st + wait::st inside softmax_row_exp)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (stamper) a.stamps[5 * 64 + n] = clock64();
      ++n;
    }
    // epilogue: O row (this warp's 32 channels) / row sum, lse
    uint32_t v[32];
    if (n > 0) {
      mbar_wait(pv_full, (n - 1) & 1);
      tc_fence_after();
      tmem_ld32(tmem_PV + lane_addr + hf * 32, v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = 0u;
    }
    const float l_tot = l_run + pair_exchange(x_mine, x_other, lq, l_run);
    if (t < a.Tq) {
      const float inv = l_tot > 0.f ? a.drop.inv_keep / l_tot : 0.f;  // 1/(1-p) of the dropout folded in (1 when off)
      bf16* op = a.o + (long long)b * a.o_bs + (long long)t * a.o_rs + h * HD + hf * 32;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o4;
        o4.x = pack_bf16(__uint_as_float(v[8 * u]) * inv, __uint_as_float(v[8 * u + 1]) * inv);
        o4.y = pack_bf16(__uint_as_float(v[8 * u + 2]) * inv, __uint_as_float(v[8 * u + 3]) * inv);
        o4.z = pack_bf16(__uint_as_float(v[8 * u + 4]) * inv, __uint_as_float(v[8 * u + 5]) * inv);
        o4.w = pack_bf16(__uint_as_float(v[8 * u + 6]) * inv, __uint_as_float(v[8 * u + 7]) * inv);
        reinterpret_cast<uint4*>(op)[u] = o4;
      }
      if (a.lse && hf == 0) a.lse[((long long)b * a.H + h) * a.Tq + t] = l_tot > 0.f ? (m_used + log2f(l_tot)) * LN2 : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FWD_SW + 1) {
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
               q_len, kv_len, quirk_mod, nullptr};
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    configured = true;
  }
  static int fdbg = -1;
  static unsigned long long* fstamps = nullptr;
  if (fdbg < 0) {
    const char* e = getenv("OMR_ATTN_DEBUG");
    fdbg = e ? atoi(e) : 0;
    if (fdbg & 512) {
      cudaMalloc(&fstamps, 6 * 64 * 8);
      cudaMemset(fstamps, 0, 6 * 64 * 8);
    }
  }
  a.stamps = fstamps;
  dim3 grid((unsigned)((Tq + BQ - 1) / BQ), (unsigned)H, (unsigned)B);
  OmrLaunch(grid, 32 * (FWD_SW + 2), FWD_SMEM, st)(attn_fwd_tc_kernel, tmQ, tmK, tmV, a);
  OMR_LAUNCHED();
  if (fstamps) {  // debugging aid: clock stamps of CTA (0,0,0), relative to the first one
    static int printed = 0;
    unsigned long long h[6 * 64];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, fstamps, sizeof(h), cudaMemcpyDeviceToHost);
    if (printed++ == 3) {
      unsigned long long t0 = ~0ull;
      for (int i = 0; i < 6 * 64; ++i) if (h[i] && h[i] < t0) t0 = h[i];
      const char* names[6] = {"kv_issued", "s_issued", "pv_issued", "sm_s_ready", "sm_pass1_done", "sm_arrived"};
      for (int r = 0; r < 6; ++r) {
        fprintf(stderr, "%-16s", names[r]);
        for (int i = 0; i < 20; ++i) fprintf(stderr, " %6lld", h[r * 64 + i] ? (long long)(h[r * 64 + i] - t0) : -1ll);
        fprintf(stderr, "\n");
      }
    }
  }
  return OMR_OK;
}

// =================================================================================================================
// Backward.  PERSISTENT: one CTA per SM walks its share of the work items (key tile, head, batch); K and V of an item stay
// resident in shared memory and the CTA walks the query tiles that can see them, in HALF tiles of 64 queries.  Everything
// is computed TRANSPOSED (rows = keys) so that each element-wise thread owns one key row and the probabilities come out
// directly in the operand layouts of the three gradient GEMMs:
//     S^T  = K Q^T                 M=128 keys, N=64 queries, K=64           (TMEM, recomputed)
//     dP^T = V dO^T                same shape
//     P^T  = exp2(scale*S^T + bias_k - lse_q)   -> packed bf16 back into TENSOR memory (A operand of the dV MMA)
//     dS^T = P^T o (dP^T - delta_q)             -> bf16 into swizzled shared-memory chunks
//     dV  += P^T  dO               A = P^T  (TMEM),     B = dO half tile (MN-major)  accumulated in TMEM over the q tiles
//     dK  += dS^T Q                A = dS^T (K-major),  B = Q  half tile (MN-major)  accumulated in TMEM over the q tiles
//     dQ_t = dS K                  A = the two dS^T chunks of a 128-query tile read as ONE MN-major operand, B = K tile
//                                  (MN-major); per q tile, ADDED to an fp32 accumulation buffer by the TMA unit
// What the round-2 measurements (OMR_ATTN_DEBUG switches and in-kernel clock stamps, scripts/attn_stamps.py) said about
// the round-1 kernel (one CTA per item, MMA -> element-wise -> MMA back to back on 8 warps, 260 us for the C3 shape):
//   * 158 us were launch, TMEM allocation, barrier set-up, K/V load latency and drain of 2432 short-lived CTAs
//     -> persistent CTAs, a scheduler warp that announces live items and loads K/V two items ahead;
//   * a half-tile TMA load takes ~2500 clk from issue to landing -> 5-stage Q / dO / statistics ring (P^T moved from
//     shared to tensor memory to make room);
//   * every synchronisation hop (mbarrier wait, fence, arrive) costs ~100 clk, ~900 clk per half tile per role
//     -> the element-wise warps work in two GROUPS of 8 that ping-pong on alternate half tiles (TMEM S^T/dP^T buffer,
//     P^T buffer and barriers of parity g belong to group g), and scores / gradients are issued by two different warps;
//   * mbarrier.try_wait suspends the thread: a role that polls two barriers with it loses ~1000 clk per visit;
//   * the dQ drain (shared-memory transpose + red.global.add.v4) cost 1770 clk per tile on the element-wise warps, the
//     dK / dV drain (scattered 16-byte stores) ~3000 clk per item -> four dedicated warps stage 4 KB tiles and hand them
//     to the TMA unit (cp.reduce.async.bulk.tensor .add for dQ, plain tensor stores for dK / dV).
// A second tiny kernel scales the fp32 dQ sums and writes them in the caller's (strided, bf16) layout.
// =================================================================================================================
namespace {

constexpr int BQH = 64;            // queries per half tile
constexpr int HTILE = BQH * 128;   // bytes of a [64 x 64] bf16 tile
constexpr int NST = 5;             // Q / dO / statistics stages
constexpr int BWD_EW = 16;         // element-wise warps: group g = warps 8g .. 8g+7 takes the half tiles of parity g
// warps 16-19: dQ / dK / dV drain (warp q owns TMEM lanes 32q..), 20: Q/dO producer, 21: score MMAs (+ TMEM allocation),
// 22: K/V producer + item scheduler, 23: gradient MMAs
constexpr int BWD_W_DQ = 16, BWD_W_QDO = 20, BWD_W_MMA_S = 21, BWD_W_KV = 22, BWD_W_MMA_G = 23, BWD_WARPS = 24;
// (K, V)[2] | Q[NST] | dO[NST] | dS^T[4] | drain staging (4 warps x 4 KB) | stats[NST][128] | barriers + announcements
constexpr int BWD_OFF_Q = 4 * TILE, BWD_OFF_DO = BWD_OFF_Q + NST * HTILE,
              BWD_OFF_DS = BWD_OFF_DO + NST * HTILE, BWD_OFF_DQS = BWD_OFF_DS + 4 * TILE, BWD_OFF_STAT = BWD_OFF_DQS + 4 * 4096,
              BWD_OFF_BAR = BWD_OFF_STAT + NST * 128 * 4, BWD_SMEM = BWD_OFF_BAR + 256;
static_assert(BWD_SMEM <= 227 * 1024, "backward attention kernel: shared memory");

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
  bf16* dk; long long dk_bs, dk_rs;  // only the dead-item zero fill writes through these (live items use the TMA maps)
  bf16* dv; long long dv_bs, dv_rs;
  float scale;
  unsigned long long* stamps;  // OMR_ATTN_DEBUG & 256: clock stamps of CTA 0 ([role][64])
  int dbg;  // OMR_ATTN_DEBUG (timing experiments only, results are wrong): 2 = no dQ reduction, 4 = no element-wise math,
            // 8 = no dQ MMAs, 16 = no dV / dK MMAs, 32 = no S^T / dP^T MMAs, 64 = no Q / dO loads, 128 = no statistics loads
};

// shared -> global with an element-wise fp32 ADD performed by the TMA unit (bulk-group completion): the dQ tiles of the
// key tiles of one (batch, head) meet in the fp32 accumulation buffer without a single SM-issued atomic
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// element-wise work of one thread on 16 query columns of a half tile: key row r, columns [16 cg, 16 cg + 16) of the half
struct BwdRow {
  uint32_t tmem_s, tmem_dp;  // TMEM addresses of this thread's 16 S^T / dP^T values
  uint32_t tmem_p;           // TMEM address of its 8 packed-bf16 P^T columns (A operand of the dV MMA)
  const float* nlse;         // smem: -lse * log2(e) of the 64 queries of the half, followed by their 64 deltas
  uint8_t* drow;             // this key row inside the dS^T chunk (128 B per row, 128B-swizzled)
  int unit0;                 // first 16-byte unit of the row this thread writes (2 units = 16 queries)
  int r, j, t0;              // key row in the tile, key, first query of this thread's 16
  int t_lo, t_hi;            // queries that can see key j
  float bias, scale_log2;
  uint64_t* free_bar;        // non-null: wait for this phase before the first store -- the P^T buffer and the dS^T chunk
  uint32_t free_par;         // are still being read by the gradient MMAs of the group's previous half while this one computes
};

template <bool MASKED, bool DROP, bool BIAS0>
__device__ __forceinline__ void bwd_half_row(const BwdRow& w, uint32_t dstream, uint32_t dkp, uint32_t thr_hi, uint32_t dsh,
                                             float inv_keep) {
  uint32_t sv[16], dv[16];
  tmem_ld16(w.tmem_s, sv);
  tmem_ld16(w.tmem_dp, dv);
  tmem_ld_wait();
  uint32_t pk[8], dk[8];
#pragma unroll
  for (int e = 0; e < 16; e += 4) {
    const float4 l4 = *reinterpret_cast<const float4*>(w.nlse + e);
    const float4 d4 = *reinterpret_cast<const float4*>(w.nlse + 64 + e);
    float ls[4] = {l4.x, l4.y, l4.z, l4.w};
    const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
    float pr[4], fk[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!BIAS0) ls[k] += w.bias;
      pr[k] = ex2_approx(fmaf(__uint_as_float(sv[e + k]), w.scale_log2, ls[k]));
      if (MASKED) {
        const int t = w.t0 + e + k;
        pr[k] = (t >= w.t_lo && t <= w.t_hi) ? pr[k] : 0.f;
      }
    }
    if (DROP) {  // mask / (1-p) of the pairs (t, j).  Keys j and j^1 (neighbouring lanes) share their 2 x 2 blocks: the
      // even lane hashes the block of queries (t, t+1), the odd lane that of (t+2, t+3); both words travel.  The 15-bit
      // field of this key is moved to the top of the word (shift 17 for even keys: bits 0-14; 1 for odd keys: bits
      // 16-30) and compared in place:  field >= thr  <=>  (word << dsh) >= (thr << 17)
      const int t = w.t0 + e;
      const uint2 mine = attn_drop_block(dstream, (uint32_t)((t >> 1) + (w.j & 1)), (uint32_t)(w.j >> 1), dkp);
      const uint32_t ox = __shfl_xor_sync(0xffffffffu, mine.x, 1), oy = __shfl_xor_sync(0xffffffffu, mine.y, 1);
      const uint2 b0 = (w.j & 1) ? make_uint2(ox, oy) : mine;  // queries (t, t+1)
      const uint2 b1 = (w.j & 1) ? mine : make_uint2(ox, oy);  // queries (t+2, t+3)
      fk[0] = (b0.x << dsh) >= thr_hi ? inv_keep : 0.f;
      fk[1] = (b0.y << dsh) >= thr_hi ? inv_keep : 0.f;
      fk[2] = (b1.x << dsh) >= thr_hi ? inv_keep : 0.f;
      fk[3] = (b1.y << dsh) >= thr_hi ? inv_keep : 0.f;
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
  if (w.free_bar) {
    mbar_wait(w.free_bar, w.free_par);
    tc_fence_after();
  }
  tmem_st8(w.tmem_p, pk);  // queries (2c, 2c+1) of the thread's 16 in column c: K-contiguous pairs, lane = key row
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int sw = ((w.unit0 + u) ^ (w.r & 7)) << 4;
    *reinterpret_cast<uint4*>(w.drow + sw) = make_uint4(dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
  }
}

// one work item = one (key tile, head, batch); items are numbered key-tile-major so that a static round-robin over the
// persistent CTAs hands everybody the same mix of long (early key tiles of a causal call) and short items
struct BwdItem {
  int kt, h, b, j0, qt0, nh;  // nh = half tiles of 64 queries (0: no query sees the tile); always even
};
__device__ __forceinline__ BwdItem bwd_item(const AttnTcArgs& a, int idx) {
  BwdItem I;
  const int bh = a.B * a.H;
  I.kt = idx / bh;
  const int rem = idx - I.kt * bh;
  I.b = rem / a.H;
  I.h = rem - I.b * a.H;
  I.j0 = I.kt * BKV;
  int qt1;
  q_tile_range(a, I.j0, I.qt0, qt1);
  I.nh = 2 * (qt1 - I.qt0);
  return I;
}

__global__ void __launch_bounds__(32 * BWD_WARPS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV,
                                                                       const __grid_constant__ CUtensorMap tmDO,
                                                                       const __grid_constant__ CUtensorMap tmDQ,
                                                                       const __grid_constant__ CUtensorMap tmDK,
                                                                       const __grid_constant__ CUtensorMap tmDV,
                                                                       AttnBwdArgs g) {
  omr_pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  const AttnTcArgs& a = g.f;
  uint8_t* sKV = smem;               // [2] x (K tile, V tile)
  uint8_t* sQ = smem + BWD_OFF_Q;    // [NST] half tiles
  uint8_t* sDO = smem + BWD_OFF_DO;  // [NST]
  uint8_t* sDS = smem + BWD_OFF_DS;  // [4] chunks of [128 keys x 64 queries]; (0,1) and (2,3) are the two dS^T tiles in flight
  uint8_t* sDQS = smem + BWD_OFF_DQS;  // [4 drain warps] x 4 KB: [32 rows x 32 fp32] (dQ) or [32 rows x 64 bf16] (dK, dV)
  float* sStat = reinterpret_cast<float*>(smem + BWD_OFF_STAT);  // [NST][-lse*log2e (64) | delta (64)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BWD_OFF_BAR);
  uint64_t* kv_full = bars;                  // [2] K, V of an item landed, item_idx / item_flag written
  uint64_t* kv_empty = bars + 2;             // [2]
  uint64_t* qd_full = bars + 4;              // [NST] Q, dO half tiles landed, statistics written
  uint64_t* qd_empty = bars + 4 + NST;       // [NST]
  uint64_t* sdp_full = bars + 4 + 2 * NST;   // [2] S^T and dP^T of a half ready in TMEM buffer uu & 1
  uint64_t* pds_full = bars + 6 + 2 * NST;   // [2] P^T (TMEM) and dS^T (smem) of a half written (count 8: one group)
  uint64_t* mma2_done = bars + 8 + 2 * NST;  // [2] dV, dK (, dQ) MMAs of a half complete
  uint64_t* dq_free = bars + 10 + 2 * NST;   // the drain warps have read the dQ accumulator of a tile (count 4)
  uint64_t* dkv_free = bars + 11 + 2 * NST;  // ... the dK / dV accumulators of an item (count 4)
  uint64_t* dq_done = bars + 12 + 2 * NST;   // the dQ MMAs of a tile are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13 + 2 * NST);
  volatile int* item_idx = reinterpret_cast<volatile int*>(bars + 14 + 2 * NST);  // [2] item of K/V buffer i, -1 = no more
  volatile int* item_flag = item_idx + 2;                                        // [2] bit 0: all key biases zero
  volatile int* dq_info = item_idx + 4;    // [2][4] (batch*H + head, first query, last tile of its item?, first key) of dQ tile n & 1
  volatile int* dq_total = item_idx + 12;  // number of dQ tiles of this CTA, -1 while the element-wise warps are running

  const int warp = (int)warp_idx_sync(), lane = threadIdx.x & 31;
  const int nkt = (a.Tk + BKV - 1) / BKV;
  const int nitems = nkt * a.H * a.B;

  if (warp == BWD_W_QDO && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDK);
    tma_prefetch_desc(&tmDV);
    *dq_total = -1;
    mbar_init(dq_free, 4);
    mbar_init(dkv_free, 4);
    mbar_init(dq_done, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);  // the gradient MMAs of the item's last half + the dQ MMAs of its last tile
      mbar_init(&sdp_full[s], 1);
      mbar_init(&pds_full[s], BWD_EW / 2);
      mbar_init(&mma2_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == BWD_W_MMA_S) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bcast0(*tmem_slot);
  const uint32_t tmem_ST = tmem_base, tmem_DPT = tmem_base + 128, tmem_DV = tmem_base + 256, tmem_DK = tmem_base + 320,
                 tmem_DQ = tmem_base + 384, tmem_PT = tmem_base + 448;  // P^T: 2 x 32 columns of packed bf16

  // All producer / issuer loops are warp-uniform; single instructions sit under elect_one() (see tc_common.cuh).
  // Every consumer role follows the same stream of live items: it waits for kv_full[it & 1] and reads item_idx[it & 1].
  if (warp == BWD_W_KV) {
    // ---- K / V producer and item scheduler: finds this CTA's live items, announces them (index + "all key biases are
    // zero" flag next to the K/V barrier) and loads their K / V tiles up to two items ahead of the consumers, so that
    // neither the liveness test (global loads) nor the K / V latency ever sits between two items.  Dead items on the way
    // -- every key of the tile masked out (bias = -inf: the padded tail of the memory, or past Tk), or no query sees the
    // tile: P = 0 throughout, dK = dV = 0 and no contribution to dQ -- get their zeros here.
    int it = 0;
    for (int idx = blockIdx.x;; idx += gridDim.x) {
      bool live = false, b0 = true;
      BwdItem I{};
      if (idx < nitems) {
        I = bwd_item(a, idx);
        bool dead = true;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = I.j0 + i * 32 + lane;
          float bv = -INFINITY;
          if (j < a.Tk) bv = a.key_bias ? a.key_bias[(long long)I.b * a.Tk + j] : 0.f;
          dead = dead && !(bv > -INFINITY);
          b0 = b0 && bv == 0.f;
        }
        dead = __all_sync(0xffffffffu, dead);
        b0 = __all_sync(0xffffffffu, b0);
        live = !dead && I.nh > 0;
        if (!live) {
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int j = I.j0 + i * 32 + lane;
            if (j < a.Tk) {
              uint4* dkp_ = reinterpret_cast<uint4*>(g.dk + (long long)I.b * g.dk_bs + (long long)j * g.dk_rs + I.h * HD);
              uint4* dvp_ = reinterpret_cast<uint4*>(g.dv + (long long)I.b * g.dv_bs + (long long)j * g.dv_rs + I.h * HD);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                dkp_[u] = make_uint4(0, 0, 0, 0);
                dvp_[u] = make_uint4(0, 0, 0, 0);
              }
            }
          }
          continue;
        }
      }
      const int kb = it & 1;
      mbar_wait(&kv_empty[kb], ((it >> 1) & 1) ^ 1);
      if (lane == 0) {
        item_idx[kb] = live ? idx : -1;  // -1: no more items
        item_flag[kb] = b0 ? 1 : 0;
      }
      __syncwarp();
      if (elect_one()) {
        if (live) {
          mbar_expect_tx(&kv_full[kb], 2 * TILE);
          tma_load_3d(sKV + kb * 2 * TILE, &tmK, &kv_full[kb], I.h * HD, I.j0, I.b);
          tma_load_3d(sKV + kb * 2 * TILE + TILE, &tmV, &kv_full[kb], I.h * HD, I.j0, I.b);
        } else {
          mbar_arrive(&kv_full[kb]);
        }
      }
      __syncwarp();
      if (!live) break;
      ++it;
    }
  } else if (warp == BWD_W_QDO) {
    // ---- Q / dO / statistics producer: follows the announced items through the ring of half tiles ----
    uint32_t uu = 0;
    for (int it = 0;; ++it) {
      mbar_wait(&kv_full[it & 1], (it >> 1) & 1);
      const int idx = item_idx[it & 1];
      if (idx < 0) break;
      const BwdItem I = bwd_item(a, idx);
      const long long stat_base = ((long long)I.b * a.H + I.h) * a.Tq;
      // the statistics of a half (-lse in log2 units and delta of its 64 queries) are fetched one half ahead, so that
      // their latency hides behind the wait for the ring slot
      float nl[2], dl[2];
      auto load_stats = [&](int u, float (&n2)[2], float (&d2)[2]) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int t = I.qt0 * BQ + u * BQH + i * 32 + lane;
          n2[i] = (t < a.Tq && !(g.dbg & 128)) ? -a.lse[stat_base + t] * LOG2E : 0.f;
          d2[i] = (t < a.Tq && !(g.dbg & 128)) ? g.delta[stat_base + t] : 0.f;
        }
      };
      load_stats(0, nl, dl);
      for (int u = 0; u < I.nh; ++u, ++uu) {
        const int s = uu % NST, q0 = I.qt0 * BQ + u * BQH;
        float nl2[2] = {0.f, 0.f}, dl2[2] = {0.f, 0.f};
        if (u + 1 < I.nh) load_stats(u + 1, nl2, dl2);
        mbar_wait(&qd_empty[s], ((uu / NST) & 1) ^ 1);
        float* st = sStat + s * 128;  // plain stores, published by the arrive below
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          st[i * 32 + lane] = nl[i];
          st[64 + i * 32 + lane] = dl[i];
          nl[i] = nl2[i];
          dl[i] = dl2[i];
        }
        __syncwarp();
        if (elect_one()) {
          if (g.dbg & 64) {
            mbar_arrive(&qd_full[s]);
          } else {
            mbar_expect_tx(&qd_full[s], 2 * HTILE);
            tma_load_3d(sQ + s * HTILE, &tmQ, &qd_full[s], I.h * HD, q0, I.b);
            tma_load_3d(sDO + s * HTILE, &tmDO, &qd_full[s], I.h * HD, q0, I.b);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == BWD_W_MMA_S) {
    // ---- score MMAs: S^T and dP^T of half uu into TMEM buffer uu & 1, as soon as its Q / dO tiles have landed and the
    // element-wise group of that parity has finished half uu - 2 (the buffer's previous tenant) ----
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);   // N = 64 queries
    constexpr uint32_t idesc_mn = make_idesc_bf16(128, 64, 1, 1);  // A MN-major (dS), B MN-major
    const uint32_t kv_base = smem_u32(sKV), q_base = smem_u32(sQ), do_base = smem_u32(sDO), ds_base = smem_u32(sDS);
    uint32_t uu = 0;
    uint32_t kq_addr = 0;  // K tile of the item the pending dQ tile belongs to
    // dQ_tile = dS K of the 128-query tile whose second (odd) half is v: A = both dS^T chunks of the tile read as ONE
    // MN-major operand (M = 128 queries, K = key rows).  Issued here, not by the gradient warp (measured: that warp was the
    // bottleneck at ~1400 clk per odd half), once the drain warps have read the previous tile out of the accumulator.
    auto issue_dq = [&](uint32_t v, uint32_t k_addr, int release_kv) {  // release_kv >= 0: last tile of the item in that K/V buffer
      if (v > 1) mbar_wait(dq_free, ((v >> 1) - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        if (!(g.dbg & 8)) {
          const uint32_t ds2 = ds_base + ((v & 3) - 1) * TILE;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j > 0)
              umma_bf16_acc(tmem_DQ, make_smem_desc(ds2 + j * 2048, TILE, 1024, 128), make_smem_desc(k_addr + j * 2048, 0, 1024, 128), idesc_mn);
            else
              umma_bf16_new(tmem_DQ, make_smem_desc(ds2, TILE, 1024, 128), make_smem_desc(k_addr, 0, 1024, 128), idesc_mn);
          }
        }
        umma_commit(dq_done);
        if (release_kv >= 0) umma_commit(&kv_empty[release_kv]);
      }
      __syncwarp();
    };
    int it = 0;
    for (;; ++it) {
      mbar_wait(&kv_full[it & 1], (it >> 1) & 1);
      const int idx = item_idx[it & 1];
      if (idx < 0) break;
      const int nh = bwd_item(a, idx).nh;
      const uint32_t k_addr = kv_base + (uint32_t)(it & 1) * 2 * TILE, v_addr = k_addr + TILE;
      for (int u = 0; u < nh; ++u, ++uu) {
        const int s = uu % NST;
        if (uu >= 2) {
          mbar_wait(&pds_full[uu & 1], ((uu >> 1) - 1) & 1);  // the element-wise group is done with half uu - 2
          if (uu & 1) issue_dq(uu - 2, u >= 2 ? k_addr : kq_addr, u >= 2 ? -1 : ((it - 1) & 1));
        }
        mbar_wait(&qd_full[s], (uu / NST) & 1);
        tc_fence_after();
        const uint32_t q_addr = q_base + (uint32_t)s * HTILE, do_addr = do_base + (uint32_t)s * HTILE;
        const uint32_t dS = tmem_ST + (uu & 1) * 64, dP = tmem_DPT + (uu & 1) * 64;
        if (elect_one()) {
          if (!(g.dbg & 32)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j > 0)
                umma_bf16_acc(dS, make_smem_desc(k_addr + j * 32, 16, 1024, 128), make_smem_desc(q_addr + j * 32, 16, 1024, 128), idesc_s);
              else
                umma_bf16_new(dS, make_smem_desc(k_addr, 16, 1024, 128), make_smem_desc(q_addr, 16, 1024, 128), idesc_s);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j > 0)
                umma_bf16_acc(dP, make_smem_desc(v_addr + j * 32, 16, 1024, 128), make_smem_desc(do_addr + j * 32, 16, 1024, 128), idesc_s);
              else
                umma_bf16_new(dP, make_smem_desc(v_addr, 16, 1024, 128), make_smem_desc(do_addr, 16, 1024, 128), idesc_s);
            }
          }
          umma_commit(&sdp_full[uu & 1]);
        }
        __syncwarp();
        if (g.stamps && blockIdx.x == 0 && lane == 0 && uu < 64) g.stamps[1 * 64 + uu] = clock64();
      }
      kq_addr = k_addr;
    }
    if (uu > 0) {  // the last tile of this CTA
      mbar_wait(&pds_full[0], ((uu - 2) >> 1) & 1);
      mbar_wait(&pds_full[1], ((uu - 1) >> 1) & 1);
      issue_dq(uu - 1, kq_addr, (it - 1) & 1);
    }
  } else if (warp == BWD_W_MMA_G) {
    // ---- gradient MMAs of half uu once its P^T / dS^T are written.  The dV / dK accumulators are single-buffered per item
    // and the dQ accumulator per tile: the first half of an item waits for the drain of the previous item's dK / dV, an odd
    // half (which ends a 128-query tile) for the drain of the previous dQ tile. ----
    constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);  // A K-major (dS^T) or TMEM (P^T), B MN-major
    const uint32_t ds_base = smem_u32(sDS), q_base = smem_u32(sQ), do_base = smem_u32(sDO);
    uint32_t uu = 0;
    for (int it = 0;; ++it) {
      mbar_wait(&kv_full[it & 1], (it >> 1) & 1);
      const int idx = item_idx[it & 1];
      if (idx < 0) break;
      const int nh = bwd_item(a, idx).nh;
      for (int u = 0; u < nh; ++u, ++uu) {
        const int s = uu % NST;
        const uint32_t q_addr = q_base + (uint32_t)s * HTILE, do_addr = do_base + (uint32_t)s * HTILE;
        const uint32_t pt_tmem = tmem_PT + (uu & 1) * 32, ds_addr = ds_base + (uu & 3) * TILE;
        if (u == 0 && it > 0) mbar_wait(dkv_free, (it - 1) & 1);
        mbar_wait(&pds_full[uu & 1], (uu >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          if (!(g.dbg & 16)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // dV += P^T dO   (reduction over the 64 queries of the half)
              if (j > 0)
                umma_bf16_ta(tmem_DV, pt_tmem + j * 8, make_smem_desc(do_addr + j * 2048, 0, 1024, 128), idesc_kn, 1u);
              else
                umma_bf16_ta(tmem_DV, pt_tmem, make_smem_desc(do_addr, 0, 1024, 128), idesc_kn, u > 0 ? 1u : 0u);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // dK += dS^T Q
              if (j > 0)
                umma_bf16_acc(tmem_DK, make_smem_desc(ds_addr + j * 32, 16, 1024, 128), make_smem_desc(q_addr + j * 2048, 0, 1024, 128), idesc_kn);
              else
                umma_bf16(tmem_DK, make_smem_desc(ds_addr, 16, 1024, 128), make_smem_desc(q_addr, 0, 1024, 128), idesc_kn, u > 0 ? 1u : 0u);
            }
          }
          umma_commit(&mma2_done[uu & 1]);
          umma_commit(&qd_empty[s]);
          if (u == nh - 1) umma_commit(&kv_empty[it & 1]);
        }
        __syncwarp();
        if (g.stamps && blockIdx.x == 0 && lane == 0 && uu < 64) g.stamps[2 * 64 + uu] = clock64();
      }
    }
  } else if (warp >= BWD_W_DQ && warp < BWD_W_DQ + 4) {
    // ---- drain warps: warp q owns TMEM lanes 32q .. 32q+31 = query rows of the dQ accumulator, key rows of dK / dV.
    // Each finished dQ tile goes, 32 channels at a time, through the warp's 4 KB staging tile to the TMA unit, which ADDS
    // it to the fp32 accumulation buffer (rows past Tq are clipped by the tensor map); after the last tile of an item the
    // warp's 32 rows of dK (x scale) and dV follow as bf16 tensor stores (rows past Tk clipped). ----
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* stage = sDQS + q * 4096;
    for (uint32_t n = 0;; ++n) {
      bool done = false;
      while (!bcast0(mbar_try(dq_done, n & 1) ? 1u : 0u)) {  // the dQ MMAs of tile n
        const int tot = *dq_total;
        if (tot >= 0 && n >= (uint32_t)tot) { done = true; break; }
      }
      if (done) break;
      tc_fence_after();
      const int bh = dq_info[(n & 1) * 4], t0 = dq_info[(n & 1) * 4 + 1], last = dq_info[(n & 1) * 4 + 2], j0 = dq_info[(n & 1) * 4 + 3];
      {
        // the whole accumulator row goes to registers first, so that the accumulator is released before any staging wait
        uint32_t v0[32], v1[32];
        tmem_ld32(tmem_DQ + lane_addr, v0);
        tmem_ld32(tmem_DQ + lane_addr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
        if (g.stamps && blockIdx.x == 0 && q == 0 && lane == 0 && n < 32) g.stamps[0 * 64 + 2 * n + 1] = clock64();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (elect_one()) tma_store_wait_read<0>();  // the staging tile's previous transfer has been read
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(stage + lane * 128 + ((u ^ (lane & 7)) << 4)) =
                c == 0 ? make_uint4(v0[4 * u], v0[4 * u + 1], v0[4 * u + 2], v0[4 * u + 3])
                       : make_uint4(v1[4 * u], v1[4 * u + 1], v1[4 * u + 2], v1[4 * u + 3]);
          fence_proxy_async();
          __syncwarp();
          if (elect_one() && !(g.dbg & 2)) {
            tma_reduce_add_3d(&tmDQ, stage, c * 32, t0 + q * 32, bh);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
      if (last) {
        mbar_wait(&mma2_done[1], n & 1);  // the gradient MMAs of the item's last half (global half 2n+1)
        tc_fence_after();
        const int b = bh / a.H, h = bh - b * a.H;
        uint32_t pk[2][32];  // 64 channels of this key row of dK (x scale) and dV, packed bf16
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const float mul = which == 0 ? g.scale : 1.f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32((which == 0 ? tmem_DK : tmem_DV) + lane_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[which][c * 16 + e] = pack_bf16(__uint_as_float(v[2 * e]) * mul, __uint_as_float(v[2 * e + 1]) * mul);
          }
        }
        tc_fence_before();  // dK and dV are in registers: the next item may start accumulating
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_free);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(stage + lane * 128 + ((u ^ (lane & 7)) << 4)) =
                make_uint4(pk[which][4 * u], pk[which][4 * u + 1], pk[which][4 * u + 2], pk[which][4 * u + 3]);
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            tma_store_3d(which == 0 ? &tmDK : &tmDV, stage, h * HD, j0 + q * 32, b);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
    }
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  } else if (warp < BWD_EW) {
    // ---- 16 element-wise warps in two groups: group grp = warp >> 3 takes the half tiles of parity grp.  Thread = key row
    // r (TMEM lane quarter warp & 3) x 32 query columns (half ch of the half tile), in two passes of 16 columns ----
    const int grp = warp >> 3, lq = warp & 3, ch = (warp >> 2) & 1;
    const int r = lq * 32 + lane;
    const int off = a.Tk - a.Tq;
    const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
    const uint32_t dseed = a.drop.thr ? attn_drop_seed(a.drop) : 0u;
    const uint32_t dkp = (uint32_t)((a.Tk + 1) >> 1), dsh = (r & 1) ? 1u : 17u, thr_hi = a.drop.thr << 17;
    const float ik = a.drop.inv_keep;
    const bool stamper = g.stamps && blockIdx.x == 0 && lane == 0 && (warp & 7) == 0;
    auto row_bias = [&](int idx) -> float {  // this thread's key bias (log2 units) in item idx
      const BwdItem I = bwd_item(a, idx);
      const int j = I.j0 + r;
      if (j >= a.Tk) return -INFINITY;
      return a.key_bias ? a.key_bias[(long long)I.b * a.Tk + j] * LOG2E : 0.f;
    };

    int it = 0;
    uint32_t uu0 = 0;  // global index of the current item's first half
    mbar_wait(&kv_full[0], 0);
    int idx = item_idx[0];
    float bias = idx >= 0 ? row_bias(idx) : 0.f;
    while (idx >= 0) {
      const BwdItem I = bwd_item(a, idx);
      const bool bias0 = (item_flag[it & 1] & 1) != 0;
      const int j = I.j0 + r;
      // queries that can see key j: t in [t_lo, t_hi]
      int t_lo = 0, t_hi = a.Tq - 1;
      if (a.causal) {
        t_lo = max(0, j - off);
        if (a.window > 0) t_hi = min(t_hi, j - off + a.window);
      }
      if (j >= a.Tk) t_hi = -1;
      int lqm, lkv;
      block_mask_of(a, I.b, I.h, lqm, lkv);
      if (j >= lkv) t_hi = min(t_hi, lqm - 1);
      const uint32_t dstream = a.drop.thr ? attn_drop_stream_of(dseed, I.b * a.H + I.h) : 0u;
      int idx_next = -1;
      float bias_next = 0.f;
      for (int u = grp; u < I.nh; u += 2) {
        const uint32_t uu = uu0 + (uint32_t)u;
        const int s = uu % NST;
        const int q0 = I.qt0 * BQ + u * BQH;
        if (u + 2 >= I.nh) {  // this group's last half of the item: the next item has long been announced -- fetch its bias
          mbar_wait(&kv_full[(it + 1) & 1], ((it + 1) >> 1) & 1);
          idx_next = item_idx[(it + 1) & 1];
          if (idx_next >= 0) bias_next = row_bias(idx_next);
        }
        if (stamper && uu < 64) g.stamps[3 * 64 + uu] = clock64();
        mbar_wait(&qd_full[s], (uu / NST) & 1);  // statistics of the half
        mbar_wait(&sdp_full[grp], (uu >> 1) & 1);
        tc_fence_after();
        if (stamper && uu < 64) g.stamps[4 * 64 + uu] = clock64();
        // every (query, key) pair of this half visible?  (then no interval tests; rows past Tk carry bias = -inf)
        bool full = q0 + BQH <= a.Tq && I.j0 + BKV <= a.Tk;
        if (a.causal) full = full && q0 >= I.j0 + BKV - 1 - off && (a.window <= 0 || q0 + BQH - 1 <= I.j0 - off + a.window);
        full = full && (I.j0 + BKV <= lkv || q0 + BQH <= lqm);
        if (g.dbg & 4) {
          if (uu >= 2) mbar_wait(&mma2_done[grp], ((uu >> 1) - 1) & 1);
        } else {
#pragma unroll 1
          for (int cg = 2 * ch; cg < 2 * ch + 2; ++cg) {
            BwdRow w;
            w.tmem_s = tmem_ST + grp * 64 + lane_addr + cg * 16;
            w.tmem_dp = tmem_DPT + grp * 64 + lane_addr + cg * 16;
            w.tmem_p = tmem_PT + grp * 32 + lane_addr + cg * 8;
            w.nlse = sStat + s * 128 + cg * 16;
            w.drow = sDS + (uu & 3) * TILE + r * 128;
            w.unit0 = cg * 2;
            w.r = r; w.j = j; w.t0 = q0 + cg * 16;
            w.t_lo = t_lo; w.t_hi = t_hi;
            w.bias = bias; w.scale_log2 = a.scale_log2;
            // the group's previous half (uu - 2): its gradient MMAs must be done before P^T buffer grp / dS^T chunk uu & 3
            // (last read two tiles ago, by MMAs issued earlier still) are written -- but not before they are computed
            w.free_bar = (uu >= 2 && cg == 2 * ch) ? &mma2_done[grp] : nullptr;
            w.free_par = ((uu >> 1) - 1) & 1;
            if (a.drop.thr) {
              if (!full) bwd_half_row<true, true, false>(w, dstream, dkp, thr_hi, dsh, ik);
              else if (bias0) bwd_half_row<false, true, true>(w, dstream, dkp, thr_hi, dsh, ik);
              else bwd_half_row<false, true, false>(w, dstream, dkp, thr_hi, dsh, ik);
            } else {
              if (!full) bwd_half_row<true, false, false>(w, dstream, dkp, thr_hi, dsh, ik);
              else if (bias0) bwd_half_row<false, false, true>(w, dstream, dkp, thr_hi, dsh, ik);
              else bwd_half_row<false, false, false>(w, dstream, dkp, thr_hi, dsh, ik);
            }
          }
          tmem_st_wait();
        }
        if (grp == 1 && (warp & 7) == 0 && lane == 0) {  // where the dQ tile completed by this (odd) half belongs
          volatile int* di = dq_info + ((uu >> 1) & 1) * 4;
          di[0] = I.b * a.H + I.h;
          di[1] = (I.qt0 + (u >> 1)) * BQ;
          di[2] = (u == I.nh - 1) ? 1 : 0;
          di[3] = I.j0;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[grp]);
        if (stamper && uu < 64) g.stamps[5 * 64 + uu] = clock64();
      }
      uu0 += (uint32_t)I.nh;
      ++it;
      idx = idx_next;
      bias = bias_next;
    }
    if (grp == 1 && (warp & 7) == 0 && lane == 0) *dq_total = (int)(uu0 >> 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BWD_W_MMA_S) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Preparation pass of the backward: delta[b,h,t] = sum_d dO * O and the zeroing of the fp32 dQ accumulation row, in one
// sweep (round 1: a 17 MB memset and a one-warp-per-row delta kernel, 7 + 13 us at the C3 shape).  Eight lanes share a
// (b, h, t) row: one 16-byte load per tensor and lane, three shuffles, two 16-byte stores of zeros.
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ o, long long o_bs, long long o_rs, const bf16* __restrict__ dO,
                                     long long do_bs, long long do_rs, float* __restrict__ delta, float* __restrict__ dq_acc,
                                     int B, int H, int Tq) {
  omr_pdl_enter();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long w = gid >> 3;  // row
  const int sub = (int)(gid & 7);
  const bool ok = w < (long long)B * H * Tq;
  float s = 0.f;
  if (ok) {
    const int t = (int)(w % Tq);
    const long long rr = w / Tq;
    const int h = (int)(rr % H), b = (int)(rr / H);
    const uint4 ov = *reinterpret_cast<const uint4*>(o + (long long)b * o_bs + (long long)t * o_rs + h * HD + sub * 8);
    const uint4 dv = *reinterpret_cast<const uint4*>(dO + (long long)b * do_bs + (long long)t * do_rs + h * HD + sub * 8);
    const uint32_t oa[4] = {ov.x, ov.y, ov.z, ov.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&oa[i]), y = *reinterpret_cast<const __nv_bfloat162*>(&da[i]);
      s = fmaf(__low2float(x), __low2float(y), s);
      s = fmaf(__high2float(x), __high2float(y), s);
    }
    float4* z = reinterpret_cast<float4*>(dq_acc + w * HD + sub * 8);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) delta[w] = s;
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
  int rc = make_head_map(&tmQ, q, q_bs, q_rs, B, Tq, H, BQH);
  if (rc) return rc;
  rc = make_head_map(&tmK, k, k_bs, k_rs, B, Tk, H, BKV);
  if (rc) return rc;
  rc = make_head_map(&tmV, v, v_bs, v_rs, B, Tk, H, BKV);
  if (rc) return rc;
  rc = make_head_map(&tmDO, dout, do_bs, do_rs, B, Tq, H, BQH);
  if (rc) return rc;
  const long long rows = (long long)B * H * Tq;
  if ((long long)B * H > 0x7fffffffll) return OMR_TC_NOT_ELIGIBLE;
  float* delta = ws;
  float* dq_acc = ws + ((rows + 3) / 4) * 4;  // keep the accumulators 16-byte aligned
  CUtensorMap tmDQ;  // fp32 [B*H][Tq][64], boxes of 32 rows x 32 channels
  {
    unsigned long long dims[3] = {(unsigned long long)HD, (unsigned long long)Tq, (unsigned long long)B * H};
    unsigned long long strides[2] = {(unsigned long long)HD * 4, (unsigned long long)Tq * HD * 4};
    unsigned int box[3] = {32u, 32u, 1u};
    rc = omr_make_tensor_map(&tmDQ, 4, dq_acc, 3, dims, strides, box, nullptr, 128);
    if (rc) return rc;
  }
  OmrLaunch((unsigned)((rows * 8 + 255) / 256), 256, 0, st)(attn_bwd_prep_kernel, (const bf16*)o, o_bs, o_rs, (const bf16*)dout, do_bs,
                                                           do_rs, delta, dq_acc, B, H, Tq);
  OMR_LAUNCHED();
  CUtensorMap tmDK, tmDV;  // bf16 stores of 32 key rows x 64 channels (the drain warps)
  rc = make_head_map(&tmDK, dk, dk_bs, dk_rs, B, Tk, H, 32);
  if (rc) return rc;
  rc = make_head_map(&tmDV, dv, dv_bs, dv_rs, B, Tk, H, 32);
  if (rc) return rc;
  AttnBwdArgs g{};
  g.f = AttnTcArgs{nullptr, 0, 0, const_cast<float*>(lse), key_bias, B, H, Tq, Tk, scale * LOG2E, causal, window,
                   omr_attn_cur_dropout(), q_len, kv_len, quirk_mod};
  g.delta = delta;
  g.dk = (bf16*)dk; g.dk_bs = dk_bs; g.dk_rs = dk_rs;
  g.dv = (bf16*)dv; g.dv_bs = dv_bs; g.dv_rs = dv_rs;
  g.scale = scale;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("OMR_ATTN_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  g.dbg = dbg;
  static unsigned long long* stamps = nullptr;
  if ((dbg & 256) && !stamps) {
    cudaMalloc(&stamps, 6 * 64 * 8);
    cudaMemset(stamps, 0, 6 * 64 * 8);
  }
  g.stamps = stamps;
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    configured = true;
  }
  const long long nitems = (long long)((Tk + BKV - 1) / BKV) * H * B;
  if (nitems > (1ll << 30)) return OMR_TC_NOT_ELIGIBLE;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const unsigned grid = (unsigned)(nitems < n_sm ? nitems : n_sm);  // persistent: one CTA per SM
  OmrLaunch(grid, 32 * BWD_WARPS, BWD_SMEM, st)(attn_bwd_tc_kernel, tmQ, tmK, tmV, tmDO, tmDQ, tmDK, tmDV, g);
  OMR_LAUNCHED();
  if (stamps) {  // debugging aid: clock stamps of CTA 0's first 64 half tiles, relative to the first one
    static int printed = 0;
    unsigned long long h[6 * 64];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost);
    if (printed++ == 3) {
      unsigned long long t0 = ~0ull;
      for (int i = 0; i < 6 * 64; ++i) if (h[i] && h[i] < t0) t0 = h[i];
      const char* names[6] = {"dq_freed(odd)", "scores_issued", "grads_issued", "ew_begin", "ew_inputs_ready", "ew_arrived"};
      for (int r = 0; r < 6; ++r) {
        fprintf(stderr, "%-16s", names[r]);
        for (int i = 0; i < 40; ++i) fprintf(stderr, " %6lld", h[r * 64 + i] ? (long long)(h[r * 64 + i] - t0) : -1ll);
        fprintf(stderr, "\n");
      }
    }
  }
  OmrLaunch((unsigned)((rows * 8 + 255) / 256), 256, 0, st)(attn_dq_finalize_kernel, dq_acc, (bf16*)dq, dq_bs, dq_rs, B, H, Tq, scale);
  OMR_LAUNCHED();
  return OMR_OK;
}
