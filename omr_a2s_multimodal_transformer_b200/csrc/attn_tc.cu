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
};

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

__global__ void __launch_bounds__(192, 2) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                             const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV, AttnTcArgs a) {
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  int kt0, kt1;
  kv_tile_range(a, q0, kt0, kt1);
  const int ntiles = kt1 - kt0;

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
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_PV = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(q_full, TILE);
      tma_load_3d(sQ, &tmQ, q_full, h * HD, q0, b);
      for (int i = 0; i < ntiles; ++i) {
        const int s = i & 1;
        mbar_wait(&kv_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * TILE);
        tma_load_3d(sK + s * TILE, &tmK, &kv_full[s], h * HD, (kt0 + i) * BKV, b);
        tma_load_3d(sV + s * TILE, &tmV, &kv_full[s], h * HD, (kt0 + i) * BKV, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && ntiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      mbar_wait(q_full, 0);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      for (int i = 0; i < ntiles; ++i) {
        const int s = i & 1;
        mbar_wait(&kv_full[s], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * TILE), v_addr = smem_u32(sV + s * TILE);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16(tmem_S, make_smem_desc(q_addr + j * 32, 16, 1024, 128), make_smem_desc(k_addr + j * 32, 16, 1024, 128), idesc_s,
                    j > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16(tmem_PV, make_smem_desc(p_addr + (j >> 2) * TILE + (j & 3) * 32, 16, 1024, 128),
                    make_smem_desc(v_addr + j * 2048, 0, 1024, 128), idesc_pv, j > 0 ? 1u : 0u);
        umma_commit(pv_full);
        umma_commit(&kv_empty[s]);
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
    // visible key interval of this row
    int k_hi = a.Tk - 1, k_lo = 0;
    if (a.causal) {
      k_hi = min(k_hi, t + off);
      if (a.window > 0) k_lo = max(0, t + off - a.window);
    }
    if (t >= a.Tq) k_hi = -1;
    for (int i = 0; i < ntiles; ++i) {
      const int j0 = (kt0 + i) * BKV;
      // key-bias tile (pre-multiplied by log2 e) -> smem, shared by the 128 rows
      softmax_bar();  // previous tile's readers are done with sBias
      {
        const int j = j0 + r;
        sBias[r] = (kb && j < a.Tk) ? kb[j] * LOG2E : 0.f;
      }
      softmax_bar();
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int j = j0 + c * 32 + e;
          const float x = fmaf(__uint_as_float(v[e]), a.scale_log2, sBias[c * 32 + e]);
          mx = fmaxf(mx, (j >= k_lo && j <= k_hi) ? x : -INFINITY);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_safe);
      if (i > 0) {
        mbar_wait(pv_full, (i - 1) & 1);
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
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const int j = j0 + c * 32 + e;
          float x0 = fmaf(__uint_as_float(v[e]), a.scale_log2, sBias[c * 32 + e]);
          float x1 = fmaf(__uint_as_float(v[e + 1]), a.scale_log2, sBias[c * 32 + e + 1]);
          float p0 = (j >= k_lo && j <= k_hi) ? exp2f(x0 - m_safe) : 0.f;
          float p1 = (j + 1 >= k_lo && j + 1 <= k_hi) ? exp2f(x1 - m_safe) : 0.f;
          // the row sum uses the bf16-rounded probabilities that the P V product will see
          __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          rs += __low2float(pb) + __high2float(pb);
          pk[e >> 1] = *reinterpret_cast<uint32_t*>(&pb);
        }
        // keys c*32 .. c*32+31 = 64 bytes = 4 sixteen-byte units of row r in chunk c/2
        uint8_t* rowp = sP + (c >> 1) * TILE + r * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int unit = (c & 1) * 4 + u;
          *reinterpret_cast<uint4*>(rowp + ((unit ^ (r & 7)) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    if (ntiles > 0) {
      mbar_wait(pv_full, (ntiles - 1) & 1);
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
  (void)quirk_mod;
  if (hd != HD || q_len || kv_len || B < 1 || H < 1 || Tq < 1 || Tk < 1) return OMR_TC_NOT_ELIGIBLE;
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
  AttnTcArgs a{(bf16*)o, o_bs, o_rs, lse, key_bias, B, H, Tq, Tk, scale * LOG2E, causal, window};
  static bool configured = false;
  if (!configured) {
    OMR_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    configured = true;
  }
  dim3 grid((unsigned)((Tq + BQ - 1) / BQ), (unsigned)H, (unsigned)B);
  attn_fwd_tc_kernel<<<grid, 192, FWD_SMEM, st>>>(tmQ, tmK, tmV, a);
  OMR_LAUNCHED();
  return OMR_OK;
}
