// attn_simt.cu -- flash-style scaled-dot-product attention (head_dim 64) on CUDA cores with exact
// fp32 accumulation: forward (online softmax, never materialises [Tq,Tk]) and backward
// (recompute from the saved log-sum-exp).  fp32-mode path and validation reference for the
// tcgen05 attention kernels.
//
// Mask algebra follows torch's multi_head_attention_forward as driven by the reference
// (SURVEY.md appendix B): additive fp32 key bias per (b, key) (0 / +1.0 / -inf), causal or
// sliding-window structure, and the mixer's (rows >= lq) x (cols >= lkv) block mask with its
// head-major repeat quirk (quirk_mod).
#include "attn_drop.cuh"
#include "simt_tile.cuh"

namespace {

constexpr int HD = 64, BT = 64, LDT = 68;
constexpr int TILE_F = BT * LDT;  // floats per smem tile

struct AttnArgs {
  const void *q, *k, *v, *o, *dout;
  void *o_w, *dq, *dk, *dv;
  float* lse; float* delta;
  const float* key_bias;
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, do_bs, do_rs, dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  int B, H, Tq, Tk;
  float scale;
  int causal, window;
  const int* q_len; const int* kv_len; int quirk_mod;
  AttnDrop drop;  // attention-probability dropout (thr == 0: off)
};

// mask-and-scale factor of pair (t, j): 0 if dropped, 1/(1-p) if kept, 1 when dropout is off
__device__ __forceinline__ float drop_factor(const AttnArgs& a, uint32_t stream, int t, int j) {
  if (a.drop.thr == 0) return 1.f;
  const uint2 blk = attn_drop_block(stream, (uint32_t)(t >> 1), (uint32_t)(j >> 1), (uint32_t)((a.Tk + 1) >> 1));
  return attn_drop_u15(blk, t, j) >= a.drop.thr ? a.drop.inv_keep : 0.f;
}

template <typename T>
__device__ __forceinline__ float fexp(float x) {
  return expf(x);
}
template <>
__device__ __forceinline__ float fexp<bf16>(float x) {
  return __expf(x);
}

// 64x64 tile: rows r0.. of a [*, rs]-strided matrix (64 contiguous columns at g) -> smem.
template <typename T, bool NAT, bool TR>
__device__ __forceinline__ void load_tile(const T* __restrict__ g, long long rs, int r0, int nrows,
                                          float* __restrict__ nat, float* __restrict__ tr, float mul) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    int idx = threadIdx.x + it * 256;
    int row = idx >> 4, c4 = idx & 15;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r0 + row < nrows) {
      load4(g + (long long)(r0 + row) * rs + c4 * 4, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] *= mul;
    }
    if (NAT) *reinterpret_cast<float4*>(nat + row * LDT + c4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
    if (TR) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tr[(c4 * 4 + k) * LDT + row] = v[k];
    }
  }
}

struct MaskCtx {
  const float* kb;  // key bias row for this batch element or nullptr
  int off;          // Tk - Tq
  int lq, lkv;      // mixer block mask (lq < 0: none)
};

__device__ __forceinline__ MaskCtx make_mask_ctx(const AttnArgs& a, int b, int h) {
  MaskCtx m;
  m.kb = a.key_bias ? a.key_bias + (long long)b * a.Tk : nullptr;
  m.off = a.Tk - a.Tq;
  m.lq = -1; m.lkv = 0;
  if (a.q_len && a.kv_len) {
    int s = a.quirk_mod > 0 ? (b * a.H + h) % a.quirk_mod : b;
    m.lq = a.q_len[s]; m.lkv = a.kv_len[s];
  }
  return m;
}

// additive mask term for (query t, key j); -inf when the pair is excluded
__device__ __forceinline__ float mask_term(const AttnArgs& a, const MaskCtx& m, int t, int j) {
  if (j >= a.Tk || t >= a.Tq) return -INFINITY;
  if (a.causal) {
    if (j > t + m.off) return -INFINITY;
    if (a.window > 0 && j < t + m.off - a.window) return -INFINITY;
  }
  if (m.lq >= 0 && t >= m.lq && j >= m.lkv) return -INFINITY;
  return m.kb ? m.kb[j] : 0.f;
}

// key-tile range [kt0, kt1) that can contain unmasked pairs for query rows [t0, t1]
__device__ __forceinline__ void key_tile_range(const AttnArgs& a, int t0, int t1, int& kt0, int& kt1) {
  int nkt = (a.Tk + BT - 1) / BT;
  kt0 = 0; kt1 = nkt;
  if (a.causal) {
    int off = a.Tk - a.Tq;
    int jmax = t1 + off;
    if (jmax < 0) { kt1 = 0; return; }
    int e = jmax / BT + 1;
    if (e < kt1) kt1 = e;
    if (a.window > 0) {
      int jmin = t0 + off - a.window;
      if (jmin > 0) kt0 = jmin / BT;
    }
  }
}
// query-tile range for key rows [j0, j1]
__device__ __forceinline__ void query_tile_range(const AttnArgs& a, int j0, int j1, int& qt0, int& qt1) {
  int nqt = (a.Tq + BT - 1) / BT;
  qt0 = 0; qt1 = nqt;
  if (a.causal) {
    int off = a.Tk - a.Tq;
    int tmin = j0 - off;  // t >= j - off
    if (tmin > 0) qt0 = tmin / BT;
    if (a.window > 0) {
      int tmax = j1 - off + a.window;  // t <= j - off + window
      if (tmax < 0) { qt1 = 0; return; }
      int e = tmax / BT + 1;
      if (e < qt1) qt1 = e;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnArgs a) {
  omr_pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* Qt = smem;               // [d][i], pre-scaled
  float* Kt = smem + TILE_F;      // [d][j]
  float* Vs = smem + 2 * TILE_F;  // [j][d]
  float* Pt = smem + 3 * TILE_F;  // [j][i]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * BT;
  const T* q = (const T*)a.q + (long long)b * a.q_bs + h * HD;
  const T* k = (const T*)a.k + (long long)b * a.k_bs + h * HD;
  const T* v = (const T*)a.v + (long long)b * a.v_bs + h * HD;
  const MaskCtx mc = make_mask_ctx(a, b, h);
  const uint32_t dstream = a.drop.thr ? attn_drop_stream(a.drop, b * a.H + h) : 0u;

  load_tile<T, false, true>(q, a.q_rs, t0, a.Tq, nullptr, Qt, a.scale);

  float m_run[4], l_run[4], oacc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -INFINITY; l_run[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) oacc[r][c] = 0.f;
  }
  int t_last = t0 + BT - 1;
  if (t_last > a.Tq - 1) t_last = a.Tq - 1;
  int kt0, kt1;
  key_tile_range(a, t0, t_last, kt0, kt1);

  for (int kt = kt0; kt < kt1; ++kt) {
    const int j0 = kt * BT;
    __syncthreads();  // previous iteration finished reading Kt/Vs/Pt (and Qt is written on the first pass)
    load_tile<T, false, true>(k, a.k_rs, j0, a.Tk, nullptr, Kt, 1.f);
    load_tile<T, true, false>(v, a.v_rs, j0, a.Tk, Vs, nullptr, 1.f);
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
    simt_mma_4x4<LDT, LDT, HD>(Qt, Kt, ty * 4, tx * 4, s);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int t = t0 + ty * 4 + r;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[r][c] += mask_term(a, mc, t, j0 + tx * 4 + c);
        mx = fmaxf(mx, s[r][c]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[r], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[r][c] = fexp<T>(s[r][c] - m_safe);
        rs += s[r][c];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      const float alpha = (m_run[r] == -INFINITY) ? 0.f : fexp<T>(m_run[r] - m_safe);
      l_run[r] = l_run[r] * alpha + rs;
      m_run[r] = m_new;
#pragma unroll
      for (int c = 0; c < 4; ++c) oacc[r][c] *= alpha;
#pragma unroll
      for (int c = 0; c < 4; ++c) Pt[(tx * 4 + c) * LDT + ty * 4 + r] = s[r][c] * drop_factor(a, dstream, t, j0 + tx * 4 + c);
    }
    __syncthreads();
    simt_mma_4x4<LDT, LDT, BT>(Pt, Vs, ty * 4, tx * 4, oacc);
  }

  T* o = (T*)a.o_w + (long long)b * a.o_bs + h * HD;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + ty * 4 + r;
    if (t >= a.Tq) continue;
    const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
    float ov[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ov[c] = oacc[r][c] * inv;
    store4(o + (long long)t * a.o_rs + tx * 4, ov);
    if (tx == 0 && a.lse)
      a.lse[((long long)b * a.H + h) * a.Tq + t] = l_run[r] > 0.f ? m_run[r] + logf(l_run[r]) : 0.f;
  }
}

// delta[b,h,t] = sum_d dO * O ; one warp per (b,h,t)
template <typename T>
__global__ void attn_delta_kernel(AttnArgs a) {
  omr_pdl_enter();
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long long total = (long long)a.B * a.H * a.Tq;
  if (w >= total) return;
  int t = (int)(w % a.Tq);
  long long r = w / a.Tq;
  int h = (int)(r % a.H), b = (int)(r / a.H);
  const T* o = (const T*)a.o + (long long)b * a.o_bs + (long long)t * a.o_rs + h * HD;
  const T* d = (const T*)a.dout + (long long)b * a.do_bs + (long long)t * a.do_rs + h * HD;
  float s = to_f(o[lane]) * to_f(d[lane]) + to_f(o[lane + 32]) * to_f(d[lane + 32]);
  s = warp_sum(s);
  if (lane == 0) a.delta[w] = s;
}

// ---------------------------------------------------------------------------------------------
// dK, dV: one block per (key tile, h, b), loops over the query tiles that can see it.
template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_dkdv_kernel(AttnArgs a) {
  omr_pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* Kt = smem;                // [d][j]
  float* Vt = smem + TILE_F;       // [d][j]
  float* Qs = smem + 2 * TILE_F;   // [i][d]
  float* Qt = smem + 3 * TILE_F;   // [d][i]
  float* dOs = smem + 4 * TILE_F;  // [i][d]
  float* dOt = smem + 5 * TILE_F;  // [d][i]
  float* Ps = smem + 6 * TILE_F;   // [i][j]  (P, then dS)
  float* row_lse = smem + 7 * TILE_F;
  float* row_delta = row_lse + BT;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * BT;
  const T* q = (const T*)a.q + (long long)b * a.q_bs + h * HD;
  const T* k = (const T*)a.k + (long long)b * a.k_bs + h * HD;
  const T* v = (const T*)a.v + (long long)b * a.v_bs + h * HD;
  const T* dO = (const T*)a.dout + (long long)b * a.do_bs + h * HD;
  const MaskCtx mc = make_mask_ctx(a, b, h);
  const uint32_t dstream = a.drop.thr ? attn_drop_stream(a.drop, b * a.H + h) : 0u;

  load_tile<T, false, true>(k, a.k_rs, j0, a.Tk, nullptr, Kt, 1.f);
  load_tile<T, false, true>(v, a.v_rs, j0, a.Tk, nullptr, Vt, 1.f);

  float dk[4][4], dv[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) { dk[r][c] = 0.f; dv[r][c] = 0.f; }

  int j_last = j0 + BT - 1;
  if (j_last > a.Tk - 1) j_last = a.Tk - 1;
  int qt0, qt1;
  query_tile_range(a, j0, j_last, qt0, qt1);
  const long long stat_base = ((long long)b * a.H + h) * a.Tq;

  for (int qt = qt0; qt < qt1; ++qt) {
    const int t0 = qt * BT;
    __syncthreads();
    load_tile<T, true, true>(q, a.q_rs, t0, a.Tq, Qs, Qt, 1.f);
    load_tile<T, true, true>(dO, a.do_rs, t0, a.Tq, dOs, dOt, 1.f);
    if (tid < BT) {
      int t = t0 + tid;
      row_lse[tid] = t < a.Tq ? a.lse[stat_base + t] : 0.f;
      row_delta[tid] = t < a.Tq ? a.delta[stat_base + t] : 0.f;
    }
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) { s[r][c] = 0.f; dp[r][c] = 0.f; }
    simt_mma_4x4<LDT, LDT, HD>(Qt, Kt, ty * 4, tx * 4, s);     // rows i, cols j
    simt_mma_4x4<LDT, LDT, HD>(dOt, Vt, ty * 4, tx * 4, dp);   // rows i, cols j
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ty * 4 + r;
      const float l = row_lse[i], dl = row_delta[i];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float mt = mask_term(a, mc, t0 + i, j0 + tx * 4 + c);
        float p = (mt == -INFINITY) ? 0.f : fexp<T>(s[r][c] * a.scale + mt - l);
        const float df = drop_factor(a, dstream, t0 + i, j0 + tx * 4 + c);
        s[r][c] = p;
        dp[r][c] = p * (dp[r][c] * df - dl);
        Ps[i * LDT + tx * 4 + c] = p * df;
      }
    }
    __syncthreads();
    simt_mma_4x4<LDT, LDT, BT>(Ps, dOs, ty * 4, tx * 4, dv);  // dV[j][d] += sum_i P[i][j] dO[i][d]
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) Ps[(ty * 4 + r) * LDT + tx * 4 + c] = dp[r][c];
    __syncthreads();
    simt_mma_4x4<LDT, LDT, BT>(Ps, Qs, ty * 4, tx * 4, dk);   // dK[j][d] += sum_i dS[i][j] Q[i][d]
  }

  T* dkp = (T*)a.dk + (long long)b * a.dk_bs + h * HD;
  T* dvp = (T*)a.dv + (long long)b * a.dv_bs + h * HD;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + ty * 4 + r;
    if (j >= a.Tk) continue;
    float kv[4], vv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { kv[c] = dk[r][c] * a.scale; vv[c] = dv[r][c]; }
    store4(dkp + (long long)j * a.dk_rs + tx * 4, kv);
    store4(dvp + (long long)j * a.dv_rs + tx * 4, vv);
  }
}

// dQ: one block per (query tile, h, b), loops over key tiles.
template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(AttnArgs a) {
  omr_pdl_enter();
  extern __shared__ __align__(16) float smem[];
  float* Qt = smem;                // [d][i]
  float* dOt = smem + TILE_F;      // [d][i]
  float* Kt = smem + 2 * TILE_F;   // [d][j]
  float* Ks = smem + 3 * TILE_F;   // [j][d]
  float* Vt = smem + 4 * TILE_F;   // [d][j]
  float* dSt = smem + 5 * TILE_F;  // [j][i]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * BT;
  const T* q = (const T*)a.q + (long long)b * a.q_bs + h * HD;
  const T* k = (const T*)a.k + (long long)b * a.k_bs + h * HD;
  const T* v = (const T*)a.v + (long long)b * a.v_bs + h * HD;
  const T* dO = (const T*)a.dout + (long long)b * a.do_bs + h * HD;
  const MaskCtx mc = make_mask_ctx(a, b, h);
  const uint32_t dstream = a.drop.thr ? attn_drop_stream(a.drop, b * a.H + h) : 0u;
  const long long stat_base = ((long long)b * a.H + h) * a.Tq;

  load_tile<T, false, true>(q, a.q_rs, t0, a.Tq, nullptr, Qt, 1.f);
  load_tile<T, false, true>(dO, a.do_rs, t0, a.Tq, nullptr, dOt, 1.f);
  float lse_r[4], delta_r[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int t = t0 + ty * 4 + r;
    lse_r[r] = t < a.Tq ? a.lse[stat_base + t] : 0.f;
    delta_r[r] = t < a.Tq ? a.delta[stat_base + t] : 0.f;
  }
  float dq[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) dq[r][c] = 0.f;

  int t_last = t0 + BT - 1;
  if (t_last > a.Tq - 1) t_last = a.Tq - 1;
  int kt0, kt1;
  key_tile_range(a, t0, t_last, kt0, kt1);
  for (int kt = kt0; kt < kt1; ++kt) {
    const int j0 = kt * BT;
    __syncthreads();
    load_tile<T, true, true>(k, a.k_rs, j0, a.Tk, Ks, Kt, 1.f);
    load_tile<T, false, true>(v, a.v_rs, j0, a.Tk, nullptr, Vt, 1.f);
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) { s[r][c] = 0.f; dp[r][c] = 0.f; }
    simt_mma_4x4<LDT, LDT, HD>(Qt, Kt, ty * 4, tx * 4, s);
    simt_mma_4x4<LDT, LDT, HD>(dOt, Vt, ty * 4, tx * 4, dp);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ty * 4 + r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float mt = mask_term(a, mc, t0 + i, j0 + tx * 4 + c);
        float p = (mt == -INFINITY) ? 0.f : fexp<T>(s[r][c] * a.scale + mt - lse_r[r]);
        dSt[(tx * 4 + c) * LDT + i] = p * (dp[r][c] * drop_factor(a, dstream, t0 + i, j0 + tx * 4 + c) - delta_r[r]);
      }
    }
    __syncthreads();
    simt_mma_4x4<LDT, LDT, BT>(dSt, Ks, ty * 4, tx * 4, dq);  // dQ[i][d] += sum_j dS[i][j] K[j][d]
  }
  T* dqp = (T*)a.dq + (long long)b * a.dq_bs + h * HD;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + ty * 4 + r;
    if (t >= a.Tq) continue;
    float qv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) qv[c] = dq[r][c] * a.scale;
    store4(dqp + (long long)t * a.dq_rs + tx * 4, qv);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    omr_set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", bytes, cudaGetErrorString(e));
    return OMR_ERR_CUDA;
  }
  return OMR_OK;
}

bool aligned4(long long v) { return (v & 3) == 0; }

}  // namespace

int omr_attn_fwd_simt(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs,
                      long long k_rs, const void* v, long long v_bs, long long v_rs, void* o, long long o_bs,
                      long long o_rs, float* lse, const float* key_bias, int B, int H, int Tq, int Tk, int hd,
                      float scale, int causal, int window, const int* q_len, const int* kv_len, int quirk_mod,
                      cudaStream_t st) {
  OMR_REQUIRE(hd == HD, "omr_attn_fwd: head_dim must be 64 (got %d)", hd);
  OMR_REQUIRE(aligned4(q_bs) && aligned4(q_rs) && aligned4(k_bs) && aligned4(k_rs) && aligned4(v_bs) && aligned4(v_rs) &&
                  aligned4(o_bs) && aligned4(o_rs),
              "omr_attn_fwd: strides must be multiples of 4 elements");
  if (B <= 0 || H <= 0 || Tq <= 0) return OMR_OK;
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o_w = o; a.lse = lse; a.key_bias = key_bias;
  a.q_bs = q_bs; a.q_rs = q_rs; a.k_bs = k_bs; a.k_rs = k_rs; a.v_bs = v_bs; a.v_rs = v_rs; a.o_bs = o_bs; a.o_rs = o_rs;
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.scale = scale; a.causal = causal; a.window = window;
  a.q_len = q_len; a.kv_len = kv_len; a.quirk_mod = quirk_mod;
  a.drop = omr_attn_cur_dropout();
  dim3 grid((unsigned)cdiv(Tq, BT), (unsigned)H, (unsigned)B);
  size_t smem = sizeof(float) * 4 * TILE_F;
  OMR_DISPATCH_DT(dt, T, {
    static bool done = false;
    if (!done) {
      int rc = set_smem(attn_fwd_kernel<T>, smem);
      if (rc) return rc;
      done = true;
    }
    OmrLaunch(grid, 256, smem, st)(attn_fwd_kernel<T>, a);
  });
  OMR_LAUNCHED();
  return OMR_OK;
}

int omr_attn_bwd_simt(int dt, const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs,
                      long long k_rs, const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs,
                      long long o_rs, const void* dout, long long do_bs, long long do_rs, const float* lse, void* dq,
                      long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs, void* dv,
                      long long dv_bs, long long dv_rs, float* delta_ws, const float* key_bias, int B, int H, int Tq,
                      int Tk, int hd, float scale, int causal, int window, const int* q_len, const int* kv_len,
                      int quirk_mod, cudaStream_t st) {
  OMR_REQUIRE(hd == HD, "omr_attn_bwd: head_dim must be 64 (got %d)", hd);
  OMR_REQUIRE(aligned4(q_bs) && aligned4(q_rs) && aligned4(k_bs) && aligned4(k_rs) && aligned4(v_bs) && aligned4(v_rs) &&
                  aligned4(o_bs) && aligned4(o_rs) && aligned4(do_bs) && aligned4(do_rs) && aligned4(dq_bs) &&
                  aligned4(dq_rs) && aligned4(dk_bs) && aligned4(dk_rs) && aligned4(dv_bs) && aligned4(dv_rs),
              "omr_attn_bwd: strides must be multiples of 4 elements");
  if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0) return OMR_OK;
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o; a.dout = dout; a.lse = const_cast<float*>(lse); a.delta = delta_ws;
  a.dq = dq; a.dk = dk; a.dv = dv; a.key_bias = key_bias;
  a.q_bs = q_bs; a.q_rs = q_rs; a.k_bs = k_bs; a.k_rs = k_rs; a.v_bs = v_bs; a.v_rs = v_rs; a.o_bs = o_bs; a.o_rs = o_rs;
  a.do_bs = do_bs; a.do_rs = do_rs; a.dq_bs = dq_bs; a.dq_rs = dq_rs; a.dk_bs = dk_bs; a.dk_rs = dk_rs;
  a.dv_bs = dv_bs; a.dv_rs = dv_rs;
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.scale = scale; a.causal = causal; a.window = window;
  a.q_len = q_len; a.kv_len = kv_len; a.quirk_mod = quirk_mod;
  a.drop = omr_attn_cur_dropout();
  long long nw = (long long)B * H * Tq;
  size_t smem_kv = sizeof(float) * (7 * TILE_F + 2 * BT);
  size_t smem_q = sizeof(float) * 6 * TILE_F;
  dim3 grid_kv((unsigned)cdiv(Tk, BT), (unsigned)H, (unsigned)B);
  dim3 grid_q((unsigned)cdiv(Tq, BT), (unsigned)H, (unsigned)B);
  OMR_DISPATCH_DT(dt, T, {
    OmrLaunch((unsigned)cdiv(nw * 32, 256), 256, 0, st)(attn_delta_kernel<T>, a);
    omr_count_launch();
    static bool done = false;
    if (!done) {
      int rc = set_smem(attn_bwd_dkdv_kernel<T>, smem_kv);
      if (rc) return rc;
      rc = set_smem(attn_bwd_dq_kernel<T>, smem_q);
      if (rc) return rc;
      done = true;
    }
    OmrLaunch(grid_kv, 256, smem_kv, st)(attn_bwd_dkdv_kernel<T>, a);
    omr_count_launch();
    OmrLaunch(grid_q, 256, smem_q, st)(attn_bwd_dq_kernel<T>, a);
  });
  OMR_LAUNCHED();
  return OMR_OK;
}
