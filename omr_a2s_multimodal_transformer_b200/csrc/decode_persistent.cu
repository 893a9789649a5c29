// decode_persistent.cu -- the batched greedy decoder as ONE persistent cooperative kernel.
//
// The reference decodes one sample at a time, re-running the whole decoder on the growing prefix and synchronising
// with the host for every token (src/transformer/model.py:170-199, 592-617).  Here a single launch of one CTA per SM
// executes `nsteps` complete decode steps for the whole batch: embedding + 1-D PE, 8 x (self-attention over the in-HBM
// KV cache, cross-attention over the pre-projected encoder memory, FFN, three post-norm LayerNorms), the vocabulary
// classifier, first-max argmax and the EOS bookkeeping -- with grid-wide barriers between dependent phases instead of
// ~100 kernel launches per token.  A step is a weight / KV stream (~25 MB per token at batch 32), so every phase is a
// bandwidth problem spread over all SMs:
//   * projection phases: the activation rows (B <= 64, D = 256) are staged in shared memory by every CTA (LayerNorm of
//     the previous block applied on the fly), output COLUMNS are dealt round-robin to the warps of all CTAs (lane = row),
//     new K/V rows are written straight into the cache;
//   * attention phases: (batch, head, key-split) items are dealt to the CTAs; 16-byte loads, 32 keys in flight per CTA
//     step; the last CTA to finish a (batch, head) combines the split partials (split-K "last arriver" pattern);
//   * argmax: one CTA per batch row; token / finished / position state never leaves the device.
// Numerics are those of the per-kernel path (fp32 accumulation everywhere); activations between phases are fp32.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int DP_D = 256, DP_HD = 64, DP_THREADS = 512, DP_WARPS = 16, DP_KL = DP_THREADS / 8;  // DP_KL key lanes
constexpr int DP_LDX = DP_D + 1;  // padded fp32 row of the staged activations: lane = row reads are conflict-free

template <typename T>
struct LayerW {
  const T* w_in; const float* b_in;     // self-attn packed in-proj [3D, D]
  const T* w_o; const float* b_o;       // self-attn out-proj [D, D]
  const T* wc_q; const float* bc_q;     // cross-attn q rows of the packed in-proj [D, D]
  const T* wc_o; const float* bc_o;     // cross-attn out-proj
  const T* w1; const float* b1;         // FFN
  const T* w2; const float* b2;
  const float *g1, *be1, *g2, *be2, *g3, *be3;  // LayerNorm affine
  T* self_kv;                           // [B, Tmax, 2D]
  const T* cross_kv;                    // [B, S, 2D]
};

struct DPArgs {
  const void* layers;  // device array of LayerW<T>
  int L;
  const void* emb; const float* pe; const void* w_out; const float* b_out;
  int B, H, V, S, Tmax, nsteps, window;
  long long* tok; float* val; int* finished; long long* out_tokens; float* out_vals; int out_ld; int* pos;
  long long eos, pad;
  const float* mem_bias; long long mem_bias_bs;
  float ln_eps, scale;
  // scratch (fp32): s [B,D] pre-LN sums, x [B,D] residual stream, q [B,3D], a [B,D], h [B,D], logits [B,Vld]
  float *s, *x, *q, *a, *h, *logits;
  long long vld;
  float *ws_o, *ws_ml;  // attention split partials [B*H*MAXSPLIT, 64] / [.., 2]
  int* cnt;             // [B*H] arrival counters (zero)
  unsigned* bar;        // grid barrier counter (zero)
  int max_split;
  long long* timing;  // optional [16] cycle counters per phase kind (CTA 0), NULL = off
};

// 8 consecutive elements as raw registers (so that many independent 16-byte loads can be in flight per thread)
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { v = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void zero() { a = make_float4(0, 0, 0, 0); b = a; }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v, spin = 0;
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spin > (1u << 25)) __trap();  // a lost CTA becomes a launch failure, not a hung GPU
    } while ((int)(v - target) < 0);
    __threadfence();
  }
  __syncthreads();
}

// rows [B, D] fp32 from global -> smem (optionally LayerNorm(src) * gamma + beta); CTA 0 may publish the result
__device__ void stage_rows(float* xs, const float* __restrict__ src, int B, const float* gamma, const float* beta, float eps,
                           float* publish) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < B; r += DP_WARPS) {
    float v[8];
    const float* row = src + (long long)r * DP_D;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(row + k * 32 + lane);  // L2 (written by other CTAs before the barrier)
    if (gamma) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += v[k];
      const float mean = warp_sum(sum) * (1.f / DP_D);
      float var = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; var = fmaf(d, d, var); }
      const float rstd = rsqrtf(warp_sum(var) * (1.f / DP_D) + eps);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (v[k] - mean) * rstd * gamma[k * 32 + lane] + beta[k * 32 + lane];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) xs[r * DP_LDX + k * 32 + lane] = v[k];
    if (publish && blockIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) publish[(long long)r * DP_D + k * 32 + lane] = v[k];
    }
  }
  __syncthreads();
}

// out(m, n) for every column n of W [N, 256] (row-major, K contiguous): columns dealt to (CTA, warp), lane = row m
template <typename T, typename Epi>
__device__ void gemm_cols(const float* xs, float* wrow_all, const T* __restrict__ W, const float* __restrict__ bias, int N, int B,
                          Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wrow = wrow_all + warp * DP_D;
  for (int n = blockIdx.x + gridDim.x * warp; n < N; n += gridDim.x * DP_WARPS) {
    {
      float w8[8];
      const T* wr = W + (long long)n * DP_D + lane * 8;
      load4(wr, *reinterpret_cast<float(*)[4]>(w8));
      load4(wr + 4, *reinterpret_cast<float(*)[4]>(w8 + 4));
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 8; ++e) wrow[lane * 8 + e] = w8[e];
      __syncwarp();
    }
    const float bn = bias ? bias[n] : 0.f;
    for (int mb = 0; mb < B; mb += 32) {
      const int m = mb + lane;
      const float* xr = xs + (m < B ? m : 0) * DP_LDX;
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
      for (int k = 0; k < DP_D; k += 2) {
        acc0 = fmaf(xr[k], wrow[k], acc0);
        acc1 = fmaf(xr[k + 1], wrow[k + 1], acc1);
      }
      if (m < B) epi(m, n, acc0 + acc1 + bn);
    }
  }
}

// single-query attention over keys [j_lo, tk) of every (b, h): split items dealt to CTAs, last arriver combines into a[b, h*64..]
template <typename T>
__device__ void attn_phase(const DPArgs& p, float* sm, const float* __restrict__ qbuf, long long q_rs, const T* __restrict__ kv,
                           long long kv_bs, int tk, int j_lo, const float* __restrict__ kbias, long long kb_bs) {
  const int tid = threadIdx.x, c = tid & 7, g = tid >> 3;  // DP_KL key lanes x 8 dim chunks
  float* sc = sm;                 // [chunk]
  const int nkeys = tk - j_lo;
  int nsplit = (nkeys + 255) / 256;
  const int bh_total = p.B * p.H;
  const int want = (int)((2LL * gridDim.x + bh_total - 1) / bh_total);
  if (nsplit > want) nsplit = want;
  if (nsplit > p.max_split) nsplit = p.max_split;
  if (nsplit < 1) nsplit = 1;
  const int chunk = (nkeys + nsplit - 1) / nsplit;
  float* red = sm + ((chunk + 3) & ~3);  // [DP_KL][64] key-lane partial outputs
  __shared__ float s_red[33];
  __shared__ int s_last;
  const int items = bh_total * nsplit;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int bh = item / nsplit, sp = item - bh * nsplit;
    const int b = bh / p.H, h = bh - b * p.H;
    int j0 = j_lo + sp * chunk, j1 = j0 + chunk;
    if (j1 > tk) j1 = tk;
    const int n = j1 > j0 ? j1 - j0 : 0;
    const T* kp = kv + (long long)b * kv_bs + h * DP_HD;
    const T* vp = kp + DP_D;
    const float* kb = kbias ? kbias + (long long)b * kb_bs : nullptr;
    float q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = __ldcg(qbuf + (long long)b * q_rs + h * DP_HD + c * 8 + e) * p.scale;
    __syncthreads();  // previous item's readers of sc / red are done
    float mx = -INFINITY;
    constexpr int U = sizeof(T) == 2 ? 8 : 4;  // keys in flight per thread
    for (int jb = 0; jb < n; jb += U * DP_KL) {
      Raw8<T> raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = jb + u * DP_KL + g;
        if (j < n) raw[u].load(kp + (long long)(j0 + j) * (2 * DP_D) + c * 8);
        else raw[u].zero();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float kvv[8];
        raw[u].get(kvv);
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d = fmaf(q[e], kvv[e], d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        const int j = jb + u * DP_KL + g;
        if (j < n) {
          if (kb) d += kb[j0 + j];
          if (c == 0) sc[j] = d;
          mx = fmaxf(mx, d);
        }
      }
    }
    mx = block_max(mx, s_red);
    const float msafe = (mx == -INFINITY) ? 0.f : mx;
    float sum = 0.f;
    for (int j = tid; j < n; j += DP_THREADS) {
      const float pr = expf(sc[j] - msafe);
      sc[j] = pr;
      sum += pr;
    }
    sum = block_sum(sum, s_red);  // contains the __syncthreads that publishes sc[]
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int jb = 0; jb < n; jb += U * DP_KL) {
      Raw8<T> raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = jb + u * DP_KL + g;
        if (j < n) raw[u].load(vp + (long long)(j0 + j) * (2 * DP_D) + c * 8);
        else raw[u].zero();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = jb + u * DP_KL + g;
        const float pj = j < n ? sc[j] : 0.f;
        float vv[8];
        raw[u].get(vv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vv[e], acc[e]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) red[g * DP_HD + c * 8 + e] = acc[e];
    __syncthreads();
    if (tid < DP_HD) {
      float o = 0.f;
#pragma unroll
      for (int l = 0; l < DP_KL; ++l) o += red[l * DP_HD + tid];
      const long long w = (long long)bh * p.max_split + sp;
      __stcg(p.ws_o + w * DP_HD + tid, o);
      if (tid == 0) { __stcg(p.ws_ml + w * 2, mx); __stcg(p.ws_ml + w * 2 + 1, sum); }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(p.cnt + bh, 1) == nsplit - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {  // every split of (b, h) has been published: combine
      __threadfence();
      if (tid < DP_HD) {
        float M = -INFINITY;
        for (int s2 = 0; s2 < nsplit; ++s2) M = fmaxf(M, __ldcg(p.ws_ml + ((long long)bh * p.max_split + s2) * 2));
        float Lsum = 0.f, o = 0.f;
        for (int s2 = 0; s2 < nsplit; ++s2) {
          const long long w = (long long)bh * p.max_split + s2;
          const float m = __ldcg(p.ws_ml + w * 2);
          const float f = (m == -INFINITY) ? 0.f : expf(m - M);
          Lsum = fmaf(__ldcg(p.ws_ml + w * 2 + 1), f, Lsum);
          o = fmaf(__ldcg(p.ws_o + w * DP_HD + tid), f, o);
        }
        p.a[(long long)b * DP_D + h * DP_HD + tid] = Lsum > 0.f ? o / Lsum : 0.f;
        if (tid == 0) p.cnt[bh] = 0;
      }
    }
  }
}

// phase kinds for the optional timing counters
enum { PH_EMBED = 0, PH_QKV, PH_SELF, PH_OUT, PH_CQ, PH_CROSS, PH_COUT, PH_FFN1, PH_FFN2, PH_VOCAB, PH_ARGMAX };
#define DP_SYNC(kind)                                                        \
  do {                                                                       \
    grid_sync(p.bar, target);                                                \
    if (p.timing && blockIdx.x == 0 && threadIdx.x == 0) {                   \
      const long long now = clock64();                                       \
      p.timing[kind] += now - t_prev;                                        \
      t_prev = now;                                                          \
    }                                                                        \
  } while (0)

template <typename T>
__global__ void __launch_bounds__(DP_THREADS, 1) decode_persistent_kernel(DPArgs p) {
  long long t_prev = clock64();
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                   // [B][DP_LDX]
  float* wrow = xs + (size_t)p.B * DP_LDX;            // [8][256]
  float* att = wrow + DP_WARPS * DP_D;                // attention scratch: scores chunk + [32][64]
  const LayerW<T>* layers = reinterpret_cast<const LayerW<T>*>(p.layers);
  const T* emb = reinterpret_cast<const T*>(p.emb);
  const T* w_out = reinterpret_cast<const T*>(p.w_out);
  unsigned target = 0;
  const int B = p.B;
  const int pos0 = *p.pos;

  for (int step = 0; step < p.nsteps; ++step) {
    const int pos = pos0 + step;  // position of the token being consumed; keys 0..pos are visible
    if (pos >= p.Tmax) break;
    // ---- embedding + PE -> x (grid-strided), residual stream in fp32 ----
    for (int i = blockIdx.x * DP_THREADS + threadIdx.x; i < B * DP_D; i += gridDim.x * DP_THREADS) {
      const int b = i / DP_D, d = i - b * DP_D;
      p.x[i] = to_f(emb[__ldcg(p.tok + b) * DP_D + d]) + p.pe[(long long)pos * DP_D + d];  // tok: written by another CTA
    }
    DP_SYNC(PH_EMBED);
    for (int l = 0; l < p.L; ++l) {
      const LayerW<T>& W = layers[l];
      // P1: x (= LN3 of the previous layer's sum, or the embedding) -> q | k | v ; k, v go straight into the cache
      if (l == 0) stage_rows(xs, p.x, B, nullptr, nullptr, 0.f, nullptr);
      else stage_rows(xs, p.s, B, layers[l - 1].g3, layers[l - 1].be3, p.ln_eps, p.x);
      {
        T* cache = W.self_kv;
        float* qb = p.q;
        const int Tmax = p.Tmax;
        gemm_cols<T>(xs, wrow, W.w_in, W.b_in, 3 * DP_D, B, [=](int m, int n, float v) {
          if (n < DP_D) qb[(long long)m * DP_D + n] = v;
          else cache[((long long)m * Tmax + pos) * (2 * DP_D) + (n - DP_D)] = from_f<T>(v);
        });
      }
      DP_SYNC(PH_QKV);
      // P2: causal / windowed self-attention over the cache
      {
        int j_lo = 0;
        if (p.window > 0 && pos - p.window > 0) j_lo = pos - p.window;
        attn_phase<T>(p, att, p.q, DP_D, W.self_kv, (long long)p.Tmax * 2 * DP_D, pos + 1, j_lo, nullptr, 0);
      }
      DP_SYNC(PH_SELF);
      // P3: s = x + out_proj(a)
      stage_rows(xs, p.a, B, nullptr, nullptr, 0.f, nullptr);
      {
        float* sb = p.s; const float* xb = p.x;
        gemm_cols<T>(xs, wrow, W.w_o, W.b_o, DP_D, B, [=](int m, int n, float v) {
          sb[(long long)m * DP_D + n] = v + __ldcg(xb + (long long)m * DP_D + n);
        });
      }
      DP_SYNC(PH_OUT);
      // P4: x1 = LN1(s) -> cross query
      stage_rows(xs, p.s, B, W.g1, W.be1, p.ln_eps, p.x);
      {
        float* qb = p.q;
        gemm_cols<T>(xs, wrow, W.wc_q, W.bc_q, DP_D, B, [=](int m, int n, float v) { qb[(long long)m * DP_D + n] = v; });
      }
      DP_SYNC(PH_CQ);
      // P5: cross-attention over the projected encoder memory
      attn_phase<T>(p, att, p.q, DP_D, W.cross_kv, (long long)p.S * 2 * DP_D, p.S, 0, p.mem_bias, p.mem_bias_bs);
      DP_SYNC(PH_CROSS);
      // P6: s = x1 + cross out_proj(a)
      stage_rows(xs, p.a, B, nullptr, nullptr, 0.f, nullptr);
      {
        float* sb = p.s; const float* xb = p.x;
        gemm_cols<T>(xs, wrow, W.wc_o, W.bc_o, DP_D, B, [=](int m, int n, float v) {
          sb[(long long)m * DP_D + n] = v + __ldcg(xb + (long long)m * DP_D + n);
        });
      }
      DP_SYNC(PH_COUT);
      // P7: x2 = LN2(s) -> h = relu(W1 x2 + b1)
      stage_rows(xs, p.s, B, W.g2, W.be2, p.ln_eps, p.x);
      {
        float* hb = p.h;
        gemm_cols<T>(xs, wrow, W.w1, W.b1, DP_D, B, [=](int m, int n, float v) { hb[(long long)m * DP_D + n] = fmaxf(v, 0.f); });
      }
      DP_SYNC(PH_FFN1);
      // P8: s = x2 + W2 h + b2
      stage_rows(xs, p.h, B, nullptr, nullptr, 0.f, nullptr);
      {
        float* sb = p.s; const float* xb = p.x;
        gemm_cols<T>(xs, wrow, W.w2, W.b2, DP_D, B, [=](int m, int n, float v) {
          sb[(long long)m * DP_D + n] = v + __ldcg(xb + (long long)m * DP_D + n);
        });
      }
      DP_SYNC(PH_FFN2);
    }
    // ---- classifier on LN3(s) of the last layer ----
    stage_rows(xs, p.s, B, layers[p.L - 1].g3, layers[p.L - 1].be3, p.ln_eps, nullptr);
    {
      float* lg = p.logits; const long long vld = p.vld;
      gemm_cols<T>(xs, wrow, w_out, p.b_out, p.V, B, [=](int m, int n, float v) { lg[(long long)m * vld + n] = v; });
    }
    DP_SYNC(PH_VOCAB);
    // ---- first-max argmax + EOS bookkeeping: one CTA per batch row ----
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      __shared__ float sv[DP_THREADS];
      __shared__ int si[DP_THREADS];
      const float* row = p.logits + (long long)b * p.vld;
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = threadIdx.x; i < p.V; i += DP_THREADS) {
        // the per-kernel path rounds logits to the storage type before the argmax; do the same so that ties resolve alike
        const float v = to_f(from_f<T>(__ldcg(row + i)));
        if (v > best) { best = v; bi = i; }
      }
      sv[threadIdx.x] = best; si[threadIdx.x] = bi;
      __syncthreads();
      for (int s2 = DP_THREADS / 2; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) {
          const float ov = sv[threadIdx.x + s2];
          const int oi = si[threadIdx.x + s2];
          if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        long long t = si[0];
        float v = sv[0];
        if (__ldcg(p.finished + b)) { t = p.pad; v = 0.f; }
        else if (t == p.eos) p.finished[b] = 1;
        p.tok[b] = t;
        p.val[b] = v;
        if (pos < p.out_ld) {
          p.out_tokens[(long long)b * p.out_ld + pos] = t;
          p.out_vals[(long long)b * p.out_ld + pos] = v;
        }
      }
      __syncthreads();
    }
    DP_SYNC(PH_ARGMAX);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int done = p.nsteps;
    if (pos0 + done > p.Tmax) done = p.Tmax - pos0;
    *p.pos = pos0 + (done > 0 ? done : 0);
  }
}

}  // namespace

extern "C" long long omr_decode_persistent_scratch_floats(int B, int H, int D, int V) {
  auto r4 = [](long long n) { return (n + 3) / 4 * 4; };
  const long long vld = (V + 3) / 4 * 4;
  return 5 * r4((long long)B * D) + r4((long long)B * vld) + r4((long long)B * H * 8 * DP_HD) + r4((long long)B * H * 8 * 2) +
         r4((long long)B * H) + 64;
}

extern "C" int omr_decode_persistent(int dt, const void* layers_dev, int L, const void* emb, const float* pe, const void* w_out,
                                     const float* b_out, int B, int H, int D, int V, int S, int Tmax, int nsteps, int window,
                                     long long* tok, float* val, int* finished, long long* out_tokens, float* out_vals, int out_ld,
                                     int* pos, long long eos, long long pad, const float* mem_bias, long long mem_bias_bs,
                                     float ln_eps, float* scratch, long long scratch_floats, long long* timing,
                                     omr_stream_t stream) {
  OMR_REQUIRE(D == DP_D && H * DP_HD == D, "omr_decode_persistent: d_model must be 256 with 64-wide heads");
  OMR_REQUIRE(B >= 1 && B <= 64, "omr_decode_persistent: batch must be in [1, 64] (got %d)", B);
  OMR_REQUIRE(L >= 1 && V >= 1 && S >= 1 && Tmax >= 1 && nsteps >= 0, "omr_decode_persistent: bad sizes");
  if (nsteps == 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  int dev = 0, sms = 0;
  OMR_CUDA(cudaGetDevice(&dev));
  OMR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DPArgs p{};
  p.layers = layers_dev; p.L = L; p.emb = emb; p.pe = pe; p.w_out = w_out; p.b_out = b_out;
  p.B = B; p.H = H; p.V = V; p.S = S; p.Tmax = Tmax; p.nsteps = nsteps; p.window = window;
  p.tok = tok; p.val = val; p.finished = finished; p.out_tokens = out_tokens; p.out_vals = out_vals; p.out_ld = out_ld; p.pos = pos;
  p.eos = eos; p.pad = pad; p.mem_bias = mem_bias; p.mem_bias_bs = mem_bias_bs; p.ln_eps = ln_eps; p.scale = 0.125f;
  p.max_split = 8;
  p.timing = timing;
  // scratch carve-up (floats): s, x, a, h [B*D] each, q [B*D], logits [B*vld], ws_o [B*H*8*64], ws_ml [B*H*8*2], cnt [B*H], bar [4]
  p.vld = (V + 3) / 4 * 4;
  long long off = 0;
  auto take = [&](long long n) { float* r = scratch + off; off += (n + 3) / 4 * 4; return r; };
  p.s = take((long long)B * D); p.x = take((long long)B * D); p.a = take((long long)B * D); p.h = take((long long)B * D);
  p.q = take((long long)B * D); p.logits = take((long long)B * p.vld);
  p.ws_o = take((long long)B * H * p.max_split * DP_HD); p.ws_ml = take((long long)B * H * p.max_split * 2);
  p.cnt = reinterpret_cast<int*>(take((long long)B * H));
  p.bar = reinterpret_cast<unsigned*>(take(4));
  OMR_REQUIRE(off <= scratch_floats, "omr_decode_persistent: scratch too small (%lld < %lld floats)", scratch_floats, off);
  OMR_CUDA(cudaMemsetAsync(p.cnt, 0, sizeof(int) * (size_t)B * H, st));
  OMR_CUDA(cudaMemsetAsync(p.bar, 0, 16, st));
  const int max_keys = S > Tmax ? S : Tmax;
  const size_t smem = sizeof(float) * ((size_t)B * DP_LDX + DP_WARPS * DP_D + (size_t)((max_keys + 3) & ~3) + DP_KL * DP_HD + 16);
  OMR_REQUIRE(smem <= 200 * 1024, "omr_decode_persistent: memory / sequence too long for the score buffer (%zu B)", smem);
  void* args[] = {&p};
  const void* fn = dt == OMR_BF16 ? (const void*)decode_persistent_kernel<bf16> : (const void*)decode_persistent_kernel<float>;
  OMR_REQUIRE(dt == OMR_BF16 || dt == OMR_F32, "omr_decode_persistent: bad dtype");
  static bool cfg[2] = {false, false};
  if (!cfg[dt]) {
    OMR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg[dt] = true;
  }
  OMR_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)sms), dim3(DP_THREADS), args, smem, st));
  omr_count_launch();
  return OMR_OK;
}
