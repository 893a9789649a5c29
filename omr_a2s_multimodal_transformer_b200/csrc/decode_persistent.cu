// decode_persistent.cu -- the batched greedy decoder as ONE persistent kernel: a 4-CTA thread-block cluster per sample.
//
// The reference decodes one sample at a time, re-running the whole decoder on the growing prefix and synchronising
// with the host for every token (src/transformer/model.py:170-199, 592-617).  Greedy decoding has no coupling between
// samples (attention is per sequence, LayerNorm per token), so here every sample of the batch is owned by one cluster
// of 4 CTAs (4 x 512 threads on 4 neighbouring SMs; batch 32 -> 128 of the 148 SMs) that runs ALL decode steps of that
// sample inside a single launch -- embedding + 1-D PE, 8 x (KV-cached self-attention, cross-attention over the
// pre-projected encoder memory, FFN, three post-norm LayerNorms), vocabulary classifier, first-max argmax, EOS -- and
// synchronises only with the hardware cluster barrier (barrier.cluster, ~0.2 us) between dependent phases: no grid-wide
// barrier, no host round trip, no kernel launch per token.  A decode step is a weight / KV stream:
//   * attention phases: CTA r of the cluster owns head r (4 heads): it streams that head's K/V rows (16-byte loads, 8
//     keys in flight per thread, 64 key lanes x 8 dim chunks), softmax in shared memory, no split-K combine needed;
//   * projection phases (GEMV: 256-vector x [N,256] weight): the N output columns are dealt to the 64 warps of the
//     cluster, every lane holds 8 elements of the input vector, one 16-byte weight load per lane and column, shuffle
//     reduction; LayerNorm of the previous block is recomputed by every warp on the fly (256 values);
//   * vectors travel between the CTAs of a cluster through a small fp32 scratch in L2; barrier.cluster (release /
//     acquire) orders them; the new K/V rows go straight into the in-HBM cache.
// Numerics are those of the per-kernel path (fp32 accumulation; logits rounded to the storage type before the argmax).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int DP_D = 256, DP_HD = 64, DP_H = 4, DP_THREADS = 512, DP_WARPS = 16, DP_KL = DP_THREADS / 8;
constexpr int DP_CL = 4;                      // CTAs per cluster = heads
constexpr int DP_CW = DP_CL * DP_WARPS;       // warps per cluster
constexpr int DP_SCR = 6 * DP_D + 16;         // fp32 scratch per sample: x, s, q, a, h (256 each), argmax candidates

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
  int B, V, S, Tmax, nsteps, window;
  long long* tok; float* val; int* finished; long long* out_tokens; float* out_vals; int out_ld; int* pos;
  long long eos, pad;
  const float* mem_bias; long long mem_bias_bs;
  float ln_eps, scale;
  float* scratch;      // [B][DP_SCR]
  long long* timing;   // optional [16] cycle counters per phase kind (cluster 0, rank 0), NULL = off
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

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}

// lane-distributed 256-vector: lane holds elements lane*8 .. lane*8+7.  Optional LayerNorm (every warp redundantly).
__device__ __forceinline__ void load_vec(const float* __restrict__ src, float (&v)[8], const float* gamma, const float* beta, float eps) {
  const int lane = threadIdx.x & 31;
  const float4 a = __ldcg(reinterpret_cast<const float4*>(src + lane * 8));
  const float4 b = __ldcg(reinterpret_cast<const float4*>(src + lane * 8 + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
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
    for (int k = 0; k < 8; ++k) v[k] = (v[k] - mean) * rstd * gamma[lane * 8 + k] + beta[lane * 8 + k];
  }
}
__device__ __forceinline__ void store_vec(float* dst, const float (&v)[8]) {
  const int lane = threadIdx.x & 31;
  *reinterpret_cast<float4*>(dst + lane * 8) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(dst + lane * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// out(n) = <x, W[n, :]> + bias[n] for the columns of this warp (cluster-wide warp id wg of DP_CW), 4 columns in flight
template <typename T, typename Epi>
__device__ __forceinline__ void gemv_cols(const float (&x)[8], const T* __restrict__ W, const float* __restrict__ bias, int N, int wg,
                                          Epi epi) {
  const int lane = threadIdx.x & 31;
  for (int n0 = wg; n0 < N; n0 += 4 * DP_CW) {
    Raw8<T> raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = n0 + u * DP_CW;
      if (n < N) raw[u].load(W + (long long)n * DP_D + lane * 8);
      else raw[u].zero();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = n0 + u * DP_CW;
      float w[8];
      raw[u].get(w);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(x[e], w[e], d);
      d = warp_sum(d);
      if (n < N && lane == 0) epi(n, d + (bias ? bias[n] : 0.f));
    }
  }
}

// single-query attention of ONE (sample, head) by the whole CTA: keys [j_lo, tk); result out[0..63] (global, fp32)
template <typename T>
__device__ void attn_head(float* sm, const float* __restrict__ q, const T* __restrict__ kp, int tk, int j_lo,
                          const float* __restrict__ kb, float scale, float* __restrict__ out) {
  const int tid = threadIdx.x, c = tid & 7, g = tid >> 3;  // DP_KL key lanes x 8 dim chunks
  __shared__ float s_red[33];
  float* sc = sm;                                // [n]
  const int n = tk - j_lo;
  float* red = sm + ((n + 3) & ~3);              // [DP_KL][64]
  const T* vp = kp + DP_D;
  float qv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) qv[e] = __ldcg(q + c * 8 + e) * scale;
  constexpr int U = sizeof(T) == 2 ? 8 : 4;  // keys in flight per thread
  float mx = -INFINITY;
  for (int jb = 0; jb < n; jb += U * DP_KL) {
    Raw8<T> raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * DP_KL + g;
      if (j < n) raw[u].load(kp + (long long)(j_lo + j) * (2 * DP_D) + c * 8);
      else raw[u].zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float kvv[8];
      raw[u].get(kvv);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(qv[e], kvv[e], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      const int j = jb + u * DP_KL + g;
      if (j < n) {
        if (kb) d += kb[j_lo + j];
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
      if (j < n) raw[u].load(vp + (long long)(j_lo + j) * (2 * DP_D) + c * 8);
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
#pragma unroll
  for (int e = 0; e < 8; ++e) red[g * DP_HD + c * 8 + e] = acc[e];
  __syncthreads();
  if (tid < DP_HD) {
    float o = 0.f;
#pragma unroll 8
    for (int l = 0; l < DP_KL; ++l) o += red[l * DP_HD + tid];
    out[tid] = sum > 0.f ? o / sum : 0.f;
  }
  __syncthreads();  // sc / red may be reused by the next call
}

enum { PH_EMBED = 0, PH_QKV, PH_SELF, PH_OUT, PH_CQ, PH_CROSS, PH_COUT, PH_FFN1, PH_FFN2, PH_VOCAB, PH_ARGMAX };
#define DP_SYNC(kind)                                   \
  do {                                                  \
    cluster_sync_all();                                 \
    if (timed) {                                        \
      const long long now = clock64();                  \
      p.timing[kind] += now - t_prev;                   \
      t_prev = now;                                     \
    }                                                   \
  } while (0)

template <typename T>
__global__ void __launch_bounds__(DP_THREADS, 1) decode_persistent_kernel(DPArgs p) {
  extern __shared__ __align__(16) float smem[];  // attention scores + key-lane partials
  __shared__ float cand_v[DP_WARPS];
  __shared__ int cand_i[DP_WARPS];
  const LayerW<T>* layers = reinterpret_cast<const LayerW<T>*>(p.layers);
  const T* emb = reinterpret_cast<const T*>(p.emb);
  const T* w_out = reinterpret_cast<const T*>(p.w_out);
  const int rank = (int)cluster_rank();          // = head owned by this CTA
  const int b = (int)cluster_id_x();             // = sample owned by this cluster
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = rank * DP_WARPS + warp;         // cluster-wide warp id
  const bool timed = p.timing && b == 0 && rank == 0 && threadIdx.x == 0;
  long long t_prev = clock64();
  float* scr = p.scratch + (long long)b * DP_SCR;
  float *xv = scr, *sv = scr + DP_D, *qv = scr + 2 * DP_D, *av = scr + 3 * DP_D, *hv = scr + 4 * DP_D;
  float* cand = scr + 5 * DP_D;                  // [4][2] per-CTA argmax candidates
  const int pos0 = *p.pos;
  const float* kbias = p.mem_bias ? p.mem_bias + (long long)b * p.mem_bias_bs : nullptr;
  long long tok = p.tok[b];
  bool fin = p.finished[b] != 0;

  for (int step = 0; step < p.nsteps && !fin; ++step) {
    const int pos = pos0 + step;  // position of the token being consumed; keys 0..pos are visible
    if (pos >= p.Tmax) break;
    // ---- x = emb[tok] + pe[pos] : written once (rank 0, warp 0), fp32 residual stream ----
    if (rank == 0 && warp == 0) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = to_f(emb[tok * DP_D + lane * 8 + k]) + p.pe[(long long)pos * DP_D + lane * 8 + k];
      store_vec(xv, v);
    }
    DP_SYNC(PH_EMBED);
    for (int l = 0; l < p.L; ++l) {
      const LayerW<T>& W = layers[l];
      float x[8];
      // P1: q | k | v of x (= the embedding, or LN3 of the previous layer's sum); k, v go straight into the cache
      if (l == 0) load_vec(xv, x, nullptr, nullptr, 0.f);
      else load_vec(sv, x, layers[l - 1].g3, layers[l - 1].be3, p.ln_eps);
      {
        T* crow = W.self_kv + ((long long)b * p.Tmax + pos) * (2 * DP_D);
        gemv_cols<T>(x, W.w_in, W.b_in, 3 * DP_D, wg, [&](int n, float v) {
          if (n < DP_D) qv[n] = v;
          else crow[n - DP_D] = from_f<T>(v);
        });
      }
      if (l > 0 && wg == 0) store_vec(xv, x);  // publish x = LN3(s): nobody reads xv in P1, P3 reads it two barriers later
      DP_SYNC(PH_QKV);
      // P2: causal / windowed self-attention of head `rank` over the cache
      {
        int j_lo = 0;
        if (p.window > 0 && pos - p.window > 0) j_lo = pos - p.window;
        attn_head<T>(smem, qv + rank * DP_HD, W.self_kv + (long long)b * p.Tmax * 2 * DP_D + rank * DP_HD, pos + 1, j_lo, nullptr,
                     p.scale, av + rank * DP_HD);
      }
      DP_SYNC(PH_SELF);
      // P3: s = x + out_proj(a)
      load_vec(av, x, nullptr, nullptr, 0.f);
      gemv_cols<T>(x, W.w_o, W.b_o, DP_D, wg, [&](int n, float v) { sv[n] = v + __ldcg(xv + n); });
      DP_SYNC(PH_OUT);
      // P4: x1 = LN1(s) -> cross query
      load_vec(sv, x, W.g1, W.be1, p.ln_eps);
      gemv_cols<T>(x, W.wc_q, W.bc_q, DP_D, wg, [&](int n, float v) { qv[n] = v; });
      if (wg == 0) store_vec(xv, x);  // x1: read again in P6 (two barriers later); P4 reads only sv
      DP_SYNC(PH_CQ);
      // P5: cross-attention of head `rank` over the projected encoder memory
      attn_head<T>(smem, qv + rank * DP_HD, W.cross_kv + (long long)b * p.S * 2 * DP_D + rank * DP_HD, p.S, 0, kbias, p.scale,
                   av + rank * DP_HD);
      DP_SYNC(PH_CROSS);
      // P6: s = x1 + cross out_proj(a)
      load_vec(av, x, nullptr, nullptr, 0.f);
      gemv_cols<T>(x, W.wc_o, W.bc_o, DP_D, wg, [&](int n, float v) { sv[n] = v + __ldcg(xv + n); });
      DP_SYNC(PH_COUT);
      // P7: x2 = LN2(s) -> h = relu(W1 x2 + b1)
      load_vec(sv, x, W.g2, W.be2, p.ln_eps);
      gemv_cols<T>(x, W.w1, W.b1, DP_D, wg, [&](int n, float v) { hv[n] = fmaxf(v, 0.f); });
      if (wg == 0) store_vec(xv, x);  // x2
      DP_SYNC(PH_FFN1);
      // P8: s = x2 + W2 h + b2
      load_vec(hv, x, nullptr, nullptr, 0.f);
      gemv_cols<T>(x, W.w2, W.b2, DP_D, wg, [&](int n, float v) { sv[n] = v + __ldcg(xv + n); });
      DP_SYNC(PH_FFN2);
    }
    // ---- classifier on LN3(s) of the last layer, fused with the first-max argmax ----
    {
      float x[8];
      load_vec(sv, x, layers[p.L - 1].g3, layers[p.L - 1].be3, p.ln_eps);
      float best = -INFINITY;
      int bi = 0x7fffffff;
      gemv_cols<T>(x, w_out, p.b_out, p.V, wg, [&](int n, float v) {
        // the per-kernel path rounds logits to the storage type before the argmax; do the same so that ties resolve alike
        const float r = to_f(from_f<T>(v));
        if (r > best || (r == best && n < bi)) { best = r; bi = n; }
      });
      if (lane == 0) { cand_v[warp] = best; cand_i[warp] = bi; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < DP_WARPS; ++w)
          if (cand_v[w] > best || (cand_v[w] == best && cand_i[w] < bi)) { best = cand_v[w]; bi = cand_i[w]; }
        cand[rank * 2] = best;
        cand[rank * 2 + 1] = __int_as_float(bi);
      }
    }
    DP_SYNC(PH_VOCAB);
    // ---- every thread resolves the 4 CTA candidates identically; rank 0 publishes ----
    {
      float best = __ldcg(cand);
      int bi = __float_as_int(__ldcg(cand + 1));
#pragma unroll
      for (int r = 1; r < DP_CL; ++r) {
        const float v = __ldcg(cand + 2 * r);
        const int i = __float_as_int(__ldcg(cand + 2 * r + 1));
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
      }
      tok = bi;
      if (tok == p.eos) fin = true;
      if (rank == 0 && threadIdx.x == 0) {
        if (pos < p.out_ld) {
          p.out_tokens[(long long)b * p.out_ld + pos] = tok;
          p.out_vals[(long long)b * p.out_ld + pos] = best;
        }
        p.tok[b] = tok;
        p.val[b] = best;
        if (fin) p.finished[b] = 1;
      }
    }
    DP_SYNC(PH_ARGMAX);  // candidates are consumed before the next step overwrites them
  }
  if (b == 0 && rank == 0 && threadIdx.x == 0) {
    int done = p.nsteps;
    if (pos0 + done > p.Tmax) done = p.Tmax - pos0;
    *p.pos = pos0 + (done > 0 ? done : 0);
  }
}

}  // namespace

extern "C" long long omr_decode_persistent_scratch_floats(int B, int H, int D, int V) {
  (void)H; (void)D; (void)V;
  return (long long)B * DP_SCR + 64;
}

extern "C" int omr_decode_persistent(int dt, const void* layers_dev, int L, const void* emb, const float* pe, const void* w_out,
                                     const float* b_out, int B, int H, int D, int V, int S, int Tmax, int nsteps, int window,
                                     long long* tok, float* val, int* finished, long long* out_tokens, float* out_vals, int out_ld,
                                     int* pos, long long eos, long long pad, const float* mem_bias, long long mem_bias_bs,
                                     float ln_eps, float* scratch, long long scratch_floats, long long* timing,
                                     omr_stream_t stream) {
  OMR_REQUIRE(D == DP_D && H == DP_H, "omr_decode_persistent: d_model must be 256 with 4 heads of 64");
  OMR_REQUIRE(B >= 1, "omr_decode_persistent: empty batch");
  OMR_REQUIRE(L >= 1 && V >= 1 && S >= 1 && Tmax >= 1 && nsteps >= 0, "omr_decode_persistent: bad sizes");
  OMR_REQUIRE(dt == OMR_BF16 || dt == OMR_F32, "omr_decode_persistent: bad dtype");
  OMR_REQUIRE(scratch_floats >= (long long)B * DP_SCR, "omr_decode_persistent: scratch too small");
  if (nsteps == 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  DPArgs p{};
  p.layers = layers_dev; p.L = L; p.emb = emb; p.pe = pe; p.w_out = w_out; p.b_out = b_out;
  p.B = B; p.V = V; p.S = S; p.Tmax = Tmax; p.nsteps = nsteps; p.window = window;
  p.tok = tok; p.val = val; p.finished = finished; p.out_tokens = out_tokens; p.out_vals = out_vals; p.out_ld = out_ld; p.pos = pos;
  p.eos = eos; p.pad = pad; p.mem_bias = mem_bias; p.mem_bias_bs = mem_bias_bs; p.ln_eps = ln_eps; p.scale = 0.125f;
  p.scratch = scratch; p.timing = timing;
  const int max_keys = S > Tmax ? S : Tmax;
  const size_t smem = sizeof(float) * ((size_t)((max_keys + 3) & ~3) + DP_KL * DP_HD + 16);
  OMR_REQUIRE(smem <= 200 * 1024, "omr_decode_persistent: memory / sequence too long for the score buffer (%zu B)", smem);
  static bool cfg[2] = {false, false};
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)(B * DP_CL));
  lc.blockDim = dim3(DP_THREADS);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = DP_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  if (dt == OMR_BF16) {
    if (!cfg[1]) { OMR_CUDA(cudaFuncSetAttribute(decode_persistent_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg[1] = true; }
    OMR_CUDA(cudaLaunchKernelEx(&lc, decode_persistent_kernel<bf16>, p));
  } else {
    if (!cfg[0]) { OMR_CUDA(cudaFuncSetAttribute(decode_persistent_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg[0] = true; }
    OMR_CUDA(cudaLaunchKernelEx(&lc, decode_persistent_kernel<float>, p));
  }
  omr_count_launch();
  return OMR_OK;
}
