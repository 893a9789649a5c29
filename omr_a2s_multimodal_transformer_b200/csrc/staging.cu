// staging.cu -- the steps either side of the model:
//   * batch collation on the device: ragged samples -> padded [B,1,Hmax,Wmax] + frame counts, ragged transcripts ->
//     y_in / y_out (reference src/data/preprocessing.py:55-144, src/data/ar_dataset.py:439-442);
//   * token-level late fusion of two decoders stepped in lock-step: softmax of both logit rows, alpha-mix, first-max
//     argmax with EOS bookkeeping (reference src/multimodal/weighted_multimodal/test.py:47-70);
//   * Levenshtein distance of token-id sequences and the Sym-ER / Seq-ER sums (reference src/utils/metrics.py:52-88).
// All HBM / latency-bound integer and elementwise work: coalesced vector accesses, one CTA per independent sequence pair.
#include "common.cuh"

namespace {

// ---- padded image / spectrogram batch ---------------------------------------------------------------------
// dst[b, 0, y, x] = src_b[y, x] inside the sample, pad_value outside; 4 output pixels per thread (16-byte stores;
// the ragged source rows are read scalar because their starts are not 16-byte aligned in general).
__global__ void __launch_bounds__(256) pad_collate_kernel(const float* __restrict__ flat, const long long* __restrict__ offs,
                                                          const int* __restrict__ hs, const int* __restrict__ ws,
                                                          float* __restrict__ dst, int B, int Hmax, int Wmax, int W4,
                                                          float pad_value, int* __restrict__ n_frames, int red_h,
                                                          int red_w) {
  omr_pdl_enter();
  const long long total = (long long)B * Hmax * W4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % W4);
    const long long r = i / W4;
    const int y = (int)(r % Hmax), b = (int)(r / Hmax);
    const int h = hs[b], w = ws[b];
    const float* src = flat + offs[b] + (long long)y * w;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x4 * 4 + k;
      v[k] = (y < h && x < w) ? __ldg(src + x) : pad_value;
    }
    float* d = dst + ((long long)b * Hmax + y) * Wmax + (long long)x4 * 4;
    if (x4 * 4 + 3 < Wmax && (Wmax & 3) == 0) {
      *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x4 * 4 + k < Wmax) d[k] = v[k];
    }
    if (n_frames && y == 0 && x4 == 0) n_frames[b] = ((h + red_h - 1) / red_h) * ((w + red_w - 1) / red_w);
  }
}

// ---- padded transcripts -----------------------------------------------------------------------------------
// y_in[b, t] = tok_b[t] (t < len_b - 1), y_out[b, t] = tok_b[t + 1] (t < len_b - 1), pad elsewhere.
__global__ void __launch_bounds__(256) pad_transcripts_kernel(const long long* __restrict__ flat,
                                                              const long long* __restrict__ offs, int B, int T,
                                                              long long* __restrict__ y_in, long long* __restrict__ y_out,
                                                              long long pad_id) {
  omr_pdl_enter();
  const long long total = (long long)B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T), b = (int)(i / T);
    const long long o = offs[b];
    const long long len = offs[b + 1] - o;
    const bool in = t < len - 1;
    y_in[i] = in ? flat[o + t] : pad_id;
    y_out[i] = in ? flat[o + t + 1] : pad_id;
  }
}

// ---- softmax-mix + argmax step -----------------------------------------------------------------------------
struct MaxSum { float m, s; };
__device__ __forceinline__ MaxSum ms_merge(MaxSum a, MaxSum b) {
  if (b.m == -INFINITY) return a;
  if (a.m == -INFINITY) return b;
  float m = fmaxf(a.m, b.m);
  return MaxSum{m, a.s * expf(a.m - m) + b.s * expf(b.m - m)};
}

template <typename T>
__global__ void __launch_bounds__(256) mix_argmax_step_kernel(const T* __restrict__ la, long long lda,
                                                              const T* __restrict__ lb, long long ldb, int V, float alpha,
                                                              long long* __restrict__ tok, float* __restrict__ val,
                                                              int* __restrict__ finished, long long eos_id,
                                                              long long pad_id, long long* __restrict__ out_tokens,
                                                              float* __restrict__ out_vals, int out_ld, int step,
                                                              const int* __restrict__ step_dev) {
  omr_pdl_enter();
  if (step_dev) step = *step_dev;
  __shared__ float sm_a[256], ss_a[256], sm_b[256], ss_b[256];
  __shared__ float sv[256];
  __shared__ int si[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const T* xa = la + (long long)b * lda;
  const T* xb = lb + (long long)b * ldb;
  // pass 1: online (max, sum of exp) of both rows
  MaxSum a{-INFINITY, 0.f}, c{-INFINITY, 0.f};
  for (int i = tid; i < V; i += 256) {
    a = ms_merge(a, MaxSum{to_f(xa[i]), 1.f});
    c = ms_merge(c, MaxSum{to_f(xb[i]), 1.f});
  }
  sm_a[tid] = a.m; ss_a[tid] = a.s; sm_b[tid] = c.m; ss_b[tid] = c.s;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
      MaxSum r = ms_merge(MaxSum{sm_a[tid], ss_a[tid]}, MaxSum{sm_a[tid + s], ss_a[tid + s]});
      sm_a[tid] = r.m; ss_a[tid] = r.s;
      r = ms_merge(MaxSum{sm_b[tid], ss_b[tid]}, MaxSum{sm_b[tid + s], ss_b[tid + s]});
      sm_b[tid] = r.m; ss_b[tid] = r.s;
    }
    __syncthreads();
  }
  const float ma = sm_a[0], mb = sm_b[0];
  const float wa = alpha / ss_a[0], wb = (1.f - alpha) / ss_b[0];
  // pass 2: mixed probability, first-max argmax
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < V; i += 256) {
    float p = wa * expf(to_f(xa[i]) - ma) + wb * expf(to_f(xb[i]) - mb);
    if (p > best) { best = p; bi = i; }
  }
  sv[tid] = best; si[tid] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
      float ov = sv[tid + s];
      int oi = si[tid + s];
      if (ov > sv[tid] || (ov == sv[tid] && oi < si[tid])) { sv[tid] = ov; si[tid] = oi; }
    }
    __syncthreads();
  }
  if (tid == 0) {
    long long t = si[0];
    float v = sv[0];
    if (finished) {
      if (finished[b]) { t = pad_id; v = 0.f; }
      else if (t == eos_id) finished[b] = 1;
    }
    tok[b] = t;
    if (val) val[b] = v;
    if (out_tokens && step < out_ld) out_tokens[(long long)b * out_ld + step] = t;
    if (out_vals && step < out_ld) out_vals[(long long)b * out_ld + step] = v;
  }
}

// ---- Levenshtein -------------------------------------------------------------------------------------------
// One CTA per (truth, hypothesis) pair; the DP table D[i][j] (i over truth, j over hypothesis) is swept by
// anti-diagonals: every cell of diagonal d = i + j depends only on diagonals d-1 and d-2, so a diagonal is one
// parallel step.  Three rotating diagonals live in shared memory, indexed by i.  Unit costs, as the reference.
__global__ void __launch_bounds__(256) levenshtein_kernel(const long long* __restrict__ ta, const long long* __restrict__ oa,
                                                          const long long* __restrict__ tb, const long long* __restrict__ ob,
                                                          int P, int* __restrict__ ed_out,
                                                          unsigned long long* __restrict__ sums, int max_len) {
  omr_pdl_enter();
  extern __shared__ int sm[];
  const int p = blockIdx.x;
  if (p >= P) return;
  const long long* a = ta + oa[p];
  const long long* b = tb + ob[p];
  const int n = (int)(oa[p + 1] - oa[p]), m = (int)(ob[p + 1] - ob[p]);
  int* d0 = sm;                      // diagonal d-2
  int* d1 = sm + (max_len + 1);      // diagonal d-1
  int* d2 = sm + 2 * (max_len + 1);  // diagonal d
  int ed;
  if (n == 0 || m == 0) {
    ed = n + m;
  } else {
    // diagonal 0: D[0][0] = 0 ; diagonal 1: D[0][1] = 1, D[1][0] = 1
    if (threadIdx.x == 0) { d0[0] = 0; d1[0] = 1; d1[1] = 1; }
    __syncthreads();
    for (int d = 2; d <= n + m; ++d) {
      const int ilo = max(0, d - m), ihi = min(n, d);
      for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
        const int j = d - i;
        int v;
        if (i == 0) v = j;
        else if (j == 0) v = i;
        else {
          const int sub = d0[i - 1] + (a[i - 1] != b[j - 1] ? 1 : 0);  // D[i-1][j-1]
          const int del = d1[i - 1] + 1;                              // D[i-1][j]
          const int ins = d1[i] + 1;                                  // D[i][j-1]
          v = min(sub, min(del, ins));
        }
        d2[i] = v;
      }
      __syncthreads();
      int* t = d0; d0 = d1; d1 = d2; d2 = t;
    }
    ed = d1[n];  // after the final rotation d1 is diagonal n+m, whose only cell is D[n][m]
  }
  if (threadIdx.x == 0) {
    ed_out[p] = ed;
    if (sums) {
      atomicAdd(&sums[0], (unsigned long long)ed);
      atomicAdd(&sums[1], (unsigned long long)n);
      atomicAdd(&sums[2], (unsigned long long)(ed > 0 ? 1 : 0));
    }
  }
}

}  // namespace

extern "C" int omr_pad_collate(const float* flat, const long long* offsets, const int* heights, const int* widths,
                               float* dst, int B, int Hmax, int Wmax, float pad_value, int* n_frames,
                               int height_reduction, int width_reduction, omr_stream_t stream) {
  OMR_REQUIRE(B >= 0 && Hmax >= 0 && Wmax >= 0, "omr_pad_collate: negative size");
  OMR_REQUIRE(height_reduction > 0 && width_reduction > 0, "omr_pad_collate: reductions must be positive");
  if ((long long)B * Hmax * Wmax == 0) return OMR_OK;
  OMR_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "omr_pad_collate: dst must be 16-byte aligned");
  const int W4 = (Wmax + 3) / 4;
  const long long total = (long long)B * Hmax * W4;
  long long blocks = cdiv(total, 256);
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  OmrLaunch((unsigned)blocks, 256, 0, as_stream(stream))(pad_collate_kernel, flat, offsets, heights, widths, dst, B, Hmax,
                                                        Wmax, W4, pad_value, n_frames, height_reduction, width_reduction);
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_pad_transcripts(const long long* flat, const long long* offsets, int B, int T, long long* y_in,
                                   long long* y_out, long long pad_id, omr_stream_t stream) {
  OMR_REQUIRE(B >= 0 && T >= 0, "omr_pad_transcripts: negative size");
  if ((long long)B * T == 0) return OMR_OK;
  long long blocks = cdiv((long long)B * T, 256);
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  OmrLaunch((unsigned)blocks, 256, 0, as_stream(stream))(pad_transcripts_kernel, flat, offsets, B, T, y_in, y_out, pad_id);
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_mix_argmax_step(int dt, const void* logits_a, long long lda, const void* logits_b, long long ldb, int B,
                                   int V, float alpha, long long* tok, float* val, int* finished, long long eos_id,
                                   long long pad_id, long long* out_tokens, float* out_vals, int out_ld, int step,
                                   const int* step_dev, omr_stream_t stream) {
  if (B <= 0) return OMR_OK;
  OMR_REQUIRE(V > 0, "omr_mix_argmax_step: empty vocabulary");
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)B, 256, 0, as_stream(stream))(mix_argmax_step_kernel<T>, (const T*)logits_a, lda,
                             (const T*)logits_b, ldb, V, alpha, tok, val, finished, eos_id, pad_id, out_tokens, out_vals,
                             out_ld, step, step_dev)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_levenshtein(const long long* truth, const long long* truth_offsets, const long long* hyp,
                               const long long* hyp_offsets, int pairs, int max_len, int* ed_out,
                               unsigned long long* sums, omr_stream_t stream) {
  OMR_REQUIRE(pairs >= 0 && max_len >= 0, "omr_levenshtein: negative size");
  if (pairs == 0) return OMR_OK;
  const size_t smem = sizeof(int) * 3 * ((size_t)max_len + 1);
  OMR_REQUIRE(smem <= 200 * 1024, "omr_levenshtein: sequences longer than %d tokens are not supported (got %d)",
              (int)(200 * 1024 / 12 - 1), max_len);
  if (smem > 48 * 1024) {
    OMR_CUDA(cudaFuncSetAttribute(levenshtein_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  OmrLaunch((unsigned)pairs, 256, smem, as_stream(stream))(levenshtein_kernel, truth, truth_offsets, hyp, hyp_offsets, pairs,
                                                          ed_out, sums, max_len);
  OMR_LAUNCHED();
  return OMR_OK;
}
