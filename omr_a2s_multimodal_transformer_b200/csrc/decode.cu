// decode.cu -- kernels of the batched greedy decoder (KV cache resident in HBM):
// single-query attention over a cache (split over keys + combine), KV append, first-max argmax
// with EOS bookkeeping.  All HBM-bound (K/V streaming).
#include "common.cuh"

namespace {

constexpr int HD = 64;

struct DecArgs {
  const void *q, *k, *v; void* o;
  long long q_bs, k_bs, k_rs, v_bs, v_rs, o_bs;
  const float* key_bias; long long kb_bs;
  float* ws_o; float* ws_ml;
  int B, H, Tk, window, chunk, nsplit;
  const int* pos_dev;  // when non-NULL the live key count is *pos_dev + 1 (Tk is then the sizing bound)
  float scale;
};

template <typename T>
__device__ __forceinline__ float dot64(const T* __restrict__ row, const float (&q)[HD]) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    float v[4];
    load4(row + c * 4, v);
    s = fmaf(v[0], q[c * 4], s); s = fmaf(v[1], q[c * 4 + 1], s);
    s = fmaf(v[2], q[c * 4 + 2], s); s = fmaf(v[3], q[c * 4 + 3], s);
  }
  return s;
}

// grid (B*H, nsplit), block 128
template <typename T>
__global__ void __launch_bounds__(128) attn_decode_partial_kernel(DecArgs a) {
  omr_pdl_enter();
  extern __shared__ float sc[];  // chunk scores, then 2*64 floats for the group reduction
  __shared__ float red[33];
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H, sp = blockIdx.y;
  const int tid = threadIdx.x;
  int tk = a.pos_dev ? (*a.pos_dev + 1) : a.Tk;
  if (tk > a.Tk) tk = a.Tk;
  int j_lo = 0;
  if (a.window > 0 && tk - 1 - a.window > 0) j_lo = tk - 1 - a.window;  // query position tk-1 sees keys >= tk-1-window
  int j0 = sp * a.chunk;
  int j1 = j0 + a.chunk;
  if (j0 < j_lo) j0 = j_lo;
  if (j1 > tk) j1 = tk;
  const int n = j1 > j0 ? j1 - j0 : 0;
  const T* qp = (const T*)a.q + (long long)b * a.q_bs + h * HD;
  const T* kp = (const T*)a.k + (long long)b * a.k_bs + h * HD;
  const T* vp = (const T*)a.v + (long long)b * a.v_bs + h * HD;
  const float* kb = a.key_bias ? a.key_bias + (long long)b * a.kb_bs : nullptr;
  // scores: thread = (key lane g of 16, 8-wide dim chunk c): one 16-byte load per key and thread (a warp covers four
  // whole 128-byte key rows per instruction), 8-lane shuffle reduction, four keys in flight per thread
  const int c = tid & 7, g = tid >> 3;
  float q[8];
  {
    float t0[4], t1[4];
    load4(qp + c * 8, t0);
    load4(qp + c * 8 + 4, t1);
#pragma unroll
    for (int e = 0; e < 4; ++e) { q[e] = t0[e] * a.scale; q[4 + e] = t1[e] * a.scale; }
  }
  float mx = -INFINITY;
  for (int jb = 0; jb < n; jb += 64) {
    float part[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb + u * 16 + g;
      float kv[8];
      if (j < n) {
        const T* kr = kp + (long long)(j0 + j) * a.k_rs + c * 8;
        load4(kr, *reinterpret_cast<float(*)[4]>(kv));
        load4(kr + 4, *reinterpret_cast<float(*)[4]>(kv + 4));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) kv[e] = 0.f;
      }
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(q[e], kv[e], d);
      part[u] = d;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float d = part[u];
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      const int j = jb + u * 16 + g;
      if (j < n) {
        if (kb) d += kb[j0 + j];
        if (c == 0) sc[j] = d;
        mx = fmaxf(mx, d);
      }
    }
  }
  mx = block_max(mx, red);
  const float msafe = (mx == -INFINITY) ? 0.f : mx;
  float sum = 0.f;
  for (int j = tid; j < n; j += 128) {
    float p = expf(sc[j] - msafe);
    sc[j] = p;
    sum += p;
  }
  sum = block_sum(sum, red);  // contains the __syncthreads that publishes sc[]
  // P V: thread = (key lane g of 16, 8-wide dim chunk c): 16 keys in flight per step with 16-byte loads, unrolled x4,
  // so the cache stream is bandwidth- rather than latency-bound; the 16 key lanes are then combined in shared memory
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  int j = g;
  for (; j + 48 < n; j += 64) {
    float v0[8], v1[8], v2[8], v3[8];
    const T* r0 = vp + (long long)(j0 + j) * a.v_rs + c * 8;
    const T* r1 = r0 + 16 * a.v_rs;
    const T* r2 = r1 + 16 * a.v_rs;
    const T* r3 = r2 + 16 * a.v_rs;
    load4(r0, *reinterpret_cast<float(*)[4]>(v0)); load4(r0 + 4, *reinterpret_cast<float(*)[4]>(v0 + 4));
    load4(r1, *reinterpret_cast<float(*)[4]>(v1)); load4(r1 + 4, *reinterpret_cast<float(*)[4]>(v1 + 4));
    load4(r2, *reinterpret_cast<float(*)[4]>(v2)); load4(r2 + 4, *reinterpret_cast<float(*)[4]>(v2 + 4));
    load4(r3, *reinterpret_cast<float(*)[4]>(v3)); load4(r3 + 4, *reinterpret_cast<float(*)[4]>(v3 + 4));
    const float p0 = sc[j], p1 = sc[j + 16], p2 = sc[j + 32], p3 = sc[j + 48];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p0, v0[e], fmaf(p1, v1[e], fmaf(p2, v2[e], fmaf(p3, v3[e], acc[e]))));
  }
  for (; j < n; j += 16) {
    float v0[8];
    const T* r0 = vp + (long long)(j0 + j) * a.v_rs + c * 8;
    load4(r0, *reinterpret_cast<float(*)[4]>(v0)); load4(r0 + 4, *reinterpret_cast<float(*)[4]>(v0 + 4));
    const float p0 = sc[j];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p0, v0[e], acc[e]);
  }
  float* gs = sc + a.chunk;  // [16][64]
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) gs[g * HD + c * 8 + e] = acc[e];
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int l = 0; l < 16; ++l) o += gs[l * HD + tid];
    long long w = ((long long)bh * a.nsplit + sp);
    a.ws_o[w * HD + tid] = o;
    if (tid == 0) { a.ws_ml[w * 2] = mx; a.ws_ml[w * 2 + 1] = sum; }
  }
}

// grid B*H, block 64
template <typename T>
__global__ void attn_decode_combine_kernel(DecArgs a) {
  omr_pdl_enter();
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H, d = threadIdx.x;
  float M = -INFINITY;
  for (int s = 0; s < a.nsplit; ++s) M = fmaxf(M, a.ws_ml[((long long)bh * a.nsplit + s) * 2]);
  float L = 0.f, acc = 0.f;
  for (int s = 0; s < a.nsplit; ++s) {
    long long w = (long long)bh * a.nsplit + s;
    float m = a.ws_ml[w * 2];
    float f = (m == -INFINITY) ? 0.f : expf(m - M);
    L = fmaf(a.ws_ml[w * 2 + 1], f, L);
    acc = fmaf(a.ws_o[w * HD + d], f, acc);
  }
  T* op = (T*)a.o + (long long)b * a.o_bs + h * HD;
  op[d] = from_f<T>(L > 0.f ? acc / L : 0.f);
}

template <typename T>
__global__ void kv_append_kernel(const T* __restrict__ src, long long src_rs, T* __restrict__ cache, int B, int Tmax,
                                 int width, int pos, const int* __restrict__ pos_dev) {
  omr_pdl_enter();
  if (pos_dev) pos = *pos_dev;
  if (pos < 0 || pos >= Tmax) return;
  long long n = (long long)B * width;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int c = (int)(i % width), b = (int)(i / width);
    cache[((long long)b * Tmax + pos) * width + c] = src[(long long)b * src_rs + c];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) argmax_step_kernel(const T* __restrict__ logits, long long ld, int V,
                                                          long long* __restrict__ tok, float* __restrict__ val,
                                                          int* __restrict__ finished, long long eos_id,
                                                          long long pad_id, long long* __restrict__ out_tokens,
                                                          float* __restrict__ out_vals, int out_ld, int step,
                                                          const int* __restrict__ step_dev) {
  omr_pdl_enter();
  if (step_dev) step = *step_dev;
  __shared__ float sv[256];
  __shared__ int si[256];
  const int b = blockIdx.x;
  const T* x = logits + (long long)b * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float v = to_f(x[i]);
    if (v > best) { best = v; bi = i; }  // strict > keeps the first maximum of this thread's slice
  }
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      float ov = sv[threadIdx.x + s];
      int oi = si[threadIdx.x + s];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
        sv[threadIdx.x] = ov; si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
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

}  // namespace

extern "C" int omr_attn_decode(int dt, const void* q, long long q_bs, const void* k, long long k_bs, long long k_rs,
                               const void* v, long long v_bs, long long v_rs, void* o, long long o_bs,
                               const float* key_bias, long long kb_bs, float* ws, long long ws_floats, int B, int H,
                               int Tk, int hd, float scale, int window, const int* pos_dev, omr_stream_t stream) {
  OMR_REQUIRE(hd == HD, "omr_attn_decode: head_dim must be 64 (got %d)", hd);
  OMR_REQUIRE(((q_bs | k_bs | k_rs | v_bs | v_rs | o_bs) & 3) == 0, "omr_attn_decode: strides must be multiples of 4");
  if (B <= 0 || H <= 0 || Tk <= 0) return OMR_OK;
  long long bh = (long long)B * H;
  int nsplit = (int)cdiv(148LL * 8, bh);  // ~8 resident CTAs per SM: the KV stream is latency-bound, not compute-bound
  int max_split = (int)cdiv(Tk, 128);
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  int chunk = (int)cdiv(Tk, nsplit);
  nsplit = (int)cdiv(Tk, chunk);
  OMR_REQUIRE(ws_floats >= bh * nsplit * (HD + 2), "omr_attn_decode: workspace too small (%lld < %lld floats)", ws_floats,
              bh * nsplit * (HD + 2));
  DecArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = o; a.q_bs = q_bs; a.k_bs = k_bs; a.k_rs = k_rs; a.v_bs = v_bs; a.v_rs = v_rs;
  a.o_bs = o_bs; a.key_bias = key_bias; a.kb_bs = kb_bs; a.ws_o = ws; a.ws_ml = ws + bh * nsplit * HD;
  a.B = B; a.H = H; a.Tk = Tk; a.window = window; a.chunk = chunk; a.nsplit = nsplit; a.scale = scale;
  a.pos_dev = pos_dev;
  size_t smem = sizeof(float) * (chunk + 16 * HD);
  OMR_REQUIRE(smem <= 200 * 1024, "omr_attn_decode: chunk too large");
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)bh, (unsigned)nsplit);
  OMR_DISPATCH_DT(dt, T, {
    if (smem > 48 * 1024) {
      static bool done = false;
      if (!done) {
        OMR_CUDA(cudaFuncSetAttribute(attn_decode_partial_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      200 * 1024));
        done = true;
      }
    }
    OmrLaunch(grid, 128, smem, st)(attn_decode_partial_kernel<T>, a);
    omr_count_launch();
    OmrLaunch((unsigned)bh, HD, 0, st)(attn_decode_combine_kernel<T>, a);
  });
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_kv_append(int dt, const void* src, long long src_rs, void* cache, int B, int Tmax, int width, int pos,
                             const int* pos_dev, omr_stream_t stream) {
  OMR_REQUIRE(pos_dev || (pos >= 0 && pos < Tmax), "omr_kv_append: position %d outside the cache (Tmax %d)", pos, Tmax);
  long long n = (long long)B * width;
  if (n <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)cdiv(n, 256), 256, 0, as_stream(stream))(kv_append_kernel<T>, 
                             (const T*)src, src_rs, (T*)cache, B, Tmax, width, pos, pos_dev)));
  OMR_LAUNCHED();
  return OMR_OK;
}

extern "C" int omr_argmax_step(int dt, const void* logits, long long ld, int B, int V, long long* tok, float* val,
                               int* finished, long long eos_id, long long pad_id, long long* out_tokens,
                               float* out_vals, int out_ld, int step, const int* step_dev, omr_stream_t stream) {
  if (B <= 0) return OMR_OK;
  OMR_REQUIRE(V > 0, "omr_argmax_step: empty vocabulary");
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((unsigned)B, 256, 0, as_stream(stream))(argmax_step_kernel<T>, 
                             (const T*)logits, ld, V, tok, val, finished, eos_id, pad_id, out_tokens, out_vals, out_ld,
                             step, step_dev)));
  OMR_LAUNCHED();
  return OMR_OK;
}
