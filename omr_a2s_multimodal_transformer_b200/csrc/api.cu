// api.cu -- library state (error string, launch counter) and the small elementwise/layout kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "kernels.h"

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void omr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool omr_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OMR_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
void omr_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int omr_abi_version(void) { return 5; }  // 3: omr_conv3x3_wgrad takes a scratch pointer, omr_proj_ce_* added; 4: omr_decode_layer w_o / wc_o / w2 as column slices; 5: head-major K/V caches in omr_decode_layer
extern "C" const char* omr_last_error(void) { return g_err; }
extern "C" long long omr_launch_count(void) { return g_launches.load(); }

// ---- cast -----------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long n) {
  omr_pdl_enter();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = from_f<TD>(to_f(s[i]));
}

static int grid_for(long long n, int threads, int per_thread = 4) {
  long long b = cdiv(n, (long long)threads * per_thread);
  if (b < 1) b = 1;
  if (b > 148LL * 32) b = 148LL * 32;
  return (int)b;
}

extern "C" int omr_cast(int src_dt, int dst_dt, const void* src, void* dst, long long n, omr_stream_t stream) {
  if (n <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  int g = grid_for(n, 256);
  if (src_dt == OMR_F32 && dst_dt == OMR_BF16)
    OmrLaunch(g, 256, 0, st)(cast_kernel<float, bf16>, (const float*)src, (bf16*)dst, n);
  else if (src_dt == OMR_BF16 && dst_dt == OMR_F32)
    OmrLaunch(g, 256, 0, st)(cast_kernel<bf16, float>, (const bf16*)src, (float*)dst, n);
  else if (src_dt == OMR_F32 && dst_dt == OMR_F32)
    OmrLaunch(g, 256, 0, st)(cast_kernel<float, float>, (const float*)src, (float*)dst, n);
  else if (src_dt == OMR_BF16 && dst_dt == OMR_BF16)
    OmrLaunch(g, 256, 0, st)(cast_kernel<bf16, bf16>, (const bf16*)src, (bf16*)dst, n);
  else
    OMR_REQUIRE(false, "omr_cast: bad dtypes %d -> %d", src_dt, dst_dt);
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- relu backward / add ------------------------------------------------------------------------
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, long long n) {
  omr_pdl_enter();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dx[i] = to_f(y[i]) > 0.f ? dy[i] : from_f<T>(0.f);
}
extern "C" int omr_relu_bwd(int dt, const void* y, const void* dy, void* dx, long long n, omr_stream_t stream) {
  if (n <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n, 256), 256, 0, as_stream(stream))(relu_bwd_kernel<T>, 
                             (const T*)y, (const T*)dy, (T*)dx, n)));
  OMR_LAUNCHED();
  return OMR_OK;
}

template <typename T>
__global__ void relu_mask_scale_kernel(T* __restrict__ dx, const T* __restrict__ m, float scale, long long n) {
  omr_pdl_enter();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dx[i] = to_f(m[i]) > 0.f ? from_f<T>(to_f(dx[i]) * scale) : from_f<T>(0.f);
}
int omr_relu_mask_scale(int dt, void* dx, const void* mask, float scale, long long n, cudaStream_t st) {
  if (n <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n, 256), 256, 0, st)(relu_mask_scale_kernel<T>, (T*)dx, (const T*)mask, scale, n)));
  OMR_LAUNCHED();
  return OMR_OK;
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long long n) {
  omr_pdl_enter();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) o[i] = from_f<T>(to_f(a[i]) + to_f(b[i]));
}
extern "C" int omr_add(int dt, const void* a, const void* b, void* out, long long n, omr_stream_t stream) {
  if (n <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n, 256), 256, 0, as_stream(stream))(add_kernel<T>, 
                             (const T*)a, (const T*)b, (T*)out, n)));
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- weight packing ---------------------------------------------------------------------------
// w [Co,Ci,3,3] -> transpose==0: out[co][tap][ci] ; transpose==1: out[ci][tap][co]
template <typename T>
__global__ void pack_conv_w_kernel(const float* __restrict__ w, T* __restrict__ out, int Co, int Ci, int transpose) {
  omr_pdl_enter();
  long long n = (long long)Co * Ci * 9;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int tap = (int)(i % 9);
    int ci = (int)((i / 9) % Ci);
    int co = (int)(i / (9LL * Ci));
    long long o = transpose ? ((long long)ci * 9 + tap) * Co + co : ((long long)co * 9 + tap) * Ci + ci;
    out[o] = from_f<T>(w[i]);
  }
}
extern "C" int omr_pack_conv_weight(int dt, const float* w, void* out, int Co, int Ci, int transpose,
                                    omr_stream_t stream) {
  long long n = (long long)Co * Ci * 9;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n, 256, 1), 256, 0, as_stream(stream))(pack_conv_w_kernel<T>, 
                             w, (T*)out, Co, Ci, transpose)));
  OMR_LAUNCHED();
  return OMR_OK;
}

// w [C,1,3,3] -> out [tap][c]
template <typename T>
__global__ void pack_dw_w_kernel(const float* __restrict__ w, T* __restrict__ out, int C) {
  omr_pdl_enter();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * 9) {
    int tap = i % 9, c = i / 9;
    out[tap * C + c] = from_f<T>(w[i]);
  }
}
extern "C" int omr_pack_dw_weight(int dt, const float* w, void* out, int C, omr_stream_t stream) {
  OMR_DISPATCH_DT(dt, T, (OmrLaunch((int)cdiv(C * 9, 256), 256, 0, as_stream(stream))(pack_dw_w_kernel<T>, w, (T*)out, C)));
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- PE-2D add into the fused memory, row-block copy ------------------------------------------------
template <typename T>
__global__ void pe2d_add_kernel(const T* __restrict__ x, const float* __restrict__ pe, T* __restrict__ out, int B,
                                int h, int w, int C, int pe_w, int out_rows, int row_off) {
  omr_pdl_enter();
  // one thread per 4 channels
  long long n4 = (long long)B * h * w * (C / 4);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  int c4n = C / 4;
  for (; i < n4; i += stride) {
    int c4 = (int)(i % c4n);
    long long r = i / c4n;
    int p = (int)(r % ((long long)h * w));
    int b = (int)(r / ((long long)h * w));
    int ph = p / w, pw = p % w;
    float v[4], e[4];
    load4(x + ((long long)b * h * w + p) * C + c4 * 4, v);
    load4(pe + ((long long)ph * pe_w + pw) * C + c4 * 4, e);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += e[k];
    store4(out + ((long long)b * out_rows + row_off + p) * C + c4 * 4, v);
  }
}
extern "C" int omr_pe2d_add(int dt, const void* x, const float* pe, void* out, int B, int h, int w, int C, int pe_w,
                            int out_rows, int row_off, omr_stream_t stream) {
  OMR_REQUIRE(C % 4 == 0, "omr_pe2d_add: C must be a multiple of 4 (got %d)", C);
  OMR_REQUIRE(w <= pe_w, "omr_pe2d_add: feature map wider than the PE table (%d > %d)", w, pe_w);
  long long n4 = (long long)B * h * w * (C / 4);
  if (n4 <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n4, 256, 1), 256, 0, as_stream(stream))(pe2d_add_kernel<T>, 
                             (const T*)x, pe, (T*)out, B, h, w, C, pe_w, out_rows, row_off)));
  OMR_LAUNCHED();
  return OMR_OK;
}

template <typename T>
__global__ void copy_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int rows, int C, int src_rows,
                                 int src_off) {
  omr_pdl_enter();
  long long n4 = (long long)B * rows * (C / 4);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  int c4n = C / 4;
  for (; i < n4; i += stride) {
    int c4 = (int)(i % c4n);
    long long r = i / c4n;
    int p = (int)(r % rows);
    int b = (int)(r / rows);
    float v[4];
    load4(src + ((long long)b * src_rows + src_off + p) * C + c4 * 4, v);
    store4(dst + ((long long)b * rows + p) * C + c4 * 4, v);
  }
}
extern "C" int omr_copy_rows(int dt, const void* src, void* dst, int B, int rows, int C, int src_rows, int src_off,
                             omr_stream_t stream) {
  OMR_REQUIRE(C % 4 == 0, "omr_copy_rows: C must be a multiple of 4 (got %d)", C);
  long long n4 = (long long)B * rows * (C / 4);
  if (n4 <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n4, 256, 1), 256, 0, as_stream(stream))(copy_rows_kernel<T>, 
                             (const T*)src, (T*)dst, B, rows, C, src_rows, src_off)));
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- masks -------------------------------------------------------------------------------------
__global__ void key_bias_len_kernel(float* __restrict__ bias, const int* __restrict__ lens, int B, int S, int seg_off,
                                    int seg_len, float value) {
  omr_pdl_enter();
  long long n = (long long)B * seg_len;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int j = (int)(i % seg_len), b = (int)(i / seg_len);
    bias[(long long)b * S + seg_off + j] = (j >= lens[b]) ? value : 0.f;
  }
}
extern "C" int omr_key_bias_from_lengths(float* bias, const int* lens, int B, int S, int seg_off, int seg_len,
                                         float value, omr_stream_t stream) {
  OMR_REQUIRE(seg_off >= 0 && seg_off + seg_len <= S, "omr_key_bias_from_lengths: segment out of range");
  long long n = (long long)B * seg_len;
  if (n <= 0) return OMR_OK;
  OmrLaunch((int)cdiv(n, 256), 256, 0, as_stream(stream))(key_bias_len_kernel, bias, lens, B, S, seg_off, seg_len, value);
  OMR_LAUNCHED();
  return OMR_OK;
}

__global__ void key_bias_tok_kernel(float* __restrict__ bias, const long long* __restrict__ tok, long long n,
                                    long long pad_id, float value) {
  omr_pdl_enter();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bias[i] = (tok[i] == pad_id) ? value : 0.f;
}
extern "C" int omr_key_bias_from_tokens(float* bias, const long long* tokens, long long n, long long pad_id,
                                        float value, omr_stream_t stream) {
  if (n <= 0) return OMR_OK;
  OmrLaunch((int)cdiv(n, 256), 256, 0, as_stream(stream))(key_bias_tok_kernel, bias, tokens, n, pad_id, value);
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- embedding + PE-1D --------------------------------------------------------------------------
template <typename T>
__global__ void embed_pe_kernel(const long long* __restrict__ tok, const T* __restrict__ table,
                                const float* __restrict__ pe, T* __restrict__ out, int B, int Tn, int D, int pos0,
                                const int* __restrict__ pos_dev) {
  omr_pdl_enter();
  if (pos_dev) pos0 = *pos_dev;
  long long n4 = (long long)B * Tn * (D / 4);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  int d4n = D / 4;
  for (; i < n4; i += stride) {
    int d4 = (int)(i % d4n);
    long long r = i / d4n;
    int t = (int)(r % Tn);
    long long id = tok[r];
    float v[4], e[4];
    load4(table + id * D + d4 * 4, v);
    load4(pe + (long long)(pos0 + t) * D + d4 * 4, e);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += e[k];
    store4(out + r * D + d4 * 4, v);
  }
}
extern "C" int omr_embed_pe_fwd(int dt, const long long* tokens, const void* table, const float* pe, void* out, int B,
                                int T_, int D, int pos0, const int* pos_dev, omr_stream_t stream) {
  OMR_REQUIRE(D % 4 == 0, "omr_embed_pe_fwd: D must be a multiple of 4");
  long long n4 = (long long)B * T_ * (D / 4);
  if (n4 <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n4, 256, 1), 256, 0, as_stream(stream))(embed_pe_kernel<T>, 
                             tokens, (const T*)table, pe, (T*)out, B, T_, D, pos0, pos_dev)));
  OMR_LAUNCHED();
  return OMR_OK;
}

template <typename T>
__global__ void embed_bwd_kernel(const long long* __restrict__ tok, const T* __restrict__ dout,
                                 float* __restrict__ dtable, long long rows, int D, long long padding_idx) {
  omr_pdl_enter();
  long long n = rows * D;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long r = i / D;
    int d = (int)(i % D);
    long long id = tok[r];
    if (id != padding_idx) atomicAdd(dtable + id * D + d, to_f(dout[i]));
  }
}
extern "C" int omr_embed_bwd(int dt, const long long* tokens, const void* dout, float* dtable, long long rows, int D,
                             long long padding_idx, omr_stream_t stream) {
  long long n = rows * D;
  if (n <= 0) return OMR_OK;
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n, 256, 1), 256, 0, as_stream(stream))(embed_bwd_kernel<T>, 
                             tokens, (const T*)dout, dtable, rows, D, padding_idx)));
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- column sum (bias gradients) ----------------------------------------------------------------------
// grid.x tiles columns by 32, grid.y splits rows; block (32, 8)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long rows, int N, long long ld, float* __restrict__ out) {
  omr_pdl_enter();
  __shared__ float sm[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < N) {
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8)
      acc += to_f(x[r * ld + c]);
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][threadIdx.x];
    atomicAdd(out + c, s);
  }
}
// vectorised variant: a thread owns 4 consecutive columns (one 8/16-byte load per row), 256 threads =
// (N/4 column quads) x (row lanes); lanes are combined in shared memory, one atomic per column and block
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, long long rows, int N, long long ld,
                                                         float* __restrict__ out, int rows_per_block) {
  omr_pdl_enter();
  __shared__ float sm[256 * 4];
  const int quads = N / 4;  // <= 256 and a divisor of 256
  const int lanes = 256 / quads;
  const int cq = threadIdx.x % quads, lane = threadIdx.x / quads;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long r = r0 + lane; r < r1; r += lanes) {
    float v[4];
    load4(x + r * ld + cq * 4, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) sm[(lane * quads + cq) * 4 + k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += 256) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sm[l * N + c];
    atomicAdd(out + c, s);
  }
}

extern "C" int omr_colsum(int dt, const void* x, long long rows, int N, long long ld, float* out, int accumulate,
                          omr_stream_t stream) {
  if (N <= 0) return OMR_OK;
  cudaStream_t st = as_stream(stream);
  if (!accumulate) OMR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  if (rows <= 0) return OMR_OK;
  {
    const int esz = dt == OMR_F32 ? 4 : 2;
    if (N % 4 == 0 && N <= 1024 && 256 % (N / 4) == 0 && (ld * esz) % (4 * esz) == 0 &&
        (reinterpret_cast<uintptr_t>(x) % (4 * esz)) == 0) {
      static int colsum_ctas = 0;  // CTAs per launch: every CTA ends with N atomics onto the same N addresses
      if (!colsum_ctas) {
        const char* e = getenv("OMR_COLSUM_CTAS");
        colsum_ctas = e ? atoi(e) : 148 * 4;
        if (colsum_ctas < 1) colsum_ctas = 148 * 4;
      }
      long long per = cdiv(rows, (long long)colsum_ctas);
      const long long min_per = 256 / (N / 4) * 8;
      if (per < min_per) per = min_per;
      const unsigned blocks = (unsigned)cdiv(rows, per);
      OMR_DISPATCH_DT(dt, T, (OmrLaunch(blocks, 256, 0, st)(colsum_vec_kernel<T>, (const T*)x, rows, N, ld, out, (int)per)));
      OMR_LAUNCHED();
      return OMR_OK;
    }
  }
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)(rows >= 8 * 64 ? (cdiv(rows, 8 * 16) > 512 ? 512 : cdiv(rows, 8 * 16)) : 1));
  dim3 block(32, 8);
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid, block, 0, st)(colsum_kernel<T>, (const T*)x, rows, N, ld, out)));
  OMR_LAUNCHED();
  return OMR_OK;
}

// ---- dropout (recomputable from the seed: the same call applied to dy is the backward) -------------
__device__ __forceinline__ uint32_t mix32(uint32_t a, uint32_t b) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u + (a << 6) + (a >> 2));
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// keep decision of element `key`: one hash serves the two 16-bit uniforms of keys (2m, 2m+1); dropped iff uniform < thr
__device__ __forceinline__ uint32_t drop_pair_bits(uint32_t seed, long long key) {
  return mix32(seed ^ (uint32_t)(key >> 33) * 0x632BE5ABu, (uint32_t)(key >> 1));
}
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, int C, long long per_sample,
                               uint32_t thr, float scale, uint32_t seed, int channelwise, const int* __restrict__ seed_off) {
  omr_pdl_enter();
  if (seed_off) seed += (uint32_t)(*seed_off) * 0x9E3779B9u;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long key = channelwise ? (i / per_sample) * C + (i % C) : i;
    const uint32_t u = (drop_pair_bits(seed, key) >> ((uint32_t)(key & 1) * 16)) & 0xFFFFu;
    y[i] = u < thr ? from_f<T>(0.f) : from_f<T>(to_f(x[i]) * scale);
  }
}
// same mask function, 4 elements per thread (8/16-byte accesses, two hashes); needs n % 4 == 0 and, for the
// channel-wise form, C % 4 == 0 so that a quad never straddles a sample/channel-row boundary (its first key is even)
template <typename T>
__global__ void dropout_vec_kernel(const T* __restrict__ x, T* __restrict__ y, long long n4, int C, long long per_sample,
                                   uint32_t thr, float scale, uint32_t seed, int channelwise, const int* __restrict__ seed_off) {
  omr_pdl_enter();
  if (seed_off) seed += (uint32_t)(*seed_off) * 0x9E3779B9u;
  long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i4 < n4; i4 += stride) {
    const long long i = i4 * 4;
    float v[4];
    load4(x + i, v);
    const long long key0 = channelwise ? (i / per_sample) * C + (i % C) : i;
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const uint32_t h = drop_pair_bits(seed, key0 + k);
      v[k] = (h & 0xFFFFu) < thr ? 0.f : v[k] * scale;
      v[k + 1] = (h >> 16) < thr ? 0.f : v[k + 1] * scale;
    }
    store4(y + i, v);
  }
}
// Wide form (round 2): 16-byte accesses, four of them in flight per thread, no 64-bit division in the loop.  Element-wise
// mode (CW == 0) walks the tensor as one flat array (key = element index); channel-wise mode (CW == 1) runs one grid row
// per sample with a stride that is a multiple of C, so a thread keeps ONE group of channels -- its keep factors are
// computed once and the loop only multiplies.  Same mask function as dropout_kernel (drop_pair_bits).
template <typename T> struct DropVec;
template <> struct DropVec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void apply(uint4& t, const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(__low2float(h[i]) * f[2 * i], __high2float(h[i]) * f[2 * i + 1]);
  }
};
template <> struct DropVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void apply(uint4& t, const float (&f)[4]) {
    float* v = reinterpret_cast<float*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] *= f[i];
  }
};
template <typename T, int CW>
__global__ void __launch_bounds__(256) dropout_wide_kernel(const T* __restrict__ x, T* __restrict__ y, long long per, int C,
                                                           uint32_t thr, float scale, uint32_t seed,
                                                           const int* __restrict__ seed_off) {
  omr_pdl_enter();
  constexpr int VEC = DropVec<T>::N, U = 4;
  if (seed_off) seed += (uint32_t)(*seed_off) * 0x9E3779B9u;
  const long long base = (long long)blockIdx.y * per;
  const long long step = (long long)gridDim.x * 256 * VEC;
  long long e = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC;
  auto factors = [&](long long key0, float (&f)[VEC]) {  // key0 is even
#pragma unroll
    for (int k = 0; k < VEC; k += 2) {
      const uint32_t h = drop_pair_bits(seed, key0 + k);
      f[k] = (h & 0xFFFFu) < thr ? 0.f : scale;
      f[k + 1] = (h >> 16) < thr ? 0.f : scale;
    }
  };
  float fc[VEC];
  if (CW) factors((long long)blockIdx.y * C + (e % C), fc);
  const uint4* xp = reinterpret_cast<const uint4*>(x + base);
  uint4* yp = reinterpret_cast<uint4*>(y + base);
  for (; e < per; e += step * U) {
    uint4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * step < per) r[u] = xp[(e + u * step) / VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * step < per) {
        if (CW) {
          DropVec<T>::apply(r[u], fc);
        } else {
          float f[VEC];
          factors(base + e + u * step, f);
          DropVec<T>::apply(r[u], f);
        }
        yp[(e + u * step) / VEC] = r[u];
      }
  }
}

extern "C" int omr_dropout(int dt, const void* x, void* y, long long n, int C, long long per_sample, float p,
                           long long seed, int channelwise, const int* seed_offset, omr_stream_t stream) {
  OMR_REQUIRE(p >= 0.f && p < 1.f, "omr_dropout: p must be in [0,1) (got %f)", p);
  OMR_REQUIRE(C > 0 && per_sample > 0, "omr_dropout: bad channel geometry");
  if (n <= 0) return OMR_OK;
  const uint32_t thr = (uint32_t)(p * 65536.f + 0.5f);  // 16-bit uniforms: p = 0.1 -> 0.100006, 0.25 and 0.5 exact
  {
    const int vec = dt == OMR_F32 ? 4 : 8;
    const bool al16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    const float scale = 1.f / (1.f - p);
    if (al16 && !channelwise && n % vec == 0) {
      long long blocks = (n / vec + 256 * 4 - 1) / (256 * 4);
      if (blocks > 148 * 8) blocks = 148 * 8;
      OMR_DISPATCH_DT(dt, T, (OmrLaunch(dim3((unsigned)blocks, 1), 256, 0, as_stream(stream))(dropout_wide_kernel<T, 0>, (const T*)x, (T*)y,
                                                                                              n, C, thr, scale, (uint32_t)seed, seed_offset)));
      OMR_LAUNCHED();
      return OMR_OK;
    }
    if (al16 && channelwise && C % vec == 0 && (256 * vec) % C == 0 && per_sample % vec == 0 && n % per_sample == 0 &&
        n / per_sample <= 65535 && (per_sample * (dt == OMR_F32 ? 4 : 2)) % 16 == 0) {
      const long long ns = n / per_sample;
      long long blocks = (148LL * 8 + ns - 1) / ns;
      const long long most = (per_sample / vec + 256 * 4 - 1) / (256 * 4);
      if (blocks > most) blocks = most;
      if (blocks < 1) blocks = 1;
      OMR_DISPATCH_DT(dt, T, (OmrLaunch(dim3((unsigned)blocks, (unsigned)ns), 256, 0, as_stream(stream))(
                                 dropout_wide_kernel<T, 1>, (const T*)x, (T*)y, per_sample, C, thr, scale, (uint32_t)seed, seed_offset)));
      OMR_LAUNCHED();
      return OMR_OK;
    }
  }
  {
    const int esz = dt == OMR_F32 ? 4 : 2;
    if (n % 4 == 0 && (!channelwise || (C % 4 == 0 && per_sample % 4 == 0)) && (reinterpret_cast<uintptr_t>(x) % (4 * esz)) == 0 &&
        (reinterpret_cast<uintptr_t>(y) % (4 * esz)) == 0) {
      OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n / 4, 256, 2), 256, 0, as_stream(stream))(dropout_vec_kernel<T>, 
                                 (const T*)x, (T*)y, n / 4, C, per_sample, thr, 1.f / (1.f - p), (uint32_t)seed, channelwise, seed_offset)));
      OMR_LAUNCHED();
      return OMR_OK;
    }
  }
  OMR_DISPATCH_DT(dt, T, (OmrLaunch(grid_for(n, 256), 256, 0, as_stream(stream))(dropout_kernel<T>, 
                             (const T*)x, (T*)y, n, C, per_sample, thr, 1.f / (1.f - p), (uint32_t)seed, channelwise, seed_offset)));
  OMR_LAUNCHED();
  return OMR_OK;
}
